#!/bin/bash
# Focused check of a kernel change: the tests that cover it, the HBM micro-benchmarks (new and previous kernels), then
# the whole GPU suite and one short bench line.   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh tag'
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_glue.py tests/test_gpu_kernels.py -q -x -s > $out/${tag}_pytest_kernels.log 2>&1; echo "kernel tests rc=$?"
tail -4 $out/${tag}_pytest_kernels.log
grep -h "upsample (" $out/${tag}_pytest_kernels.log | head -8
timeout 200 python tools/hbm_bench.py 32 > $out/${tag}_hbm_bench.md 2>&1; echo "hbm_bench rc=$?"
EFFQ_QA_V3=0 EFFQ_FQ_STATE_F64=1 timeout 200 python tools/hbm_bench.py 32 > $out/${tag}_hbm_bench_prev.md 2>&1; echo "hbm_bench(prev) rc=$?"
cat $out/${tag}_hbm_bench.md; grep "quantize_act\|fakequant_state" $out/${tag}_hbm_bench_prev.md
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["value"], "fp", d["fp_pass_s"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
k=d["kernels"]
print(" | ".join("%s %.1f"%(n,v["ms_per_step"]) for n,v in sorted(k.items(), key=lambda kv:-kv[1]["ms_per_step"])[:40]))
P
tail -c 2000 $out/${tag}_bench.err > $out/${tag}_bench.err.tail; rm -f $out/${tag}_bench.err

#!/bin/bash
# Quick A/B of a change: GPU tests, then bench lines with the given env settings.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag "ENV1=a ENV2=b" "ENV1=c"'
tag=${1:-q}; shift
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $out/${tag}_pytest.log
i=0
for envs in "$@"; do
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > $out/${tag}_bench_$i.json 2> $out/${tag}_bench_$i.err; echo "bench[$envs] rc=$?"
  python - <<P
import json
d=json.loads(open("$out/${tag}_bench_$i.json").read().strip().splitlines()[-1])
print("$envs", "ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["value"], "fp", d["fp_pass_s"])
k=d["kernels"]
print(" | ".join("%s %.0f"%(n,v["ms_per_step"]) for n,v in sorted(k.items(), key=lambda kv:-kv[1]["ms_per_step"])[:14]))
print(" | ".join("%s %.0f"%(l["layer"].split(".")[-3] if l["layer"].count(".")>2 else l["layer"], l["gpu_ms"]) for l in d["layers"]))
P
  tail -c 3000 $out/${tag}_bench_$i.err > $out/${tag}_bench_$i.err.tail; rm -f $out/${tag}_bench_$i.err
  i=$((i+1))
done

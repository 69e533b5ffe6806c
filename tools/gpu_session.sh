#!/bin/bash
# One gpurun call that refreshes every round-level measurement, each stage with its own timeout and its own
# files under gpurun_out/ (a failing stage does not stop the later ones).  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02'
# Stages: GPU tests | bench (1 GPU, default flags) | reference arm | evaluation throughput | reproducibility probe |
# conv micro-sweep | ncu launch list of a short bench run (only after the plain run exited 0).
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_gpu.txt 2>&1

timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 400 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 200 python tools/eval_bench.py 4 > $out/${tag}_eval_bench.log 2>&1; echo "eval_bench rc=$?"
REPRO_TAG=$tag timeout 120 python tools/repro_check.py > $out/${tag}_repro_check.log 2>&1; echo "repro_check rc=$?"
timeout 300 python tools/conv_sweep.py > $out/${tag}_conv_sweep.log 2> $out/${tag}_conv_sweep.err; echo "conv_sweep rc=$?"
if timeout 200 python bench.py --workload brats_w4a4_2x64 --steps 1 --warmup 3 --no-cpu > $out/${tag}_bench_2x64.json 2> $out/${tag}_bench_2x64.err; then
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file $out/${tag}_launches_bench_2x64.csv \
    python bench.py --workload brats_w4a4_2x64 --steps 1 --warmup 3 --no-cpu > $out/${tag}_ncu_bench.json 2> $out/${tag}_ncu_bench.err
  echo "ncu launch list rc=$?"
  python tools/summarize_launches.py $out/${tag}_launches_bench_2x64.csv > $out/${tag}_launches_bench_2x64.md 2>&1
fi
tail -3 $out/${tag}_pytest.log
cat $out/${tag}_bench_n1.json | head -c 600

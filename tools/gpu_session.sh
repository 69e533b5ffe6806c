#!/bin/bash
# Round-level measurements, each stage with its own timeout and its own SMALL files under gpurun_out/ (gpurun copies
# back at most 64 MiB: the ncu stages live in tools/gpu_session_ncu.sh).  Usage (from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/gpu_session.sh r02'
# Stages: GPU tests | bench (1 GPU, default flags, CPU arm = unmodified reference) | reference arm | LiTS line |
# evaluation throughput | reproducibility probe | conv micro-sweep | solve accuracy.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
rm -f $out/r02_one_step.txt $out/r02_spd_inverse.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 500 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 300 python bench.py --workload lits_w2a2_4x160 --steps 3 --warmup 3 --no-cpu > $out/${tag}_bench_lits.json 2> $out/${tag}_bench_lits.err; echo "lits rc=$?"
timeout 200 python tools/eval_bench.py 4 > $out/${tag}_eval_bench.log 2>&1; echo "eval_bench rc=$?"
REPRO_TAG=$tag timeout 120 python tools/repro_check.py > $out/${tag}_repro_check.log 2>&1; echo "repro_check rc=$?"
timeout 300 python tools/conv_sweep.py > $out/${tag}_conv_sweep.log 2> $out/${tag}_conv_sweep.err; echo "conv_sweep rc=$?"
timeout 120 python tools/solve_accuracy.py > $out/${tag}_solve_accuracy.txt 2>&1; echo "solve_accuracy rc=$?"
for f in $out/${tag}_*.err; do tail -c 20000 $f > $f.tail; rm -f $f; done
tail -3 $out/${tag}_pytest.log
head -c 600 $out/${tag}_bench_n1.json
du -sh $out

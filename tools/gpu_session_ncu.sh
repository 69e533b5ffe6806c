#!/bin/bash
# ncu stages (each only after the same command exited 0 without ncu).  Outputs are kept small: gpurun copies back at
# most 64 MiB.   gpurun --timeout 1800 -- 'bash tools/gpu_session_ncu.sh r02'
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
# (1) launch list of a short bench run: every launch with its device time (cold-cache, serialised: compare SHARES)
if [ -z "$NCU_SKIP_LIST" ] && timeout 200 python bench.py --workload brats_w4a4_2x64 --steps 1 --warmup 3 --no-cpu > $out/${tag}_bench_2x64.json 2> /dev/null; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file /tmp/launches.csv \
    python bench.py --workload brats_w4a4_2x64 --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
  echo "ncu launch list rc=$?"
  python tools/summarize_launches.py /tmp/launches.csv > $out/${tag}_launches_bench_2x64.md 2>&1
fi
# (2) --set full of one level-1 layer's kernels (8 volumes), one or two launches per kernel
if timeout 200 python tools/profile_layer.py 8 > $out/${tag}_profile_layer.log 2>&1; then
  timeout 500 ncu --set full --clock-control none \
    -k "regex:gram_tc_kernel|quadform_delta_kernel|potrf_tile_kernel|solve_gemm_tc_kernel|conv3d_tc_kernel|quantize_act_ndhwc|scale_search_rows|fakequant_f32_kernel|fakequant_state|glue_" \
    -c 44 -o /tmp/layer_full python tools/profile_layer.py 8 > $out/${tag}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ncu -i /tmp/layer_full.ncu-rep --page raw --csv > $out/${tag}_layer_ncu_raw.csv 2> /dev/null
  python tools/ncu_summary.py $out/${tag}_layer_ncu_raw.csv > $out/${tag}_layer_ncu.md 2>&1
  ls -la /tmp/layer_full.ncu-rep
fi
tail -c 3000 $out/${tag}_ncu_full.log > $out/${tag}_ncu_full.log.tail; rm -f $out/${tag}_ncu_full.log
du -sh $out

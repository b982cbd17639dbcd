#!/bin/bash
# sanitizer passes over every registered scenario, the one-env-per-warp intersection reset experiment, and the
# final bench lines (so that their roofline objects quote the ncu summaries committed with the same kernels)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck python tools/all_envs_probe.py 12 > gpurun_out/r2_memcheck_all_envs.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/r2_memcheck_all_envs.log
timeout 420 compute-sanitizer --tool synccheck python tools/all_envs_probe.py 12 > gpurun_out/r2_synccheck_all_envs.log 2>&1
echo "synccheck rc=$?"; tail -2 gpurun_out/r2_synccheck_all_envs.log
export QB_CONFIGS=cfg3
python tools/quick_bench.py cfg3_default 2>&1 | tee gpurun_out/r2w_variants.txt
HWYB200_RESET_WIDE=1 python tools/quick_bench.py cfg3_reset_wide 2>&1 | tee -a gpurun_out/r2w_variants.txt
unset QB_CONFIGS
bash tools/gpu_round.sh r2 bench

#!/bin/bash
# rank repair + float-key observation ranking; network: inline rectangle test, warp-per-env roundabout reset,
# group-parallel spawn lane search
mkdir -p gpurun_out
bash tools/gpu_round.sh r2v tests
V=$PWD/highwayenv_b200/csrc/variants
export QB_CONFIGS=cfg2,cfg1,cfg5
HWYB200_LIB=$V/libhwyb200_lean1.so python tools/quick_bench.py lean1 2>&1 | tee gpurun_out/r2v_variants.txt
python tools/quick_bench.py lean3_repair 2>&1 | tee -a gpurun_out/r2v_variants.txt
HWYB200_LIB=$V/libhwyb200_norepair.so python tools/quick_bench.py lean3_norepair 2>&1 | tee -a gpurun_out/r2v_variants.txt
python tools/quick_bench.py lean3_repair_again 2>&1 | tee -a gpurun_out/r2v_variants.txt
export QB_CONFIGS=cfg3,cfg4
HWYB200_LIB=$V/libhwyb200_lean2.so python tools/quick_bench.py net_lean2 2>&1 | tee -a gpurun_out/r2v_variants.txt
python tools/quick_bench.py net_new 2>&1 | tee -a gpurun_out/r2v_variants.txt
HWYB200_LIB=$V/libhwyb200_lean2.so python tools/quick_bench.py net_lean2_again 2>&1 | tee -a gpurun_out/r2v_variants.txt
python tools/quick_bench.py net_new_again 2>&1 | tee -a gpurun_out/r2v_variants.txt
bash tools/ncu_one.sh r2v_cfg2 cfg2 step_kernel 6 2

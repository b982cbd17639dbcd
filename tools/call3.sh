#!/bin/bash
# rank hints + float-key observation ranking on the lean highway kernel; which network change hurt cfg3
mkdir -p gpurun_out
bash tools/gpu_round.sh r2u tests
V=$PWD/highwayenv_b200/csrc/variants
export QB_CONFIGS=cfg2,cfg1,cfg5
HWYB200_LIB=$V/libhwyb200_lean1.so python tools/quick_bench.py lean1 2>&1 | tee gpurun_out/r2u_variants.txt
python tools/quick_bench.py lean2_hints 2>&1 | tee -a gpurun_out/r2u_variants.txt
HWYB200_LIB=$V/libhwyb200_lean1.so python tools/quick_bench.py lean1_again 2>&1 | tee -a gpurun_out/r2u_variants.txt
python tools/quick_bench.py lean2_hints_again 2>&1 | tee -a gpurun_out/r2u_variants.txt
export QB_CONFIGS=cfg3,cfg4
HWYB200_LIB=$V/libhwyb200_head.so python tools/quick_bench.py net_head 2>&1 | tee -a gpurun_out/r2u_variants.txt
python tools/quick_bench.py net_all_lean 2>&1 | tee -a gpurun_out/r2u_variants.txt
for v in RULES RECT COLLIDE; do HWYB200_LIB=$V/libhwyb200_inl_$v.so python tools/quick_bench.py net_inline_$v 2>&1 | tee -a gpurun_out/r2u_variants.txt; done

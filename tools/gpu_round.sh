#!/bin/bash
# One gpurun call: GPU tests, the default bench line, the ncu launch list of the bench command and one
# `ncu --set full` capture per BASELINE config (step + reset + observe kernels).  Outputs under gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag] [what]'     what: all | tests | bench | ncu
TAG=${1:-r2}
WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$WHAT" = all ] || [ "$WHAT" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -25
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
  echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
  timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_n1.json 2>> gpurun_out/${TAG}_bench_n1.err
fi
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 4 --warmup 3 --other-steps 4 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_bench.out 2>&1
  for cfg in cfg2 cfg3 cfg4 cfg5 cfg1; do
    # the .ncu-rep files are 30-50 MB each (gpurun_out/ is capped at 64 MiB): keep them in /tmp on the box and
    # bring back the raw-page CSV plus the per-source-line aggregation
    timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'step_kernel|reset_kernel|observe_kernel|classify' -s 12 -c 8 -f -o /tmp/${TAG}_${cfg} \
      python tools/ncu_target.py $cfg > gpurun_out/${TAG}_${cfg}_ncu.log 2>&1
    echo "ncu $cfg rc=$?"
    ncu -i /tmp/${TAG}_${cfg}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${cfg}_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_${cfg}.ncu-rep --page source --csv --print-source cuda,sass -k regex:step_kernel -c 2 \
      > /tmp/${TAG}_${cfg}_src.csv 2>/dev/null
    python tools/ncu_lines.py /tmp/${TAG}_${cfg}_src.csv 60 > gpurun_out/${TAG}_${cfg}_lines.txt 2>&1
  done
fi
ls -la gpurun_out | tail -30

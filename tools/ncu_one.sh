#!/bin/bash
# One `ncu --set full` capture of one kernel of one bench config; brings back the raw-page CSV and the per-source-line
# aggregation (the .ncu-rep itself stays on the box: 30-50 MB each against gpurun_out's 64 MiB cap).
#   bash tools/ncu_one.sh <tag> <cfg> <kernel regex> [skip] [count]
TAG=$1; CFG=$2; KRE=$3; SKIP=${4:-6}; CNT=${5:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o /tmp/${TAG} \
  python tools/ncu_target.py $CFG > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu $TAG rc=$?"
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source cuda,sass -c 1 > /tmp/${TAG}_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/${TAG}_src.csv 70 > gpurun_out/${TAG}_lines.txt 2>&1
head -3 gpurun_out/${TAG}_lines.txt
# executed counts of the CALL instructions (which call sites of a __noinline__ helper are hot)
python tools/ncu_calls.py /tmp/${TAG}_src.csv 70 > gpurun_out/${TAG}_calls.txt 2>&1

#!/bin/bash
# Turn the raw-page CSVs of tools/gpu_round.sh (gpurun_out/<tag>_cfgN_raw.csv) into the committed summaries
# profiles/r2_ncu_*.json that bench.py's roofline objects quote, and copy the per-line attributions.
#   bash tools/collect_profiles.sh <tag>
TAG=${1:-r2}
CMD="ncu --set full --clock-control none --import-source on -k regex:'step_kernel|reset_kernel|observe_kernel|classify' -s 12 -c 8 python tools/ncu_target.py"
s() { python tools/ncu_summary.py gpurun_out/${TAG}_$1_raw.csv profiles/$3 --kernel "$2" --vehicle-substeps $4 --command "$CMD $1" > /dev/null 2>&1 || echo "missing: $3"; }
s cfg2 highway_step_kernel r2_ncu_highway_step_v51.json $((4096*51*5))
s cfg1 highway_step_kernel r2_ncu_highway_step_v21.json $((4096*21*5))
s cfg5 highway_step_kernel r2_ncu_highway_step_v101.json $((8192*101*15))
s cfg4 network_step_kernel r2_ncu_network_step_roundabout.json $((8192*5*15))
s cfg3 "network_step_kernel<16" r2_ncu_network_step_intersection.json 0
s cfg3 "network_step_kernel<32" r2_ncu_network_step_intersection_g32.json 0
s cfg3 intersection_reset_kernel r2_ncu_intersection_reset.json 0
s cfg4 roundabout_reset_kernel r2_ncu_roundabout_reset.json 0
s cfg4 network_observe_kernel r2_ncu_network_observe.json 0
s cfg2 highway_reset_kernel r2_ncu_highway_reset.json 0
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do [ -f gpurun_out/${TAG}_${c}_lines.txt ] && cp gpurun_out/${TAG}_${c}_lines.txt profiles/r2_lines_${c}.txt; done
[ -f gpurun_out/${TAG}_launches_bench.csv ] && cp gpurun_out/${TAG}_launches_bench.csv profiles/r2_launches_bench.csv
[ -f gpurun_out/${TAG}_bench_n1.json ] && cp gpurun_out/${TAG}_bench_n1.json profiles/r2_bench_n1.json
[ -f gpurun_out/${TAG}_bench_reference_n1.json ] && cp gpurun_out/${TAG}_bench_reference_n1.json profiles/r2_bench_reference_n1.json
[ -f gpurun_out/${TAG}_pytest.log ] && tail -3 gpurun_out/${TAG}_pytest.log > profiles/r2_gpu_tests.txt
ls profiles | grep r2_

"""Summarise one kernel of an `.ncu-rep` (ncu --set full) into the JSON committed under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_ncu_x.json --vehicle-substeps N \
        [--kernel regex] [--command "..."] [--note "..."]

Reads `ncu -i rep --page raw --csv` (works without a GPU).  With several captured launches of the kernel the
metrics are averaged.  `--vehicle-substeps` = envs x vehicle slots doing work x substeps of one launch, for the
thread-instructions-per-vehicle-substep figure bench.py's roofline object quotes."""
from __future__ import annotations

import argparse
import csv
import io
import json
import re
import subprocess

WANTED = {
    "gpu__time_duration.sum": "gpu_time_duration_ns",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_inst_executed",
    "smsp__thread_inst_executed.sum": "thread_inst_executed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_cycles_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic_per_block",
    "launch__shared_mem_per_block_static": "smem_static_per_block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_inst",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum": "local_load_bytes",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum": "local_store_bytes",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_load_inst",
    "smsp__sass_inst_executed_op_local_st.sum": "local_store_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|"
                    r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def num(x: str):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--vehicle-substeps", type=float, default=0)
    ap.add_argument("--command", default="")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    if a.rep.endswith(".csv"):  # `ncu -i x.ncu-rep --page raw --csv` saved on the GPU box
        raw = open(a.rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    name_col = hdr.index("Kernel Name")
    data = [r for r in rows[2:] if len(r) == len(hdr) and (not a.kernel or re.search(a.kernel, r[name_col]))]
    if not data:
        raise SystemExit("no matching kernel in the report")
    out = {"kernel": data[0][name_col][:120], "launches_averaged": len(data), "command": a.command, "report": a.rep + " (scratch, not committed)"}
    stalls = {}
    for c, h in enumerate(hdr):
        vals = [num(r[c]) for r in data]
        vals = [v for v in vals if v is not None]
        if not vals:
            continue
        v = sum(vals) / len(vals)
        if h in WANTED:
            out[WANTED[h]] = v
            if h == "gpu__time_duration.sum":
                u = units[c]
                out["gpu_time_duration_ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
                del out["gpu_time_duration_ns"]
            if h.startswith("dram__bytes"):
                out[WANTED[h]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[c], 1)
        m = STALLS.match(h)
        if m:
            stalls[m.group(1) or m.group(2)] = round(v, 3)
    if stalls:
        out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
    if a.vehicle_substeps and "thread_inst_executed" in out:
        out["vehicle_substeps_per_launch"] = a.vehicle_substeps
        out["thread_inst_per_vehicle_substep"] = out["thread_inst_executed"] / a.vehicle_substeps
    elif a.vehicle_substeps and "warp_inst_executed" in out and "threads_per_warp_inst" in out:
        out["vehicle_substeps_per_launch"] = a.vehicle_substeps
        out["thread_inst_per_vehicle_substep"] = out["warp_inst_executed"] * out["threads_per_warp_inst"] / a.vehicle_substeps
    out["note"] = a.note or ("ncu replays the launch without the bench's L2 flush, so the state can stay L2-resident "
                             "between replays; durations under ncu are not bench values")
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

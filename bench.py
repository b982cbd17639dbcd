#!/usr/bin/env python
"""bench.py — env-steps/s of the batched highway hot path on B200 (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

Workload (BASELINE.json configs[1]): highway-fast-v0, vehicles_count=50 (V = 51), 3 lanes,
5 substeps per step, Kinematics observation, DiscreteMetaAction, E = 4096 envs per GPU,
i.i.d. uniform random actions from torch.Generator(1234), SameStep autoreset.  One bench
"step" = one env.step of every env of the batch.  Multi-GPU: one process per GPU (torchrun),
each rank owns a contiguous env-index range; no collective on the data path (weak scaling).

`value`      device-timed throughput through the C ABI with actions resident in HBM
             (CUDA events around every step, L2 flushed between timed steps).
`e2e`        same metric through the public API (hb.make(...).step) with HOST action buffers
             (pinned) copied H2D and obs/reward/terminated/truncated copied D2H every step.
`roofline`   dominant kernel (highway_step_kernel) against the measured HBM peak.
`cpu_baseline` / --impl reference: the CPU oracle port (oracle/hwy_oracle.c, a scalar C
             restatement of the reference's per-vehicle loop; the Python reference cannot
             travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (batched), highway-fast-v0 50 veh"
UNIT = "env-steps/s"
VEHICLES_COUNT = 50
ALGO_BYTES_PER_ENV_STEP = 9902  # SURVEY.md §8(d) cfg 2: 2*51*96 + 4 + 100 + 6
WORKLOAD = "highway-fast-v0, vehicles_count=50, 3 lanes, 5 substeps/step, Kinematics[5,5], DiscreteMetaAction"


def env_config() -> dict:
    return {"vehicles_count": VEHICLES_COUNT}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ------------------------------------------------------------------ CPU arm
def cpu_port_run(n_envs: int, steps: int, warmup: int, threads: int) -> float:
    """env-steps/s of the C oracle port (SameStep autoreset, same workload/actions scheme)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwy_oracle as ho

    from highwayenv_b200.config import default_config

    cfg = default_config("highway-fast-v0")
    cfg.update(env_config())
    cfg["_others_check_collisions"] = 0
    ob = ho.OracleBatch(ho.cfg_from_dict(cfg), n_envs, seeds=range(n_envs), threads=threads)
    ob.reset()
    rng = np.random.default_rng(1234)
    acts = rng.integers(0, 5, size=(warmup + steps, n_envs)).astype(np.int32)
    for t in range(warmup):
        ob.step(acts[t], autoreset=True)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        ob.step(acts[t], autoreset=True)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n_envs = 256 * cores  # bounded sample (thread start-up amortised over 256 envs per thread)
    t0 = time.perf_counter()
    value = cpu_port_run(n_envs, args.steps, args.warmup, cores)
    sample = (f"{n_envs} envs x {args.steps} steps (+{args.warmup} warm-up) of {WORKLOAD}, SameStep autoreset, "
              f"{cores} host threads, {cpu_model()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n_envs / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": n_envs, "autoreset": "SameStep"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import highwayenv_b200 as hb
    from highwayenv_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    E, K, W = args.envs_per_gpu, args.steps, args.warmup

    env = hb.make("highway-fast-v0", num_envs=E, config=env_config(), device=dev,
                  env_index_offset=rank * E)
    env.reset(seed=0)
    lib, P, S = env._lib, env._params, env._state
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    actions = torch.randint(0, 5, (W + K, E), generator=gen, device=dev, dtype=torch.int32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def abi_step(t):
        N.check(lib.hwy_highway_step(
            C.byref(P), C.byref(S), actions[t].data_ptr(), None, env._obs.data_ptr(), env._reward.data_ptr(),
            env._terminated.data_ptr(), env._truncated.data_ptr(), env._info_speed.data_ptr(),
            env._info_crashed.data_ptr(), N.AUTORESET_SAME_STEP, None, sp))

    # ---- warm-up
    for t in range(W):
        abi_step(t)
    barrier()

    # ---- timed: K steps, CUDA events per step, L2 flushed between steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    launches0 = lib.hwy_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        abi_step(W + k)  # ONE launch: substeps + observation + reward + SameStep autoreset
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    launches = int(lib.hwy_launch_count() - launches0)
    step_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]
    kern_ms = step_ms
    total_ms = sum(step_ms)
    kern_ms_avg = sum(kern_ms) / K

    # ---- e2e through the public API with host buffers
    h_actions = torch.empty(E, dtype=torch.int32).pin_memory()
    host_pool = torch.randint(0, 5, (K, E), dtype=torch.int32)
    h_obs = torch.empty(tuple(env._obs.shape), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(E, dtype=torch.float64).pin_memory()
    h_term = torch.empty(E, dtype=torch.bool).pin_memory()
    h_trunc = torch.empty(E, dtype=torch.bool).pin_memory()
    d_actions = torch.empty(E, dtype=torch.int32, device=dev)
    for k in range(min(3, K)):
        h_actions.copy_(host_pool[k])
        d_actions.copy_(h_actions, non_blocking=True)
        env.step(d_actions)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        h_actions.copy_(host_pool[k])  # the policy's host-side output
        d_actions.copy_(h_actions, non_blocking=True)
        obs, rew, term, trunc, _ = env.step(d_actions)
        h_obs.copy_(obs, non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_term.copy_(term, non_blocking=True)
        h_trunc.copy_(trunc, non_blocking=True)
        torch.cuda.synchronize(dev)  # the caller reads the results before acting again
    barrier()
    e2e_eager_s = time.perf_counter() - t0
    # the same loop through env.host_stepper(): upload + kernel + downloads replayed as one CUDA graph
    e2e_s, e2e_api = e2e_eager_s, "env.step + explicit pinned copies"
    try:
        hs = env.host_stepper()
        pool_np = host_pool.numpy()
        for k in range(min(3, K)):
            hs.actions[:] = pool_np[k]
            hs.step()
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            hs.actions[:] = pool_np[k]  # the policy's host-side output
            o, r_, te, tr = hs.step()   # returns after the results are in host memory
        barrier()
        e2e_graph_s = time.perf_counter() - t0
        if e2e_graph_s < e2e_s:
            e2e_s, e2e_api = e2e_graph_s, "env.host_stepper().step() (one CUDA graph: H2D + step kernel + D2H)"
    except Exception as exc:  # graph capture unavailable: keep the eager number, say why
        e2e_api += f" (host_stepper unavailable: {type(exc).__name__})"

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s, kern_ms_avg, e2e_eager_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kern_ms_avg, e2e_eager = (float(x) for x in t.cpu())
    total_launches = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_launches)

    if rank == 0:
        n_total = E * world
        value = n_total * K / (total_ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = ALGO_BYTES_PER_ENV_STEP * E / (kern_ms_avg * 1e-3) / 1e9
        cores = host_cores()
        cpu_envs, cpu_steps = 256 * cores, 20
        cpu_value = cpu_port_run(cpu_envs, cpu_steps, 3, cores) if not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{WORKLOAD}, {E} envs/GPU, SameStep autoreset (device-side, reference RNG streams)",
                "envs_per_gpu": E, "envs_total": n_total, "vehicles_per_env": VEHICLES_COUNT + 1,
                "l2": "flushed (256 MiB write) between timed steps, outside the event pairs",
                "timing": "CUDA events per step on the launch stream, summed; max over ranks",
                "parallelism": f"env-range sharding x{world}, no collective",
            },
            "clocks": clocks,
            "e2e": {
                "value": n_total * K / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": E * 4,
                "d2h_bytes_per_step": E * (env.K * 5 * 4 + 8 + 1 + 1),
                "api": e2e_api, "eager_value": n_total * K / e2e_eager,
                "note": "per GPU; pinned host actions -> device, public API, obs/reward/terminated/truncated -> pinned host, sync every step",
            },
            "gpu_launches": int(total_launches.item()),
            "roofline": {
                "bound": "hbm", "kernel": "highway_step_kernel<64>",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "peak_source": peak_src,
                "traffic_source": "profiles/r1_ncu_highway_step.json (dram__bytes_read.sum + dram__bytes_write.sum, "
                                  "one ncu --set full capture of this kernel at this size; ncu replays without the L2 "
                                  "flush, so the state stays L2-resident)",
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * E,
                "kernel_ms": kern_ms_avg,
                "note": "fp64 compute/latency bound: see DESIGN.md roofline discussion",
            },
            "cpu_baseline": None if cpu_value is None else {
                "value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cpu_envs} envs x {cpu_steps} steps of the same workload, C oracle port, {cores} threads, {cpu_model()}",
            },
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _ncu_traffic():
    """DRAM bytes per launch of the step kernel from the committed ncu capture summary (None if absent)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_ncu_highway_step.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu_arm(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched IPP environment hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU numpy port of the reference path)

One "step" = one environment timestep of every env of the batch (all A UAVs act): comm matrix,
local + global map fusion, reward, masks / uniform-random masked policy / moves, measurement and
Bayesian update at the new positions — `ipp_step` through the C ABI.  Episodes are 15 steps; the
per-episode reset (`ipp_reset`) happens inside the timed region every 15 steps.

Workload (config.workload): synthetic 50x50-cell grid (FoV 90/90, 10x10 px),
A = 4 UAVs, B = 8192 envs per GPU (BASELINE configs[2] env shape; 8 GPUs x 8192 = configs[3]'s
65536), weak scaling: every rank owns B envs, no data-path collective (envs are independent).
State per GPU = 8192 x 5 maps x 10 KB = 410 MB > 126 MB L2, so every step streams from HBM.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
EP_LEN = 15


def synthetic_params(n_agents, x_dim=50):
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as f:
        k = json.load(f)
    p = k["synthetic100" if x_dim == 100 else "synthetic50"]["params"]
    p["experiment"]["missions"]["n_agents"] = n_agents
    return p


def bytes_per_env_step(g, a):
    """SURVEY.md section 8d contract figure: read+write of the global and A local float32 beliefs
    plus one read of the uint8 ground truth."""
    return g * g * (2 * 4 * (a + 1) + 1)


# --------------------------------------------------------------------------------------------------
# CPU baseline: the numpy port of the reference path (oracle/numpy_oracle.py), one env per process
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    params, first_episode, seconds = args
    from oracle import numpy_oracle as no

    t0 = time.perf_counter()
    steps = 0
    ep = first_episode
    gen = 0.0
    while time.perf_counter() - t0 < seconds:
        g0 = time.perf_counter()
        env = no.OracleEnv(params, ep)
        gen += time.perf_counter() - g0
        for _ in range(EP_LEN):
            env.observe()
            env.act()
            steps += 1
        ep += 1
    return steps, time.perf_counter() - t0, gen


def cpu_baseline(params, seconds, procs, pool=None):
    """env-steps/s of `procs` independent single-env numpy processes (the reference is
    single-threaded numpy, so "all cores" = one process per core).  `pool`: reuse a worker pool."""
    import multiprocessing as mp

    t0 = time.perf_counter()
    if pool is None:
        with mp.get_context("fork").Pool(procs) as own:
            res = own.map(_cpu_worker, [(params, 1 + 1000 * i, seconds) for i in range(procs)], chunksize=1)
    else:
        res = pool.map(_cpu_worker, [(params, 1 + 1000 * i, seconds) for i in range(procs)], chunksize=1)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    rate = sum(r[0] / r[1] for r in res)
    single = max(r[0] / r[1] for r in res)
    return {"value": rate, "steps": steps, "wall_s": wall, "best_single_process": single}


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end] or self.rows[-3:]
        for _, line in rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = synthetic_params(args.agents, args.grid)
    procs = os.cpu_count() or 1
    # a "step" of this arm = a bounded sample: `sample_s` seconds of all-core numpy work
    per_step_s = min(args.ref_seconds, 90.0 / max(args.steps, 1))  # whole arm stays within ~2 minutes
    import multiprocessing as mp

    with mp.get_context("fork").Pool(procs) as pool:  # one pool for the whole run: no fork cost per step
        for _ in range(args.warmup):
            cpu_baseline(params, min(0.1, per_step_s), procs, pool)
        t0 = time.perf_counter()
        total_steps, rates = 0, []
        for _ in range(args.steps):
            r = cpu_baseline(params, per_step_s, procs, pool)
            total_steps += r["steps"]
            rates.append(r["value"])
        wall = time.perf_counter() - t0
    value = sum(rates) / len(rates)
    import numpy

    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (numpy)",
        "data": "synthetic",
        "config": {"workload": "numpy port of the reference env loop (oracle/numpy_oracle.py, pinned bit-for-bit "
                               "to /root/reference), single env per process, %dx%d cells, %d UAVs, random masked "
                               "policy, 15-step episodes incl. ground-truth generation" % (args.grid, args.grid,
                                                                                          args.agents),
                   "grid": args.grid, "agents": args.agents, "numpy": numpy.__version__},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "%d x %.1f s of %d independent single-env processes (%d env-steps)" % (
                             args.steps, per_step_s, procs, total_steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    params = synthetic_params(args.agents, args.grid)

    # CPU baseline first (fork before CUDA is initialised), rank 0 at N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        r = cpu_baseline(params, args.cpu_seconds, procs)
        cpu = {"value": r["value"], "unit": UNIT, "cores": procs, "kind": "port",
               "sample": "%.0f s x %d single-env numpy processes (oracle port of the reference loop, %dx%d cells, "
                         "%d UAVs): %d env-steps; best single process %.0f env-steps/s" % (
                             args.cpu_seconds, procs, args.grid, args.grid, args.agents, r["steps"],
                             r["best_single_process"])}

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from ipp_marl_b200 import BatchedIPPEnv

    B, A, G = args.envs, args.agents, args.grid
    env = BatchedIPPEnv(params, B, device=dev, env_id_base=rank * B)
    launches = {"n": 0}

    def episode_step(i, **kw):
        if i % EP_LEN == 0:
            env.reset()
            launches["n"] += 2
        env.step(**kw)
        launches["n"] += 2 if env.tables.map_stride // 4 <= 1024 else 3

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- headline: device-resident, K steps ----------------------------------------------------
    for i in range(max(args.warmup, 3)):
        episode_step(i)
    launches["n"] = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        episode_step(i)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    n_launch = launches["n"]
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- roofline: the map kernel alone, bracketed by events on its stream ------------------------
    evs = []

    def hook(phase, before):
        if phase == 2:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            evs.append(ev)

    for i in range(min(args.steps, 300)):
        if i % EP_LEN == 0:
            env.reset()
        env.step(_phase_hook=hook)
    torch.cuda.synchronize()
    kern_ms = sum(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(len(evs) // 2)) / (len(evs) // 2)
    peak, peak_src = measured_peak()
    traffic = None  # dram bytes per launch of the map kernel, from the committed ncu capture of this workload
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if (tj["envs"], tj["agents"], tj["grid"]) == (B, A, G):
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    alg = bytes_per_env_step(G, A) * B
    achieved = alg / (kern_ms * 1e-3) / 1e9

    # ---- e2e: host policy -> device env -> host rewards, copies inside the timed region ------------
    # BatchedIPPEnv.step_host = one C call (ipp_step_host): H2D copy of the policy probabilities from pinned host
    # memory, the two launches, D2H copies of rewards + chosen actions; the host synchronises every step because a
    # host policy needs the result before it can produce the next step's probabilities.
    probs_host = torch.rand((B, A, 6), dtype=torch.float32).pin_memory()
    rel_host, abs_host, act_host = env.host_results()  # one pinned block: the results come back in one copy
    stream = torch.cuda.current_stream()

    def e2e_step(i):
        if i % EP_LEN == 0:
            env.reset()
        env.step_host(probs_host, None, rel_host, abs_host, act_host)
        stream.synchronize()

    for i in range(3):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(e2e_ms.item()) * 1e-3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "agent_steps_per_sec": value * A,
            "config": {"workload": "ipp_step, %d envs/GPU x %d UAVs, %dx%d belief cells, uniform masked policy, "
                                   "15-step episodes, reset inside the timed region" % (B, A, G, G),
                       "envs_per_gpu": B, "global_envs": world * B, "agents": A, "grid": G, "parallelism": "env-sharded x%d, no data-path collective" % world,
                       "l2": "state %.0f MB per GPU > 126 MB L2 (no flush needed)" % (
                           B * (A + 1) * G * G * 4 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * A * 6 * 4,
                    "d2h_bytes_per_step": B * 4 * 2 + B * A * 4,
                    "what": "BatchedIPPEnv.step_host (C ABI ipp_step_host): policy probabilities copied from pinned "
                            "host memory, rewards + actions copied back, host synchronises every step"},
            "gpu_launches": n_launch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "%s<%d,true>" % ("step_tma_kernel" if env.step_variant == "tma" else
                                                    "step_direct_kernel", A), "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": alg,
                         "note": "dense contract bytes G^2*(8(A+1)+1) per env-step (SURVEY.md 8d)"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--warmup", type=int, default=15)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--grid", type=int, default=50, choices=[50, 100])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=2.0, help="--impl reference: CPU seconds per 'step'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched IPP environment hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the UNMODIFIED reference on the host cores)

One "step" = one environment timestep of every env of the batch (all A UAVs act): comm matrix, local + global map
fusion, reward, masks / uniform-random masked policy / moves, measurement and Bayesian update at the new positions —
`ipp_step` through the C ABI.  Episodes are 15 steps; the per-episode reset (`ipp_reset`) happens inside the timed
region every 15 steps.

Headline workload (config.workload): synthetic 50x50-cell grid (FoV 90/90, 10x10 px), A = 4 UAVs,
B = 8192 envs per GPU (BASELINE configs[2] env shape; 8 GPUs x 8192 = configs[3]'s 65536), weak scaling: every rank
owns B envs, no data-path collective on the env path (envs are independent).  State per GPU = 8192 x 5 maps x 10 KB =
410 MB > 126 MB L2, so every step streams from HBM.

Besides the contract keys the line carries (N = 1 unless noted):
  shapes        the other north-star shapes, each with value / ms_per_step / kernel_ms / frac
  train         BASELINE configs[2] (N = 1) / configs[3] (N = 8): full COMA actor+critic loop on the batched env with the
                NCCL gradient all-reduce (every N)
  cpu_baseline  the unmodified reference (baseline/_ref) single-env loop, one process per host core
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


def kat_params(n_agents, grid=50):
    """Reference parameter trees recorded by oracle/make_golden.py: the synthetic 50 / 100 m families of SURVEY.md
    section 8d and (grid 493) the untouched params.yaml default."""
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as f:
        k = json.load(f)
    p = k[{50: "synthetic50", 100: "synthetic100", 493: "default"}[grid]]["params"]
    p["experiment"]["missions"]["n_agents"] = n_agents
    return p


synthetic_params = kat_params  # older name (scripts/)


def bytes_per_env_step(g, a):
    """SURVEY.md section 8d contract figure: read+write of the global and A local float32 beliefs
    plus one read of the uint8 ground truth."""
    return g * g * (2 * 4 * (a + 1) + 1)


# --------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/ref_timing.py on baseline/_ref), else the numpy port
# --------------------------------------------------------------------------------------------------
def _port_worker(args):
    params, first_episode, seconds = args
    from oracle import numpy_oracle as no

    t0 = time.perf_counter()
    steps, ep = 0, first_episode
    while time.perf_counter() - t0 < seconds:
        env = no.OracleEnv(params, ep)
        for _ in range(EP_LEN):
            env.observe()
            env.act()
            steps += 1
        ep += 1
    return steps, time.perf_counter() - t0, 0.0


def port_rate(params, seconds, procs, pool=None):
    """env-steps/s of `procs` independent single-env processes of the numpy PORT (oracle/numpy_oracle.py)."""
    import multiprocessing as mp

    jobs = [(params, 1 + 1000 * i, seconds) for i in range(procs)]
    if pool is None:
        with mp.get_context("fork").Pool(procs) as own:
            res = own.map(_port_worker, jobs, chunksize=1)
    else:
        res = pool.map(_port_worker, jobs, chunksize=1)
    return {"value": sum(r[0] / r[1] for r in res), "steps": sum(r[0] for r in res),
            "best_single_process": max(r[0] / r[1] for r in res)}


def reference_available():
    from oracle import ref_timing

    return ref_timing.available()


def cpu_rate(params, seconds, procs, workload="env_loop", pool=None):
    """(rate dict, kind): the unmodified reference when baseline/_ref is installed, else the port."""
    if reference_available():
        from oracle import ref_timing

        return ref_timing.run(params, seconds, procs, workload=workload, pool=pool), "reference"
    return port_rate(params, seconds, procs, pool=pool), "port"


REF_WHAT = ("UNMODIFIED reference (baseline/_ref = dmar-bonn/ipp-marl marl_framework/, installed by "
            "scripts/install_ref.py): single-env loop of SURVEY.md 3.2 built from the reference's own Mapping / Agent / "
            "CommunicationLog / fuse_map / get_global_reward calls, uniform masked policy, ground-truth generation "
            "(Mapping.__init__) included, one process per host core")
PORT_WHAT = "numpy PORT of the reference env loop (oracle/numpy_oracle.py, pinned bit-for-bit to the reference)"


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
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
# reference arm
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    import numpy

    params = kat_params(args.agents, args.grid)
    procs = os.cpu_count() or 1
    have_ref = reference_available()
    # a "step" of this arm = a bounded sample: `per_step_s` seconds of all-core work on the headline workload
    per_step_s = min(args.ref_seconds, 90.0 / max(args.steps, 1))  # the timed loop stays within ~1.5 minutes
    with mp.get_context("fork").Pool(procs) as pool:  # one pool for the whole run: the workers import the reference once
        for _ in range(max(args.warmup, 1)):
            cpu_rate(params, min(0.2, per_step_s), procs, pool=pool)
        t0 = time.perf_counter()
        total_steps, rates, best = 0, [], 0.0
        for _ in range(args.steps):
            r, kind = cpu_rate(params, per_step_s, procs, pool=pool)
            total_steps += r["steps"]
            rates.append(r["value"])
            best = max(best, r["best_single_process"])
        wall = time.perf_counter() - t0
        value = sum(rates) / len(rates)
        extra = {}
        if not args.no_extra:
            # the other shapes BASELINE.md section 3 asks for, each a bounded all-core sample (not part of `value`)
            for name, a, g, wl, secs in (("g50_a2_env_loop", 2, 50, "env_loop", 2.0),
                                         ("g100_a8_env_loop", 8, 100, "env_loop", 3.0),
                                         ("g493_a4_env_loop", 4, 493, "env_loop", 6.0),
                                         ("g493_a2_env_loop", 2, 493, "env_loop", 6.0),
                                         ("g50_a4_episode_generator", 4, 50, "episode_generator", 4.0),
                                         ("g493_a2_episode_generator", 2, 493, "episode_generator", 6.0)):
                if not have_ref:
                    break
                try:
                    r, _ = cpu_rate(kat_params(a, g), secs, procs, workload=wl, pool=pool)
                    extra[name] = {"value": r["value"], "unit": UNIT, "cores": procs, "env_steps": r["steps"],
                                   "best_single_process": r["best_single_process"],
                                   "ground_truth_share": r.get("ground_truth_share")}
                except Exception as e:  # a shape that fails must not lose the headline
                    extra[name] = {"error": repr(e)[:200]}
            pr = port_rate(params, 2.0, procs, pool=pool)
            extra["port_g%d_a%d" % (args.grid, args.agents)] = {"value": pr["value"], "unit": UNIT, "cores": procs,
                                                               "what": PORT_WHAT}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (numpy)",
        "data": "synthetic",
        "config": {"workload": "%s, single env per process, %dx%d cells, %d UAVs, random masked policy, 15-step "
                               "episodes incl. ground-truth generation" % (REF_WHAT if have_ref else PORT_WHAT,
                                                                           args.grid, args.grid, args.agents),
                   "grid": args.grid, "agents": args.agents, "numpy": numpy.__version__},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference" if have_ref else "port",
                         "sample": "%d x %.1f s of %d independent single-env processes (%d env-steps); best single "
                                   "process %.0f env-steps/s" % (args.steps, per_step_s, procs, total_steps, best)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extra": extra,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def time_env_shape(torch, dist, dev, world, rank, params, B, A, G, steps, warmup, clocks_index=None, graphs=True):
    """Device-resident throughput of `ipp_step` on one shape + the map kernel bracketed alone.  Returns a dict."""
    from ipp_marl_b200 import BatchedIPPEnv

    env = BatchedIPPEnv(params, B, device=dev, env_id_base=rank * B)
    counts = {"n": 0}
    per_step = 2 if env.tables.n_cells <= 2560 else 3  # plan + map (+ reward finalize when an env spans > 1 item)

    state = {"skip": 0}

    def episode_step(i, **kw):
        """One timestep.  When a whole episode lies ahead it is issued as ONE CUDA-graph launch (ipp_run_steps: the
        same reset + 15 x (plan, map) kernels, captured once) and the next 14 calls have nothing left to issue."""
        if state["skip"]:
            state["skip"] -= 1
            return
        if i % EP_LEN == 0:
            if graphs and not kw and i + EP_LEN <= state["n_steps"]:
                env.run_steps(reset=True)
                counts["n"] += 2 + EP_LEN * per_step
                state["skip"] = EP_LEN - 1
                return
            env.reset()
            counts["n"] += 2
        env.step(**kw)
        counts["n"] += per_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if graphs:
        env.run_steps(reset=True)  # set-up, not a warm-up step: captures the episode graph (like a JIT / autotune pass)
    state["n_steps"] = max(warmup, 3)
    for i in range(max(warmup, 3)):
        episode_step(i)
    state["skip"], state["n_steps"] = 0, steps
    counts["n"] = 0
    sampler = None
    if clocks_index is not None:
        sampler = ClockSampler(clocks_index)
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for i in range(steps):
        episode_step(i)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end) if sampler is not None else None
    n_launch = counts["n"]
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())

    # the map kernel alone, bracketed by events on its stream
    evs = []

    def hook(phase, before):
        if phase == 2:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            evs.append(ev)

    for i in range(min(steps, 150)):
        if i % EP_LEN == 0:
            env.reset()
        env.step(_phase_hook=hook)
    torch.cuda.synchronize()
    kern_ms = sum(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(len(evs) // 2)) / (len(evs) // 2)
    peak, peak_src = measured_peak()
    alg = bytes_per_env_step(G, A) * B
    achieved = alg / (kern_ms * 1e-3) / 1e9
    kernel = "%s<%d,true>" % ("step_tma_kernel" if env.step_variant == "tma" else "step_direct_kernel", A)
    out = {"env": env, "value": world * B * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "kernel_ms": kern_ms,
           "kernel": kernel, "achieved": achieved, "peak": peak, "peak_src": peak_src, "alg": alg, "clocks": clocks,
           "launches": n_launch, "barrier": barrier}
    return out


def traffic_of(version, B, A, G):
    """DRAM bytes of one map-kernel launch from the committed ncu capture — only if it was taken with THIS kernel
    version on THIS shape (profiles/r02_traffic.json), else None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        tj = json.load(f)
    if tj.get("ipp_version") == version and (tj["envs"], tj["agents"], tj["grid"]) == (B, A, G):
        return tj["dram_bytes_read"] + tj["dram_bytes_write"]
    return None


def e2e_host_policy(torch, dist, dev, world, rank, params, B, steps):
    """End to end through the public API with a policy that lives on the HOST (BatchedIPPEnv.step_host = C ABI
    ipp_step_host): every step the policy output [B, A, 6] float32 is copied from pinned host memory to the device
    and rewards + chosen actions are copied back; the host waits for a step's results before it issues the next step
    of the same envs.  The batch is driven as TWO half-batches on two streams (two BatchedIPPEnv objects whose env ids
    continue each other, so the results are those of one batch of B envs — partition invariance is tested): the copies
    and the host turnaround of one half overlap the kernels of the other."""
    from ipp_marl_b200 import BatchedIPPEnv

    A = params["experiment"]["missions"]["n_agents"]
    halves = []
    Bh = B // 2
    for k in range(2):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            env = BatchedIPPEnv(params, Bh, device=dev, env_id_base=rank * B + k * Bh)
            probs = torch.rand((Bh, A, 6), dtype=torch.float32).pin_memory()
            res = env.host_results()
        halves.append({"stream": st, "env": env, "probs": probs, "res": res, "done": torch.cuda.Event()})
    torch.cuda.synchronize()

    def issue(h, i):
        with torch.cuda.stream(h["stream"]):
            if i % EP_LEN == 0:
                h["env"].reset()
            h["env"].step_host(h["probs"], None, *h["res"])
            h["done"].record()

    def run(n):
        for h in halves:
            issue(h, 0)
        for i in range(1, n):
            for h in halves:
                h["done"].synchronize()  # the host policy has this half's rewards / actions of step i - 1
                issue(h, i)
        for h in halves:
            h["done"].synchronize()

    run(6)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for h in halves:
        h["stream"].wait_event(e0)
    run(steps)
    for h in halves:
        torch.cuda.current_stream().wait_event(h["done"])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    for h in halves:
        h["env"].close()
    return {"value": world * B * steps / (float(ms.item()) * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": B * A * 6 * 4, "d2h_bytes_per_step": B * 4 * 2 + B * A * 4,
            "what": "BatchedIPPEnv.step_host (C ABI ipp_step_host) on two half-batches / two streams: policy "
                    "probabilities [B, A, 6] f32 copied from pinned host memory every step, rewards + chosen actions "
                    "copied back, the host waits for a half's results before issuing that half's next step"}


def train_leg(torch, dist, dev, world, rank, params, B, iters, bf16):
    """BASELINE configs[2] / [3]: rollout (env + feature kernels + actor forward) and COMA update on the batched env,
    gradients averaged over the ranks with the bucketed NCCL all-reduce of ipp_marl_b200/coma.py."""
    from ipp_marl_b200 import BatchedIPPEnv
    from ipp_marl_b200.coma import COMATrainer

    env = BatchedIPPEnv(params, B, device=dev, env_id_base=rank * B)
    tr = COMATrainer(env, params, minibatch=16384, data_passes=1,
                     compute_dtype=torch.bfloat16 if bf16 else torch.float32)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for it in range(2):  # warm-up: cuDNN algorithm selection, allocator growth, the one-piece first all-reduce
        tr.rollout(episodes=torch.arange(B) + (rank * B + 1))
        tr.update()
    sync()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_roll = t_upd = 0.0
    per_iter = []
    for it in range(iters):
        base = (it + 1) * world * B + rank * B + 1
        e[0].record()
        tr.rollout(episodes=torch.arange(B) + base)
        e[1].record()
        stats = tr.update()
        e[2].record()
        sync()
        t_roll += e[0].elapsed_time(e[1])
        t_upd += e[1].elapsed_time(e[2])
        per_iter.append([round(e[0].elapsed_time(e[1]), 2), round(e[1].elapsed_time(e[2]), 2)])
    # the collective alone: both networks' flat gradient buffers, bucket by bucket, nothing to overlap with
    ar_ms = 0.0
    n_opt = tr.optimizer_steps_per_update()
    grad_bytes = 4 * (tr.sync_actor.flat.numel() + tr.sync_critic.flat.numel())
    if world > 1:
        sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            for s in (tr.sync_critic, tr.sync_actor):
                for lo, hi, _m in s.buckets:
                    dist.all_reduce(s.flat[lo:hi])
        a1.record()
        sync()
        ar_ms = a0.elapsed_time(a1) / 10
    tt = torch.tensor([t_roll, t_upd, ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd, ar_ms = tt.tolist()
    steps = world * B * env.T * iters
    out = {"workload": "COMA rollout + update, %d envs/GPU x %d UAVs, 50x50, %d GPU(s)" % (B, env.A, world),
           "value": steps / ((t_roll + t_upd) * 1e-3), "unit": UNIT,
           "rollout_env_steps_per_sec": steps / (t_roll * 1e-3), "ms_rollout_per_iter": t_roll / iters,
           "ms_update_per_iter": t_upd / iters, "iters": iters, "rank0_ms_rollout_update_per_iter": per_iter,
           "data_passes": 1, "minibatch": 16384,
           "compute_dtype": "bf16 autocast" if bf16 else "fp32",
           "optimizer_steps_per_update": n_opt,
           "grad_allreduce": {"bytes_per_critic_plus_actor_step": grad_bytes if world > 1 else 0,
                              "buckets": len(tr.sync_actor.buckets) + len(tr.sync_critic.buckets),
                              "ms_alone_per_critic_plus_actor_step": ar_ms,
                              "ms_alone_per_update": ar_ms * n_opt / 2,
                              "share_of_update_if_not_overlapped": (ar_ms * n_opt / 2) / max(t_upd / iters, 1e-9),
                              "overlap": "bucket all-reduces are launched from backward hooks (async) and waited for "
                                         "before the optimizer step"},
           "critic_loss": float(stats["critic_loss"]), "actor_loss": float(stats["actor_loss"]),
           "model_tflops": tr.flops_per_update() * iters / ((t_roll + t_upd) * 1e-3) / 1e12 * world}
    env.close()
    return out


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    params = kat_params(args.agents, args.grid)

    # CPU baseline first (fork before CUDA is initialised), rank 0 at N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        r, kind = cpu_rate(params, args.cpu_seconds, procs)
        cpu = {"value": r["value"], "unit": UNIT, "cores": procs, "kind": kind,
               "sample": "%.0f s x %d single-env processes of the %s, %dx%d cells, %d UAVs: %d env-steps; best single "
                         "process %.0f env-steps/s" % (args.cpu_seconds, procs, REF_WHAT if kind == "reference" else
                                                       PORT_WHAT, args.grid, args.grid, args.agents, r["steps"],
                                                       r["best_single_process"])}
        if kind == "reference":
            pr = port_rate(params, 3.0, procs)
            cpu["port_value"] = pr["value"]
            cpu["port_what"] = PORT_WHAT

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from ipp_marl_b200 import _native

    B, A, G = args.envs, args.agents, args.grid
    head = time_env_shape(torch, dist, dev, world, rank, params, B, A, G, args.steps, args.warmup,
                          clocks_index=local_rank)
    head["env"].close()
    version = _native.load().ipp_version()
    e2e = e2e_host_policy(torch, dist, dev, world, rank, params, B, args.steps)

    shapes = {}
    if world == 1 and not args.no_shapes:
        # the other north-star shapes (BASELINE.json configs / SURVEY.md 8d); G = 493 is the reference's default grid
        for name, (b, a, g) in (("c2_1024x2_g50", (1024, 2, 50)), ("1024x4_g50", (1024, 4, 50)),
                                ("65536x4_g50", (65536, 4, 50)), ("c5_8192x8_g100", (8192, 8, 100)),
                                ("default_grid_2048x4_g493", (2048, 4, 493))):
            n = 30 if g == 493 else (60 if b * g * g > 3e8 else 150)
            r = time_env_shape(torch, dist, dev, 1, 0, kat_params(a, g), b, a, g, n, 15)
            r["env"].close()
            shapes[name] = {"value": r["value"], "unit": UNIT, "agent_steps_per_sec": r["value"] * a,
                            "ms_per_step": r["ms_per_step"], "kernel": r["kernel"], "kernel_ms": r["kernel_ms"],
                            "achieved_gbs": r["achieved"], "frac": r["achieved"] / r["peak"], "steps": n,
                            "state_mb": b * (a + 1) * g * g * 4 / 1e6}
            torch.cuda.empty_cache()

    train = None
    if not args.no_train:
        train = train_leg(torch, dist, dev, world, rank, kat_params(4, 50), args.train_envs, args.train_iters,
                          args.train_bf16)

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "agent_steps_per_sec": head["value"] * A,
            "config": {"workload": "ipp_step, %d envs/GPU x %d UAVs, %dx%d belief cells, uniform masked policy, "
                                   "15-step episodes, reset inside the timed region" % (B, A, G, G),
                       "envs_per_gpu": B, "global_envs": world * B, "agents": A, "grid": G,
                       "parallelism": "env-sharded x%d, no data-path collective" % world,
                       "l2": "state %.0f MB per GPU > 126 MB L2 (no flush needed)" % (B * (A + 1) * G * G * 4 / 1e6)},
            "clocks": head["clocks"],
            "e2e": e2e,
            "gpu_launches": head["launches"],
            "roofline": {"bound": "hbm", "achieved": head["achieved"], "peak": head["peak"], "unit": "GB/s",
                         "frac": head["achieved"] / head["peak"], "traffic": traffic_of(version, B, A, G),
                         "peak_source": head["peak_src"], "kernel": head["kernel"], "kernel_ms": head["kernel_ms"],
                         "algorithmic_bytes_per_launch": head["alg"], "ipp_version": version,
                         "note": "achieved = dense contract bytes G^2*(8(A+1)+1) per env-step (SURVEY.md 8d) / kernel "
                                 "time: an algorithmic-bytes throughput, not a DRAM bandwidth — the kernel is "
                                 "footprint-sparse and moves fewer bytes (`traffic`, ncu), so frac can exceed 1"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if shapes:
            line["shapes"] = shapes
        if train is not None:
            line["train"] = train
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
    ap.add_argument("--grid", type=int, default=50, choices=[50, 100, 493])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=2.0, help="--impl reference: CPU seconds per 'step'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shapes", action="store_true", help="skip the block of other north-star shapes (N = 1)")
    ap.add_argument("--no-train", action="store_true", help="skip the COMA train leg (configs[2] / configs[3])")
    ap.add_argument("--no-extra", action="store_true", help="--impl reference: skip the extra reference shapes")
    ap.add_argument("--train-envs", type=int, default=8192, help="envs per GPU of the train leg")
    ap.add_argument("--train-iters", type=int, default=2)
    ap.add_argument("--train-bf16", action="store_true", help="train leg with bf16 autocast (default: fp32 like the reference)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""Per-kernel times of the fused step (development aid): plan kernel, map kernel and reset, CUDA-event bracketed on
the launching stream, for the BASELINE shapes.  Usage: python scripts/kernel_times.py [BxAxG ...] [direct]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv  # noqa: E402

kats = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))
shapes = [a for a in sys.argv[1:] if "x" in a] or ["8192x4x50", "1024x4x50", "1024x2x50", "65536x4x50", "8192x8x100"]
variant = "direct" if "direct" in sys.argv[1:] else "tma"
for s in shapes:
    B, A, G = (int(v) for v in s.split("x"))
    params = kats["synthetic100" if G == 100 else "synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = A
    env = BatchedIPPEnv(params, B, device="cuda:0")
    env.set_step_variant(variant)
    env.reset()
    for _ in range(15):
        env.step()
    torch.cuda.synchronize()
    evs = {1: [], 2: [], 0: []}

    def hook(phase, before):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        evs[phase].append(ev)

    n_ep = 6
    for ep in range(n_ep):
        hook(0, True)
        env.reset()
        hook(0, False)
        for _ in range(15):
            env.step(_phase_hook=hook)
    torch.cuda.synchronize()

    def mean_us(lst):
        v = [lst[2 * i].elapsed_time(lst[2 * i + 1]) for i in range(len(lst) // 2)]
        return 1e3 * sum(v) / len(v), v

    plan_us, _ = mean_us(evs[1])
    map_us, mv = mean_us(evs[2])
    reset_us, _ = mean_us(evs[0])
    per_t = [1e3 * sum(mv[t::15]) / len(mv[t::15]) for t in range(15)]
    # whole steps without hooks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for ep in range(n_ep):
        env.reset()
        for _ in range(15):
            env.step()
    e1.record()
    torch.cuda.synchronize()
    step_us = 1e3 * e0.elapsed_time(e1) / (15 * n_ep)
    byt = env.algorithmic_bytes_per_env_step() * B
    print("%s %s: step %.1f us (%.2f M env-steps/s) | map %.1f us (%.0f GB/s dense-contract) plan %.1f reset %.1f | map per t: %s"
          % (variant, s, step_us, B / step_us, map_us, byt / map_us / 1e3, plan_us, reset_us,
             " ".join("%.0f" % v for v in per_t)), flush=True)
    del env

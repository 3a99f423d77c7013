"""Timing experiments on the TMA map kernel (IPP_TMA_DEBUG knobs; results of those runs are not valid maps).
Needs a library built with -DIPP_TMA_TIMING_KNOBS (add it to NVCC_FLAGS in ipp_marl_b200/build.py); the product
build compiles the knobs out."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open(os.path.join(sys.path[0], "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
for dbg in sys.argv[1:] or ["0", "1", "2", "4"]:
    os.environ["IPP_TMA_DEBUG"] = dbg
    env.reset()
    for _ in range(15): env.step()
    torch.cuda.synchronize()
    evs = []
    def hook(phase, before):
        if phase == 2:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); evs.append(ev)
    for ep in range(6):
        env.reset()
        for _ in range(15): env.step(_phase_hook=hook)
    torch.cuda.synchronize()
    ms = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(len(evs) // 2)]
    per_t = [sum(ms[t::15]) / len(ms[t::15]) * 1e3 for t in range(15)]
    print("dbg=%s map kernel mean %.1f us; per timestep: %s" % (dbg, sum(ms) / len(ms) * 1e3, " ".join("%.0f" % v for v in per_t)))

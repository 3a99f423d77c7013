"""Timing of the plan kernel and its phases (IPP_PLAN_DEBUG: 2 = no planning (agents stay), 4 = no code generation)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open(os.path.join(sys.path[0], "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
for dbg in sys.argv[1:] or ["0", "4", "2"]:
    if dbg.startswith("e"):  # e16 / e28: envs per block override
        os.environ["IPP_PLAN_EPB"] = dbg[1:]
        dbg = "0"
    os.environ["IPP_PLAN_DEBUG"] = dbg
    evs = []
    def hook(phase, before):
        if phase == 1:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); evs.append(ev)
    for ep in range(8):
        env.reset()
        for _ in range(15): env.step(_phase_hook=hook)
    torch.cuda.synchronize()
    ms = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(30, len(evs) // 2)]
    print("IPP_PLAN_DEBUG=%s EPB=%s plan kernel mean %.1f us" % (dbg, os.environ.get("IPP_PLAN_EPB", "auto"), sum(ms) / len(ms) * 1e3))

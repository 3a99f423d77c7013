"""Plan kernel timed alone (development aid): CUDA-event bracket of IPP_PHASE_MOVE over a few episodes."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv
kats = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))
for s in [a for a in sys.argv[1:] if "x" in a] or ["8192x4x50"]:
    B, A, G = (int(v) for v in s.split("x"))
    params = kats["synthetic100" if G == 100 else "synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = A
    env = BatchedIPPEnv(params, B, device="cuda:0")
    env.reset()
    for _ in range(15): env.step()
    evs = []
    def hook(phase, before):
        if phase == 1:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); evs.append(ev)
    for ep in range(6):
        env.reset()
        for _ in range(15): env.step(_phase_hook=hook)
    torch.cuda.synchronize()
    ms = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(len(evs) // 2)]
    print("%s plan kernel mean %.1f us (event bracket)" % (s, 1e3 * sum(ms) / len(ms)), flush=True)

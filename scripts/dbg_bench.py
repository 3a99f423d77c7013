"""Timing experiments on the TMA map kernel (IPP_TMA_DEBUG knobs; results of runs with bits 1/2/4 are not valid maps).
Needs a library built with the knobs:  IPP_NVCC_EXTRA="-DIPP_TMA_TIMING_KNOBS" python scripts/dbg_bench.py 0 8 1
(bit 0 = load pipeline only, 1 = no global map, 2 = no local maps, 3 (8) = dense loads of the local maps);
the product build compiles the knobs out."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if "--child" not in sys.argv:
    for dbg in [a for a in sys.argv[1:]] or ["0", "1", "2", "4", "8"]:
        env = dict(os.environ, IPP_TMA_DEBUG=dbg)
        subprocess.run([sys.executable, __file__, "--child", dbg], env=env)
    sys.exit(0)
import torch
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv
dbg = sys.argv[2]
params = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
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
print("dbg=%s map kernel mean %.1f us; per timestep: %s" % (dbg, sum(ms) / len(ms) * 1e3, " ".join("%.0f" % v for v in per_t)), flush=True)

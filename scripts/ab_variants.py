"""Development aid: step the TMA and the direct map kernel side by side and report where their maps differ."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv
kats = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))
for name, A, B in (("synthetic50", 4, 3), ("synthetic100", 8, 2)):
    params = kats[name]["params"]; params["experiment"]["missions"]["n_agents"] = A
    envs = []
    for v in ("direct", "tma"):
        e = BatchedIPPEnv(params, B, device="cuda:0"); e.set_step_variant(v); e.reset(); envs.append(e)
    for t in range(15):
        for e in envs: e.step()
        torch.cuda.synchronize()
        d, m = envs
        lo = (d._local != m._local).cpu().numpy(); go = (d._glob != m._glob).cpu().numpy()
        fl = (d._flags != m._flags).cpu().numpy()
        rw = (d.reward_rel - m.reward_rel).abs().max().item()
        if lo.any() or go.any() or fl.any() or rw > 0:
            print(name, "t", t, "local diff cells", lo.sum(), "global", go.sum(), "flags", fl.sum(), "reward", rw)
            for b in range(B):
                for i in range(A):
                    idx = np.flatnonzero(lo[b, i])
                    if idx.size: print("  env", b, "local", i, "tiles", sorted(set((idx // 128).tolist())), "n", idx.size,
                                       "comm", int(d.comm[b, i]), "pos", d.positions[t][b, i].tolist(), d.positions[t+1][b, i].tolist())
                idx = np.flatnonzero(go[b])
                if idx.size: print("  env", b, "global tiles", sorted(set((idx // 128).tolist())), "n", idx.size)
            break
    else:
        print(name, "identical over 15 steps")

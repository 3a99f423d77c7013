"""Development aid: time the step with parts of plan_kernel disabled (IPP_PLAN_DEBUG bits: 2 = no moves, 4 = no codes)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open(os.path.join(sys.path[0], "tests/golden/kats.json")))["synthetic50"]["params"]
B = 8192
env = BatchedIPPEnv(params, B, device="cuda:0")
for dbg in ("0", "2", "4", "6"):
    os.environ["IPP_PLAN_DEBUG"] = dbg
    env.reset()
    for _ in range(15): env.step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for ep in range(10):
        env.reset()
        for _ in range(15): env.step()
    e1.record(); torch.cuda.synchronize()
    print("IPP_PLAN_DEBUG=%s: %.1f us/step" % (dbg, e0.elapsed_time(e1) / 150 * 1e3))

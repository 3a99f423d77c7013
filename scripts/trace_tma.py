"""Development aid: per-item timeline of CTA 0 of the TMA map kernel.  Needs a build with the trace hooks:
    IPP_NVCC_EXTRA="-DIPP_TMA_TRACE" python scripts/trace_tma.py
Columns (us, relative to the producer's start on item 0): producer starts item | ring space granted | all copies
issued | first consumer sees the item landed | last tile task has read its quads | finisher done."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv, _native
params = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
lib = _native.load()
env.reset()
for _ in range(6): env.step()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 512)()
lib.ipp_debug_tma_trace(buf, 1)
env.step()
torch.cuda.synchronize()
lib.ipp_debug_tma_trace(buf, 0)
raw = np.array(buf[:], dtype=np.float64).reshape(64, 8)
print("ns waited for the record, summed over the item's 20 tile tasks:", raw[:56, 6].astype(int).tolist())
a = raw[:56, :6]
t0 = a[0, 0]
a = (a - t0) / 1e3
np.set_printoptions(precision=2, suppress=True, linewidth=200)
print("item  start  space  issued  landed(first seen)  read(last)  finished | issue->landed  landed->read")
for k in range(56):
    print("%3d  %6.2f %6.2f %6.2f   %6.2f   %6.2f   %6.2f | %5.2f %5.2f" % (k, *a[k], a[k, 3] - a[k, 2], a[k, 4] - a[k, 3]))

"""A/B of the host-policy step paths on one box: python-level copies + step(probs=) vs the single C call step_host."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open(os.path.join(sys.path[0], "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
B, A = 8192, 4
env = BatchedIPPEnv(params, B, device="cuda:0")
probs_host = torch.rand((B, A, 6), dtype=torch.float32).pin_memory()
probs_dev = torch.empty((B, A, 6), dtype=torch.float32, device="cuda:0")
rel_host, abs_host, act_host = env.host_results()
stream = torch.cuda.current_stream()

def py_step(i):
    if i % 15 == 0: env.reset()
    probs_dev.copy_(probs_host, non_blocking=True)
    env.step(probs=probs_dev)
    rel_host.copy_(env.reward_rel, non_blocking=True)
    abs_host.copy_(env.reward_abs, non_blocking=True)
    act_host.copy_(env.actions, non_blocking=True)
    stream.synchronize()

def c_step(i):
    if i % 15 == 0: env.reset()
    env.step_host(probs_host, None, rel_host, abs_host, act_host)
    stream.synchronize()

def dev_step(i):
    if i % 15 == 0: env.reset()
    env.step()

for name, fn, sync in (("device-only", dev_step, False), ("python copies", py_step, True), ("step_host", c_step, True)) * 2:
    for i in range(30): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 600
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-14s %.1f us/step  %.2f M env-steps/s" % (name, dt / n * 1e6, B * n / dt / 1e6))

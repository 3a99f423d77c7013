"""Quick kernel timing (development aid; bench.py is the contract benchmark)."""
import json, sys, time
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ipp_marl_b200 import BatchedIPPEnv
kats = json.load(open(__import__("os").path.join(sys.path[0], "tests/golden/kats.json")))
params = kats["synthetic50"]["params"]
import itertools
only = sys.argv[1:]
# BASELINE configs: c2 (1024 x 2), c3/c4 shape (8192 x 4 per GPU), 65536 x 4 on one GPU, c5 (8192 x 8 UAVs, 100x100)
for (B, A, G), variant in itertools.product([(1024, 2, 50), (8192, 2, 50), (8192, 4, 50), (65536, 4, 50), (8192, 8, 100)],
                                            ["direct", "tma"]):
    if only and variant not in only: continue
    params = kats["synthetic100" if G == 100 else "synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = A
    env = BatchedIPPEnv(params, B, device="cuda:0")
    env.set_step_variant(variant)
    env.reset()
    for _ in range(15): env.step()
    torch.cuda.synchronize()
    n_ep = 10
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for ep in range(n_ep):
        env.reset()
        for _ in range(15): env.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = n_ep * 15
    sps = B * steps / (ms * 1e-3)
    byt = env.algorithmic_bytes_per_env_step()
    print(f"{variant:6s} B={B} A={A} G={G}: {ms/steps*1e3:.1f} us/step  {sps:.3e} env-steps/s  {sps*byt/1e9:.0f} GB/s algorithmic")
    del env

# split mode with the network-input feature builders (observe -> features_actor -> act -> features_critic)
params = kats["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
env.reset()
obs = None
for _ in range(15):
    env.observe(); obs = env.features_actor(obs); env.act(); st = env.features_critic(obs)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for ep in range(10):
    env.reset()
    for _ in range(15):
        env.observe(); obs = env.features_actor(obs); env.act(); st = env.features_critic(obs, st)
e1.record(); torch.cuda.synchronize()
print(f"split+features B=8192 A=4: {e0.elapsed_time(e1)/150*1e3:.1f} us/step")

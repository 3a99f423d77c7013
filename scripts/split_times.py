import json, os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo/ipp_marl_b200") else os.getcwd())
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open("tests/golden/kats.json"))["synthetic50"]["params"]; params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
obs = st = None
env.reset()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
tot = [0.0] * 4
for ep in range(4):
    env.reset()
    for _ in range(15):
        ev[0].record(); env.observe(); ev[1].record(); obs = env.features_actor(obs); ev[2].record(); env.act(); ev[3].record(); st = env.features_critic(obs, st); ev[4].record()
        torch.cuda.synchronize()
        if ep: 
            for i in range(4): tot[i] += ev[i].elapsed_time(ev[i + 1])
print("split mode us/step: observe %.1f features_actor %.1f act %.1f features_critic %.1f total %.1f" % tuple([1e3 * t / 45 for t in tot] + [1e3 * sum(tot) / 45]))

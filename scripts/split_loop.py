"""Development aid: a few episodes in split mode (observe -> features_actor -> ig_plan -> act -> features_critic ->
eval metrics) so that ncu can capture the feature / planner / metric kernels."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipp_marl_b200 import BatchedIPPEnv
params = json.load(open(os.path.join(ROOT, "tests/golden/kats.json")))["synthetic50"]["params"]
params["experiment"]["missions"]["n_agents"] = 4
env = BatchedIPPEnv(params, 8192, device="cuda:0")
obs = st = None
for ep in range(2):
    env.reset()
    for _ in range(15):
        env.observe(); obs = env.features_actor(obs); acts = env.ig_plan(); env.act(actions=acts); st = env.features_critic(obs, st)
    env.observe(final=True); env.eval_metrics()
torch.cuda.synchronize()
print("ok")

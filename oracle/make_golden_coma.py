"""Golden values of the reference's learner arithmetic, written to
tests/golden/coma_kats.npz.  Build-container only: ``python -m oracle.make_golden_coma``.

Networks are NOT stored: both sides construct them under ``torch.manual_seed(seed)`` and the layer
construction order is the same (conv1, conv2, conv3, fc1, fc2, fc3), so the default initialisation
consumes the RNG identically; the fixture pins that too (first logits of both networks).
"""
import os

import numpy as np
import torch

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "coma_kats.npz")


class _Writer:
    def __getattr__(self, name):
        return lambda *a, **k: None


def main():
    rh.load()
    from marl_framework.actor.network import ActorNetwork
    from marl_framework.actor.learner import ActorLearner
    from marl_framework.critic.network import CriticNetwork
    from marl_framework.critic.learner import CriticLearner
    from marl_framework.agent.state_space import AgentStateSpace
    from batch_memory import BatchMemory
    from utils.utils import TransitionCOMA

    params = rh.synthetic_params(50, 4)
    seed, T, A = 123, 15, 4
    out = {"seed": np.array(seed)}
    # ---- TD(lambda) targets of one episode (batch_memory.py:120-162) with a seeded critic ----
    torch.manual_seed(seed)
    critic = CriticNetwork(params)
    g = torch.Generator().manual_seed(7)
    states = torch.rand((T, A, 11, 11, 12), generator=g)
    actions = torch.randint(0, 6, (T, A), generator=g)
    rewards = torch.rand((T,), generator=g) - 0.3
    mem = BatchMemory(params, None)
    for t in range(T):
        for a in range(A):
            mem.add(a, state=states[t, a], observation=states[t, a, :, :, :7], action=actions[t, a].reshape(1),
                    mask=torch.ones(6), reward=float(rewards[t]), done=(t == T - 1))
    mem.build_td_targets(critic)
    td = np.array([[float(mem.get(t, a, "td_target")) for a in range(A)] for t in range(T)])
    with torch.no_grad():
        q0, _ = critic.forward(states[0, 0])
    out.update(td_states=states.numpy(), td_actions=actions.numpy(), td_rewards=rewards.numpy(), td_targets=td,
               critic_q0=q0.numpy())
    # ---- critic / actor mini-batch losses (critic/learner.py:76-99, actor/learner.py:52-101) ----
    torch.manual_seed(seed + 1)
    actor = ActorNetwork(params)
    torch.manual_seed(seed + 2)
    critic2 = CriticNetwork(params)
    actor.device = torch.device("cpu")
    N = 60
    obs = torch.rand((N, 11, 11, 7), generator=g, dtype=torch.float64)
    st = torch.rand((N, 11, 11, 12), generator=g)
    act = torch.randint(0, 6, (N, 1), generator=g)
    masks = (torch.rand((N, 6), generator=g) > 0.25).float()
    masks[torch.arange(N), act[:, 0]] = 1.0
    tdt = torch.rand((N, 1), generator=g) - 0.2
    batch = [TransitionCOMA(st[i], obs[i], act[i], masks[i], 0.0, False, tdt[i], torch.tensor([0.0])) for i in range(N)]
    eps = 0.3
    with torch.no_grad():
        probs0, _ = actor.forward(obs.float(), eps)
        q_before, _ = critic2.forward(st)
    cl = CriticLearner(params, _Writer(), critic2)
    cl.device = torch.device("cpu")
    cl.critic.to("cpu")
    all_q, cstats = cl.learn(1, [batch], 1)
    al = ActorLearner(params, _Writer(), actor, AgentStateSpace(params))
    al.device = torch.device("cpu")
    al.actor.to("cpu")
    _, astats = al.learn([batch], all_q, eps)
    out.update(lb_obs=obs.numpy(), lb_state=st.numpy(), lb_act=act.numpy(), lb_masks=masks.numpy(),
               lb_td=tdt.numpy(), lb_eps=np.array(eps), actor_probs0=probs0.numpy(), critic_q_before=q_before.numpy(),
               critic_loss=np.array(float(cstats[0])), critic_q_after=all_q[0].detach().numpy(),
               actor_loss=np.array(float(astats[0])), actor_adv_mean=np.array(float(astats[1])))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()

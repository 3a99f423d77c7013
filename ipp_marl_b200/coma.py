"""Batched COMA training loop on top of BatchedIPPEnv.

Stays PyTorch (tensor cores through cuDNN / cuBLAS): the actor / critic CNNs are the reference's
architectures (actor/network.py:19-28, critic/network.py:18-26 — `fc2` is constructed but unused there;
kept so that the reference's layer names / checkpoints map one to one), the counterfactual-baseline policy gradient follows
actor/learner.py:52-101 and the critic regression critic/learner.py:76-99.  What changes is the shape
of the data: B environments are rolled out in lock-step on the GPU (env kernels + feature kernels),
the actor runs once per timestep on [B*A, 11, 11, 7], and TD(lambda) targets are computed for all
B*A trajectories at once.  Multi-GPU: one process per GPU, env batch sharded by rank, ONE flattened
NCCL all-reduce of the gradients per network per optimizer step.

Differences from the reference, all deliberate and listed:
  * the reference concatenates 5 episodes per agent into one list and its TD(lambda) code leaks across
    the episode boundaries (batch_memory.py:129,145-148: first transition of episodes 2..5 gets target 0;
    bootstraps from the next episode's first state).  Here every (env, agent) trajectory is one episode,
    i.e. the reference's formula for the FIRST episode of a batch — `td_lambda_targets` below, checked
    against the reference's BatchMemory.build_td_targets in tests/test_coma.py;
  * `reference_frozen_target=True` reproduces the reference's never-updated target critic
    (coma_mission.py:90 hands build_td_targets the construction-time copy, coma_wrapper.py:34); set it
    to False for the conventional hard update every `copy_rate` updates;
  * mini-batches are large (default 8192 transitions) instead of 60;
  * one rollout collects B * world episodes, so the epsilon schedule (annealed per EPISODE in the reference,
    actor/network.py:53-58) advances by that many episodes per rollout (`COMATrainer.episodes_per_rollout`);
  * arithmetic is float32 like the reference by default; bf16 autocast is opt-in (`compute_dtype`);
  * checkpoints are state dicts, not pickled modules (`mission.save_actor` / `mission.load_actor_state` reads both).
"""
import math

import torch
from torch import nn

N_ACTIONS = 6


class _Trunk(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 256, (5, 5))
        self.conv2 = nn.Conv2d(256, 256, (4, 4))
        self.conv3 = nn.Conv2d(256, 256, (4, 4))
        self.fc1 = nn.Linear(256, 256)
        self.fc2 = nn.Linear(256, 256)  # constructed but unused, as in the reference
        self.fc3 = nn.Linear(256, N_ACTIONS)

    def logits(self, x):
        """x: [N, P, P, C] channel-last like the reference's feature tensors (network.py:70-74)."""
        x = x.permute(0, 3, 1, 2)
        x = torch.relu(self.conv1(x))
        x = torch.relu(self.conv2(x))
        x = torch.relu(self.conv3(x))
        h = x.flatten(1)
        return self.fc3(torch.relu(self.fc1(h)))


class ActorNet(_Trunk):
    """actor/network.py:10-88: eps-mixed softmax policy over the 6 actions."""

    def __init__(self):
        super().__init__(7)

    def forward(self, obs, eps):
        probs = torch.softmax(self.logits(obs).float(), dim=-1)
        return (1 - eps) * probs + eps / N_ACTIONS


class CriticNet(_Trunk):
    """critic/network.py:12-47: Q(s, .) for the agent whose state tensor is given."""

    def __init__(self):
        super().__init__(12)

    def forward(self, state):
        return self.logits(state).float()


def epsilon(episode, eps_max=0.5, eps_min=0.02, anneal=10000):
    """actor/network.py:53-58."""
    if episode > anneal:
        return eps_min
    return eps_max - episode / anneal * (eps_max - eps_min)


def td_lambda_targets(rewards, q_taken, gamma, lam):
    """TD(lambda) targets as the reference codes them for one episode (batch_memory.py:120-162).

    rewards [..., T], q_taken [..., T] = Q_target(s_t)[a_t].  For every t:
        G(n) = sum_{l<n} gamma^l r_{t+l}  +  [t+n <= T-2] gamma^n q_{t+n}       n = 1 .. T-t
        target_t = (1 - lam) * sum_n lam^(n-1) G(n)
    (the last transition is never bootstrapped from: `done[t+n] or t+n+1 >= len`, :145-148; the weights
    are not renormalised for the truncated tail — as coded).
    """
    T = rewards.shape[-1]
    dev, dt = rewards.device, torch.float64
    r = rewards.to(dt)
    q = q_taken.to(dt)
    t_idx = torch.arange(T, device=dev)
    n_idx = torch.arange(1, T + 1, device=dev)
    # discounted reward prefix sums: S[t, n] = sum_{l<n} gamma^l r_{t+l}
    l_idx = torch.arange(T, device=dev)
    src = t_idx[:, None] + l_idx[None, :]                       # t + l
    valid = src < T
    disc = (gamma ** l_idx.to(dt))[None, :] * valid
    gathered = r[..., src.clamp(max=T - 1)] * disc               # [..., T(t), T(l)]
    S = torch.cumsum(gathered, dim=-1)                           # S[..., t, n-1]
    tn = t_idx[:, None] + n_idx[None, :]                         # t + n
    n_ok = tn <= T                                               # n <= T - t
    boot_ok = tn <= T - 2
    boot = q[..., tn.clamp(max=T - 1)] * (gamma ** n_idx.to(dt))[None, :] * boot_ok
    G = (S + boot) * n_ok
    w = (lam ** (n_idx - 1).to(dt))[None, :]
    return ((1 - lam) * (G * w).sum(-1)).to(rewards.dtype)


def coma_actor_loss(probs, q_values, actions, masks):
    """actor/learner.py:52-96.  probs [N,6] (eps-mixed, with grad), q_values [N,6], actions [N], masks [N,6].

    Counterfactual baseline on the masked, renormalised policy (no grad); the log-prob is the UNMASKED
    eps-mixed one and `advantage[N,1] * logp[N,1] * masks[N,6]` is averaged over all N*6 entries — so a sample
    is weighted by (#valid actions)/6, exactly as coded."""
    logp = torch.log(probs)
    with torch.no_grad():
        pm = probs * masks
        s = pm.sum(-1, keepdim=True).clamp_min(1e-5)
        pn = (pm / s).clamp_min(1e-5)
    a = actions.long().clamp_min(0)[:, None]
    q_chosen = q_values.gather(1, a)
    baseline = (pn * q_values * masks).sum(-1, keepdim=True)
    advantage = (q_chosen - baseline).detach()
    loss = -(advantage * logp.gather(1, a) * masks).mean()
    return loss, advantage


def coma_critic_loss(q_values, actions, td_targets):
    """critic/learner.py:76-94: mean squared TD error of the taken action."""
    q_chosen = q_values.gather(1, actions.long().clamp_min(0)[:, None]).squeeze(1)
    return torch.square(q_chosen - td_targets.detach()).mean()


class FlatGradAllReduce:
    """The only collective of the path: the mean of a network's gradients over the ranks, once
    per optimizer step, over NCCL.

    The gradients of all parameters live in ONE flat float32 buffer (every ``p.grad`` is a view into it, so nothing is
    copied in or out) that is cut into a few contiguous buckets.  Backward fills the buckets from the last layer to
    the first; a post-accumulate hook launches a bucket's asynchronous all-reduce as soon as its last gradient has
    arrived, so the transfer of the late layers overlaps the backward pass of the early ones.  ``finish()`` (=
    ``__call__``) waits for the outstanding work and scales by 1 / world.  Parameters that never receive a gradient
    (the reference's unused ``fc2``) are found during the first backward, which is reduced in one piece.
    Use ``zero_grad(set_to_none=False)`` with it: the views must survive."""

    def __init__(self, module, bucket_bytes=4 << 20):
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            # a view of the flat buffer with the parameter's own (dense) strides: channels_last conv weights keep the
            # layout autograd expects ("gradient layout contract"), the bytes stay in one contiguous range
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)
            off += p.numel()
        # buckets = contiguous ranges of the flat buffer, at least bucket_bytes each, built from the LAST parameter
        # backwards (the order in which backward produces gradients)
        self.buckets = []  # (lo, hi, [param indices])
        hi, members = n, []
        for i in range(len(self.params) - 1, -1, -1):
            members.append(i)
            if (hi - self.offsets[i]) * 4 >= bucket_bytes or i == 0:
                self.buckets.append((self.offsets[i], hi, members))
                hi, members = self.offsets[i], []
        self.bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self.bucket_of[i] = b
        self.used = None          # indices of the parameters that receive gradients (known after the first backward)
        self._seen = set()
        self._pending = []
        self._launched = set()
        self.bytes_reduced = 0    # bookkeeping for the benchmarks
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    @staticmethod
    def _world():
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _make_hook(self, i):
        def hook(param):
            self._seen.add(i)
            if self.used is None or self._world() == 1:
                return
            b = self.bucket_of[i]
            if b in self._launched:
                return
            if all((j in self._seen) or (j not in self.used) for j in self.buckets[b][2]):
                self._launch(b)
        return hook

    def _launch(self, b):
        import torch.distributed as dist

        lo, hi, _ = self.buckets[b]
        self._launched.add(b)
        self._pending.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        self.bytes_reduced += 4 * (hi - lo)

    def finish(self):
        world = self._world()
        if self.used is None:
            self.used = set(self._seen)
        if world > 1:
            for b in range(len(self.buckets)):  # whatever backward did not trigger (first call; unused layers)
                if b not in self._launched:
                    self._launch(b)
            for w in self._pending:
                w.wait()
            self.flat.div_(world)
        self._pending, self._launched, self._seen = [], set(), set()

    __call__ = finish


class COMATrainer:
    """``compute_dtype``: float32 like the reference (default; TF32 is NOT enabled by this class) or
    ``torch.bfloat16`` (autocast of the CNN forward / backward, opt-in, faster on tensor cores).

    Epsilon schedule: the reference anneals eps linearly in the EPISODE index (actor/network.py:53-58,
    ``eps_anneal_phase`` = 10 000 episodes).  One rollout here collects ``B * world`` episodes, so the counter advances
    by that many per rollout (``episodes_per_rollout``; pass 1 to anneal per rollout instead)."""

    def __init__(self, env, params, minibatch=8192, data_passes=None, compute_dtype=torch.float32,
                 reference_frozen_target=True, seed=0, episodes_per_rollout=None):
        self.env = env
        self.params = params
        net = params["networks"]
        self.gamma, self.lam = net["gamma"], net["lambda"]
        self.data_passes = net["data_passes"] if data_passes is None else data_passes
        self.copy_rate = net["copy_rate"]
        self.minibatch = minibatch
        self.compute_dtype = compute_dtype
        self.frozen_target = reference_frozen_target
        mis = params["experiment"]["missions"]
        self.eps_cfg = (mis["eps_max"], mis["eps_min"], mis["eps_anneal_phase"])
        dev = env.device
        torch.manual_seed(seed)  # identical initial weights on every rank
        self.actor = ActorNet().to(dev).to(memory_format=torch.channels_last)
        self.critic = CriticNet().to(dev).to(memory_format=torch.channels_last)
        self.target_critic = CriticNet().to(dev).to(memory_format=torch.channels_last)
        self.target_critic.load_state_dict(self.critic.state_dict())
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=net["actor"]["learning_rate"])
        self.opt_critic = torch.optim.Adam(self.critic.parameters(), lr=net["critic"]["learning_rate"])
        self.sync_actor = FlatGradAllReduce(self.actor)
        self.sync_critic = FlatGradAllReduce(self.critic)
        self.updates = 0
        self.episodes_done = 0
        world = FlatGradAllReduce._world()
        self.episodes_per_rollout = env.B * world if episodes_per_rollout is None else int(episodes_per_rollout)
        self._last_eps_episode = 0
        B, A, T, P = env.B, env.A, env.T, env.tables.px
        self.buf_obs = torch.empty((T, B, A, P, P, 7), dtype=torch.float32, device=dev)
        self.buf_state = torch.empty((T, B, A, P, P, 12), dtype=torch.float32, device=dev)
        self.buf_act = torch.empty((T, B, A), dtype=torch.int32, device=dev)
        self.buf_mask = torch.empty((T, B, A), dtype=torch.uint8, device=dev)
        self.buf_rew = torch.empty((T, B), dtype=torch.float32, device=dev)
        self.buf_abs = torch.empty((T, B), dtype=torch.float32, device=dev)  # absolute rewards (logged only)

    def _autocast(self):
        return torch.autocast("cuda", dtype=self.compute_dtype, enabled=self.compute_dtype != torch.float32)

    @torch.no_grad()
    def rollout(self, episodes=None, greedy=False):
        """One episode in every env (missions/episode_generator.py:38-88, batched).  Returns mean return."""
        env = self.env
        env.reset(episodes)
        eps = epsilon(self.episodes_done, *self.eps_cfg)
        self._last_eps_episode = self.episodes_done
        B, A = env.B, env.A
        for t in range(env.T):
            rel, ab = env.observe()
            env.features_actor(out=self.buf_obs[t])
            with self._autocast():
                probs = self.actor(self.buf_obs[t].flatten(0, 1), eps)
            env.act(probs=probs.view(B, A, N_ACTIONS), greedy=greedy)
            env.features_critic(self.buf_obs[t], out=self.buf_state[t])
            self.buf_act[t].copy_(env.actions)
            self.buf_mask[t].copy_(env.masks)
            self.buf_rew[t].copy_(rel)
            self.buf_abs[t].copy_(ab)
        self.episodes_done += self.episodes_per_rollout
        return self.buf_rew.sum(0).mean()

    def _masks6(self, m_u8):
        return ((m_u8[..., None].int() >> torch.arange(N_ACTIONS, device=m_u8.device)) & 1).float()

    def learn_minibatch(self, state, obs, act, masks, td, eps):
        """One critic step then one actor step on a mini-batch (coma_mission.py:93-98): critic regression on the TD
        targets (critic/learner.py:76-99), Q re-evaluated after the step without grad (:101-105), counterfactual
        policy gradient (actor/learner.py:52-101).  The gradient all-reduce of each network overlaps its backward.
        Returns (critic_loss, actor_loss, advantage, q_after)."""
        with self._autocast():
            q = self.critic(state)
        loss_c = coma_critic_loss(q, act, td)
        self.opt_critic.zero_grad(set_to_none=False)
        loss_c.backward()
        self.sync_critic.finish()
        self.opt_critic.step()
        with torch.no_grad(), self._autocast():
            q_new = self.critic(state)
        with self._autocast():
            probs = self.actor(obs, eps)
        loss_a, adv = coma_actor_loss(probs, q_new, act, masks)
        self.opt_actor.zero_grad(set_to_none=False)
        loss_a.backward()
        self.sync_actor.finish()
        self.opt_actor.step()
        return loss_c.detach(), loss_a.detach(), adv, q_new

    def update(self):
        """TD(lambda) targets + `data_passes` passes of critic / actor mini-batch steps (coma_mission.py:89-98)."""
        env = self.env
        T, B, A = env.T, env.B, env.A
        eps = epsilon(self._last_eps_episode, *self.eps_cfg)  # the eps of the last collected episode (coma_mission.py:78)
        obs = self.buf_obs.flatten(0, 2)
        state = self.buf_state.flatten(0, 2)
        act = self.buf_act.flatten()
        masks = self._masks6(self.buf_mask.flatten())
        with torch.no_grad(), self._autocast():
            tgt_net = self.target_critic
            q_all = torch.cat([tgt_net(state[i:i + 32768]) for i in range(0, state.shape[0], 32768)])
        q_taken = q_all.gather(1, act.long().clamp_min(0)[:, None]).view(T, B, A)
        rewards = self.buf_rew[:, :, None].expand(T, B, A)  # team reward shared by all agents (coma_wrapper.py:166-168)
        td = td_lambda_targets(rewards.permute(1, 2, 0), q_taken.permute(1, 2, 0), self.gamma, self.lam)  # [B, A, T]
        td = td.permute(2, 0, 1).reshape(-1)
        n = obs.shape[0]
        stats = {}
        for _ in range(self.data_passes):
            perm = torch.randperm(n, device=obs.device)
            for i in range(0, n, self.minibatch):
                idx = perm[i:i + self.minibatch]
                loss_c, loss_a, adv, _ = self.learn_minibatch(state[idx], obs[idx], act[idx], masks[idx], td[idx], eps)
                stats = {"critic_loss": loss_c, "actor_loss": loss_a, "adv_mean": adv.mean()}
        self.updates += 1
        if not self.frozen_target and self.updates % self.copy_rate == 0:
            self.target_critic.load_state_dict(self.critic.state_dict())
        return stats

    def optimizer_steps_per_update(self):
        n = self.env.T * self.env.B * self.env.A
        return 2 * self.data_passes * ((n + self.minibatch - 1) // self.minibatch)

    def flops_per_update(self):
        """Forward MACs: 20.1 M per actor observation, 21.7 M per critic state."""
        n = self.env.T * self.env.B * self.env.A
        fwd_actor, fwd_critic = 2 * 20.1e6, 2 * 21.7e6
        rollout = n * fwd_actor
        learn = n * fwd_critic + self.data_passes * n * (3 * fwd_critic + fwd_critic + 3 * fwd_actor)
        return rollout + learn

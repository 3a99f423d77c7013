"""COMA learner arithmetic: vectorised TD(lambda), critic / actor losses and the
network architectures vs values produced by the reference's own BatchMemory / CriticLearner / ActorLearner
(tests/golden/coma_kats.npz, written by oracle/make_golden_coma.py).  CPU, float32."""
import os

import numpy as np
import pytest
import torch

from ipp_marl_b200 import coma

K = np.load(os.path.join(os.path.dirname(__file__), "golden", "coma_kats.npz"))


def _critic(seed):
    torch.manual_seed(seed)
    return coma.CriticNet()


def test_architectures_and_seeded_init_match_reference():
    seed = int(K["seed"])
    critic = _critic(seed)
    with torch.no_grad():
        q0 = critic(torch.from_numpy(K["td_states"][0, 0])[None])[0]
    assert np.allclose(q0.numpy(), K["critic_q0"], rtol=1e-5, atol=1e-6)
    torch.manual_seed(seed + 1)
    actor = coma.ActorNet()
    with torch.no_grad():
        p = actor(torch.from_numpy(K["lb_obs"]).float(), float(K["lb_eps"]))
    assert np.allclose(p.numpy(), K["actor_probs0"], rtol=1e-5, atol=1e-6)
    assert sum(x.numel() for x in actor.parameters()) == 2275846
    assert sum(x.numel() for x in critic.parameters()) == 2307846


def test_td_lambda_matches_reference_build_td_targets():
    seed = int(K["seed"])
    critic = _critic(seed)
    states = torch.from_numpy(K["td_states"])            # [T, A, 11, 11, 12]
    actions = torch.from_numpy(K["td_actions"])          # [T, A]
    T, A = actions.shape
    with torch.no_grad():
        q = critic(states.flatten(0, 1)).view(T, A, 6)
    q_taken = q.gather(2, actions[..., None]).squeeze(-1)             # [T, A]
    rewards = torch.from_numpy(K["td_rewards"])[:, None].expand(T, A)
    td = coma.td_lambda_targets(rewards.T.contiguous(), q_taken.T.contiguous(), 0.99, 0.8).T
    assert np.allclose(td.numpy(), K["td_targets"], rtol=2e-5, atol=2e-6), np.abs(td.numpy() - K["td_targets"]).max()


def test_critic_and_actor_losses_match_reference_learners():
    seed = int(K["seed"])
    torch.manual_seed(seed + 1)
    actor = coma.ActorNet()
    torch.manual_seed(seed + 2)
    critic = coma.CriticNet()
    st = torch.from_numpy(K["lb_state"])
    act = torch.from_numpy(K["lb_act"])[:, 0]
    td = torch.from_numpy(K["lb_td"])[:, 0]
    q = critic(st)
    assert np.allclose(q.detach().numpy(), K["critic_q_before"], rtol=1e-5, atol=1e-6)
    loss_c = coma.coma_critic_loss(q, act, td)
    assert abs(float(loss_c) - float(K["critic_loss"])) <= 1e-6 + 1e-5 * abs(float(K["critic_loss"]))
    # one Adam step with the reference's learning rate reproduces its post-step Q values
    opt = torch.optim.Adam(critic.parameters(), lr=1e-4)
    opt.zero_grad()
    loss_c.backward()
    opt.step()
    with torch.no_grad():
        q_after = critic(st)
    assert np.allclose(q_after.numpy(), K["critic_q_after"], rtol=1e-4, atol=1e-5)
    probs = actor(torch.from_numpy(K["lb_obs"]).float(), float(K["lb_eps"]))
    loss_a, adv = coma.coma_actor_loss(probs, torch.from_numpy(K["critic_q_after"]), act,
                                       torch.from_numpy(K["lb_masks"]))
    assert abs(float(loss_a) - float(K["actor_loss"])) <= 1e-6 + 1e-5 * abs(float(K["actor_loss"]))
    assert abs(float(adv.mean()) - float(K["actor_adv_mean"])) <= 1e-6


def test_flat_grad_allreduce_is_identity_without_process_group():
    """Gradients live as views in one flat buffer; without a process group finish() changes nothing, and the views
    give the same gradients as an untouched copy of the network."""
    torch.manual_seed(5)
    net, twin = coma.CriticNet(), coma.CriticNet()
    twin.load_state_dict(net.state_dict())
    sync = coma.FlatGradAllReduce(net, bucket_bytes=1 << 20)
    assert [hi - lo for lo, hi, _ in sync.buckets] and sync.buckets[0][1] == sync.flat.numel() and sync.buckets[-1][0] == 0
    assert sum(hi - lo for lo, hi, _ in sync.buckets) == 2307846
    x = torch.rand(2, 11, 11, 12)
    for _ in range(2):  # second pass: the views survive zero_grad(set_to_none=False) and accumulate in place
        net.zero_grad(set_to_none=False)
        twin.zero_grad(set_to_none=True)
        net(x).sum().backward()
        twin(x).sum().backward()
        sync.finish()
        for (n, p), q in zip(net.named_parameters(), twin.parameters()):
            assert p.grad.data_ptr() >= sync.flat.data_ptr()
            if q.grad is None:  # fc2: constructed but unused
                assert n.startswith("fc2") and float(p.grad.abs().sum()) == 0.0
            else:
                assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-8), n
    assert sync.used is not None and len(sync.used) == 10  # 12 parameter tensors, fc2.weight / fc2.bias unused


def test_epsilon_schedule():
    assert coma.epsilon(0) == 0.5 and coma.epsilon(20000) == 0.02
    assert abs(coma.epsilon(5000) - (0.5 - 0.5 * 0.48)) < 1e-12


@pytest.mark.gpu
def test_learn_minibatch_fp32_matches_reference_learners_on_gpu():
    """COMATrainer.learn_minibatch (the arithmetic of one critic + one actor optimizer step, as update() runs it) on
    the GPU in float32 against the values of the reference's CriticLearner / ActorLearner (coma_kats.npz): critic
    loss, post-step Q values, actor loss and mean advantage."""
    import json

    from ipp_marl_b200 import BatchedIPPEnv

    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        params = json.load(open(os.path.join(root, "tests", "golden", "kats.json")))["synthetic50"]["params"]
        env = BatchedIPPEnv(params, 8, device="cuda:0")
        tr = coma.COMATrainer(env, params)
        assert tr.compute_dtype == torch.float32
        seed = int(K["seed"])
        torch.manual_seed(seed + 1)
        tr.actor.load_state_dict(coma.ActorNet().state_dict())
        torch.manual_seed(seed + 2)
        tr.critic.load_state_dict(coma.CriticNet().state_dict())
        dev = env.device
        st = torch.from_numpy(K["lb_state"]).to(dev)
        obs = torch.from_numpy(K["lb_obs"]).float().to(dev)
        act = torch.from_numpy(K["lb_act"])[:, 0].to(dev)
        td = torch.from_numpy(K["lb_td"])[:, 0].to(dev)
        masks = torch.from_numpy(K["lb_masks"]).to(dev)
        with torch.no_grad():
            assert np.allclose(tr.critic(st).cpu().numpy(), K["critic_q_before"], rtol=1e-4, atol=1e-5)
        loss_c, loss_a, adv, q_after = tr.learn_minibatch(st, obs, act, masks, td, float(K["lb_eps"]))
        assert abs(float(loss_c) - float(K["critic_loss"])) <= 1e-6 + 1e-4 * abs(float(K["critic_loss"]))
        assert np.allclose(q_after.cpu().numpy(), K["critic_q_after"], rtol=2e-4, atol=2e-5)
        assert abs(float(loss_a) - float(K["actor_loss"])) <= 2e-6 + 2e-4 * abs(float(K["actor_loss"]))
        assert abs(float(adv.mean()) - float(K["actor_adv_mean"])) <= 2e-5
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


@pytest.mark.gpu
def test_trainer_runs_and_learns_signal():
    """Two iterations of rollout + update on the GPU: finite losses, rewards identical to a policy-free
    replay of the same actions (the env under the trainer is the parity-tested env)."""
    import json

    from ipp_marl_b200 import BatchedIPPEnv

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    params = json.load(open(os.path.join(root, "tests", "golden", "kats.json")))["synthetic50"]["params"]
    env = BatchedIPPEnv(params, 64, device="cuda:0")
    tr = coma.COMATrainer(env, params, minibatch=1024, data_passes=1)
    r0 = tr.rollout(episodes=torch.arange(64) + 1)
    acts = tr.buf_act.clone()
    rew = tr.buf_rew.clone()
    stats = tr.update()
    assert all(torch.isfinite(v).all() for v in stats.values())
    assert torch.isfinite(r0)
    # replay: inject the recorded actions into a fresh env (fused step): same rewards
    env2 = BatchedIPPEnv(params, 64, device="cuda:0")
    env2.reset(torch.arange(64) + 1)
    for t in range(env2.T):
        rel, _, _ = env2.step(actions=acts[t])
        assert torch.allclose(rel, rew[t], rtol=1e-5, atol=1e-5), t
    r1 = tr.rollout(episodes=torch.arange(64) + 100)
    assert torch.isfinite(r1)


def test_best_model_rule_and_scalar_log(tmp_path):
    """missions/coma_mission.py:425-435: running mean over ALL returns so far, compared once `patience` exist."""
    import json

    from ipp_marl_b200 import mission

    best = mission.BestModel(patience=3, path=None)
    took = [best.offer(r, None) for r in (1.0, 5.0, 0.0, 6.0, -10.0, 20.0)]
    # running means: 1, 3, 2, 3, 0.4, 3.67 -> first comparison at the 3rd return
    assert took == [False, False, True, True, False, True]
    assert abs(best.best - 22.0 / 6.0) < 1e-12
    log = mission.ScalarLog(str(tmp_path))
    log.add_scalar("trainReturn/Episode/mean", 1.5, 7)
    log.close()
    rows = [json.loads(x) for x in open(tmp_path / "scalars.jsonl")]
    assert rows == [{"tag": "trainReturn/Episode/mean", "value": 1.5, "step": 7}]


@pytest.mark.gpu
def test_mission_loop_runs_logs_and_checkpoints(tmp_path):
    """The batched training mission: updates, the reference's scalar tags, greedy evaluation with the entropy / F1
    metrics, best-model checkpoint that loads back into an ActorNet."""
    import json

    from ipp_marl_b200 import BatchedIPPEnv, mission
    from tests.helpers import load_kats

    params = load_kats()["synthetic50"]["params"]
    params["experiment"]["missions"]["patience"] = 2
    env = BatchedIPPEnv(params, 16, device="cuda:0")
    tr = coma.COMATrainer(env, params, minibatch=512, data_passes=1, compute_dtype=torch.float32)
    m = mission.COMAMission(tr, str(tmp_path), eval_every=2)
    best = m.execute(3)
    assert np.isfinite(best) and m.training_step_idx == 3 and m.environment_step_idx == 3 * 15 * 4 * 16
    tags = {json.loads(x)["tag"] for x in open(tmp_path / "scalars.jsonl")}
    for t in ("trainReturn/Episode/mean", "trainRewards/Episode/std", "trainReturn/Relative(used)/Episode/max",
              "evalReturn/Episode/mean", "evalMetrics/entropy_final", "evalMetrics/f1_final", "Training/critic_loss",
              "trainActions/0", "trainAltitudes/15"):
        assert t in tags, t
    actor = coma.ActorNet()
    actor.load_state_dict(mission.load_actor_state(tmp_path / "best_model.pth"))
    ent = [json.loads(x) for x in open(tmp_path / "scalars.jsonl") if "entropy_final" in x][0]["value"]
    assert 0.0 < ent <= 1.0


def test_checkpoint_loader_reads_state_dicts_and_pickled_modules(tmp_path):
    """mission.load_actor_state: our own format (state dict) and the reference's (whole pickled module,
    coma_mission.py:435)."""
    from ipp_marl_b200 import mission

    torch.manual_seed(1)
    actor = coma.ActorNet()
    mission.save_actor(actor, tmp_path / "a.pth")
    torch.save(actor, tmp_path / "b.pth")  # what the reference does with its ActorNetwork
    for name in ("a.pth", "b.pth"):
        sd = mission.load_actor_state(tmp_path / name)
        other = coma.ActorNet()
        other.load_state_dict(sd)
        assert all(torch.equal(x, y) for x, y in zip(actor.state_dict().values(), other.state_dict().values()))


def _allreduce_worker(tmp_path, backend, port):
    """world_size-2 run of FlatGradAllReduce: both ranks end with the mean of the two ranks' gradients, parameter by
    parameter, and identical weights after the optimizer step; step 0 reduces in one piece (and discovers the unused
    fc2), step 1 launches the bucket all-reduces from the backward hooks."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(r"""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, %r)
from ipp_marl_b200 import coma
backend = %r
if backend == "nccl":
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
else:
    dev = torch.device("cpu")
    dist.init_process_group("gloo")
torch.set_default_device(dev)
torch.backends.cudnn.allow_tf32 = False
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)                      # identical initial weights on every rank (COMATrainer does the same)
net = coma.CriticNet()
sync = coma.FlatGradAllReduce(net)
opt = torch.optim.SGD(net.parameters(), lr=0.1)
torch.manual_seed(100 + rank)             # different data per rank
twin = coma.CriticNet()                   # plain copy: the rank-local gradients
ok = True
launched_by_hooks = []
for it in range(2):                       # step 0: one-piece reduce (discovers the unused fc2); step 1: bucket hooks
    twin.load_state_dict(net.state_dict())
    x = torch.rand(4, 11, 11, 12)
    opt.zero_grad(set_to_none=False)
    twin.zero_grad(set_to_none=True)
    net(x).square().mean().backward()
    launched_by_hooks.append(len(sync._launched))
    twin(x).square().mean().backward()
    local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in twin.parameters()]
    sync()
    for p, g in zip(net.parameters(), local):
        both = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(both, g)
        ok = ok and torch.allclose(p.grad, sum(both) / world, rtol=1e-5, atol=1e-7)
    opt.step()
ok = ok and launched_by_hooks[0] == 0 and launched_by_hooks[1] == len(sync.buckets)
w = torch.cat([p.detach().flatten() for p in net.parameters()])
ws = [torch.zeros_like(w) for _ in range(world)]
dist.all_gather(ws, w)
if rank == 0:
    print(json.dumps({"ok": bool(ok), "same_weights": bool(torch.equal(ws[0], ws[1])), "n": int(sync.flat.numel())}))
dist.destroy_process_group()
""" % (root, backend))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out == {"ok": True, "same_weights": True, "n": 2307846}


def test_flat_grad_allreduce_gloo(tmp_path):
    """The only collective of the path: the mean of a network's gradients over the ranks once
    per optimizer step.  world_size-2 gloo run on CPU."""
    _allreduce_worker(tmp_path, "gloo", 29641)


@pytest.mark.gpu
def test_flat_grad_allreduce_nccl_two_gpus(tmp_path):
    """The same over NCCL on two GPUs of one box (skipped on a single-GPU box; `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _allreduce_worker(tmp_path, "nccl", 29643)

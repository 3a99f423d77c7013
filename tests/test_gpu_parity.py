"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI via
ipp_marl_b200.BatchedIPPEnv; the oracle is only the checker.

Bars: belief maps BIT-EXACT vs the kernel arithmetic model (oracle/kernel_model.py, itself gated
against the reference arithmetic on CPU), and allclose(rtol=1e-5, atol=1e-5) vs the reference's
own outputs in tests/golden (SURVEY.md section 8d gate).  Integer outputs (positions, actions,
masks, comm matrix, ground truth) are bit-exact.
"""
import numpy as np
import pytest

from tests.helpers import gate_stats, golden_episodes, load_episode, load_kats

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL = 1e-5


def _env(params, episodes, variant=None):
    import torch
    from ipp_marl_b200 import BatchedIPPEnv

    env = BatchedIPPEnv(params, len(episodes), device="cuda:0")
    if variant is not None:
        env.set_step_variant(variant)
    env.reset(episodes)
    torch.cuda.synchronize()
    return env


def _bits(u8, A, width):
    return ((u8[..., None].astype(np.int64) >> np.arange(width)) & 1).astype(np.uint8)


CASES = [p for p in golden_episodes() if "g493" not in p]


@pytest.mark.parametrize("path", CASES, ids=[p.split("episode_")[1] for p in CASES])
def test_fused_step_vs_reference_golden(path):
    """ipp_step over a whole episode vs the reference's own outputs."""
    g = load_episode(path)
    env = _env(g["params"], [g["episode"]])
    A = env.A
    assert np.array_equal(env.ground_truth[0].cpu().numpy(), g["gt"])
    keep = {int(t): i for i, t in enumerate(g["map_steps"])}
    prior_half = g["params"]["mapping"]["prior"] == 0.5
    worst = {"max_abs": 0.0, "fail_pure_rtol": 0}
    for t in range(len(g["reward_rel"])):
        assert np.array_equal(env.pos[0].cpu().numpy(), g["pos"][t])
        rel, ab, done = env.step()
        assert done == (t == g["params"]["experiment"]["constraints"]["budget"])
        assert np.array_equal(_bits(env.comm[0].cpu().numpy(), A, A), g["comm"][t])
        assert np.array_equal(_bits(env.masks[0].cpu().numpy(), A, 6), g["mask"][t].astype(np.uint8))
        assert np.array_equal(env.actions[0].cpu().numpy(), g["action"][t])
        assert np.array_equal(env.pos[0].cpu().numpy(), g["pos_next"][t])
        if t in keep:
            i = keep[t]
            for ref, got in ((g["global"][i], env.global_map[0]), (g["local_after_move"][i], env.local_maps[0])):
                s = gate_stats(ref, got.cpu().numpy(), RTOL, ATOL)
                assert s["fail_gate"] == 0, (t, s)
                worst["max_abs"] = max(worst["max_abs"], s["max_abs"])
                worst["fail_pure_rtol"] += s["fail_pure_rtol"]
        if prior_half:
            assert abs(float(rel[0]) - g["reward_rel"][t]) <= ATOL + RTOL * abs(g["reward_rel"][t]), t
            assert abs(float(ab[0]) - g["reward_abs"][t]) <= ATOL + RTOL * abs(g["reward_abs"][t]), t
        else:
            assert abs(float(rel[0]) - g["reward_rel"][t]) <= 2e-2
    print("worst", worst)


@pytest.mark.parametrize("variant", ["direct", "tma"])
def test_default_g493_vs_reference_golden(variant):
    """Reference default config (493x493 belief cells): rewards, moves and the final global map."""
    g = load_episode([p for p in golden_episodes() if "g493" in p][0])
    env = _env(g["params"], [g["episode"]], variant)
    assert np.array_equal(env.ground_truth[0].cpu().numpy(), g["gt"])
    for t in range(len(g["reward_rel"])):
        rel, ab, _ = env.step()
        assert np.array_equal(env.pos[0].cpu().numpy(), g["pos_next"][t])
        assert abs(float(rel[0]) - g["reward_rel"][t]) <= ATOL + RTOL * abs(g["reward_rel"][t]), t
        assert abs(float(ab[0]) - g["reward_abs"][t]) <= ATOL + RTOL * abs(g["reward_abs"][t]), t
        assert abs(float(env.global_map[0].double().sum()) - g["global_sum"][t]) <= 1e-6 * g["global_sum"][t]
        assert abs(float(env.local_maps[0].double().sum()) - g["local_after_move_sum"][t]) <= 1e-6 * g["local_after_move_sum"][t]
    s = gate_stats(g["global_final_f32"], env.global_map[0].cpu().numpy(), RTOL, ATOL)
    assert s["fail_gate"] == 0, s


@pytest.mark.parametrize("variant", ["direct", "tma"])
@pytest.mark.parametrize("tag,n_agents,B", [("synthetic50", 4, 64), ("synthetic50", 2, 64), ("synthetic50", 1, 16),
                                            ("synthetic50", 3, 700), ("synthetic100", 8, 8)])
def test_fused_step_bit_exact_vs_kernel_model(tag, n_agents, B, variant):
    from oracle import kernel_model as km

    params = load_kats()[tag]["params"]
    params["experiment"]["missions"]["n_agents"] = n_agents
    episodes = np.arange(1, B + 1)
    env = _env(params, episodes, variant)
    assert env.step_variant == variant
    model = km.KernelModelEnv(params, episodes)
    assert np.array_equal(env.ground_truth.cpu().numpy(), model.gt)
    assert np.array_equal(env.positions[0].cpu().numpy(), model.pos)
    assert np.array_equal(env.local_maps.cpu().numpy(), model.local)
    for t in range(env.T):
        rel, ab, _ = env.step()
        out = model.step()
        assert np.array_equal(_bits(env.comm.cpu().numpy(), env.A, env.A).astype(bool), out["comm"])
        assert np.array_equal(_bits(env.masks.cpu().numpy(), env.A, 6), out["mask"])
        assert np.array_equal(env.actions.cpu().numpy(), out["action"])
        assert np.array_equal(env.pos.cpu().numpy(), model.pos)
        # resident state (float32 odds) and its probability export, both bit for bit
        assert np.array_equal(env.global_odds.cpu().numpy(), model.glob_o), t
        assert np.array_equal(env.local_odds.cpu().numpy(), model.local_o), t
        assert np.array_equal(env.global_map.cpu().numpy(), model.glob), t
        assert np.array_equal(env.local_maps.cpu().numpy(), model.local), t
        assert np.allclose(rel.cpu().numpy(), out["reward_rel"], rtol=RTOL, atol=ATOL)
        assert np.allclose(ab.cpu().numpy(), out["reward_abs"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(env.stuck.cpu().numpy().astype(bool), model.flag_stuck) or t < env.T


@pytest.mark.parametrize("variant", ["direct", "tma"])
def test_split_observe_act_bit_exact_vs_kernel_model(variant):
    from oracle import kernel_model as km

    params = load_kats()["synthetic50"]["params"]
    params["experiment"]["uav"]["communication_range"] = 15
    params["experiment"]["uav"]["failure_rate"] = 0.25
    episodes = np.arange(3, 35)
    env = _env(params, episodes, variant)
    model = km.KernelModelEnv(params, episodes)
    rng = np.random.RandomState(0)
    for t in range(env.T):
        rel, ab = env.observe()
        mo = model.observe()
        assert np.array_equal(env.global_map.cpu().numpy(), model.glob)
        assert np.array_equal(env.local_maps.cpu().numpy(), model.local)
        assert np.allclose(rel.cpu().numpy(), mo["reward_rel"], rtol=RTOL, atol=ATOL)
        if t % 2 == 0:
            acts, _ = env.act()  # uniform policy
            ma = model.act()
        else:  # injected actions: replay what the uniform policy of the model would do, perturbed to "stay"
            probe = km.KernelModelEnv.__new__(km.KernelModelEnv)
            probe.__dict__.update(model.__dict__)
            probe.flag_stuck = model.flag_stuck.copy()
            _, _, inj = probe._choose_and_move(None)
            inj = np.where(rng.rand(*inj.shape) < 0.2, -1, inj)
            acts, _ = env.act(actions=inj)
            ma = model.act(actions=inj)
        assert np.array_equal(env.actions.cpu().numpy(), ma["action"])
        assert np.array_equal(env.pos.cpu().numpy(), model.pos)
        assert np.array_equal(env.local_odds.cpu().numpy(), model.local_o)
        assert np.array_equal(env.local_maps.cpu().numpy(), model.local)


def test_reset_matches_numpy_seeding():
    """MT19937-compatible start positions / ground truth for many episodes (state_space.py:28-51,
    ground_truths.py:42-56), including large episode numbers."""
    from oracle import numpy_oracle as no

    params = load_kats()["synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = 8
    geo = no.Geometry(params)
    episodes = np.concatenate([np.arange(1, 200), np.array([1000, 65536, 1234567, 2**31 - 1, 2**31 + 5])])
    env = _env(params, episodes)
    pos = env.positions[0].cpu().numpy()
    gt = env.ground_truth.cpu().numpy()
    for b, ep in enumerate(episodes):
        for a in range(8):
            if geo.seed * int(ep) * a < 2**32:  # numpy refuses larger seeds; the kernel wraps
                assert pos[b, a].tolist() == no.start_position(geo, a, int(ep)).tolist(), (ep, a)
        if ep < 2**32:
            assert np.array_equal(gt[b], no.ground_truth(geo, int(ep)).astype(np.uint8)), ep


def test_probs_policy_greedy_and_sampled():
    import torch

    params = load_kats()["synthetic50"]["params"]
    B = 256
    env = _env(params, np.arange(1, B + 1))
    probs = torch.rand((B, env.A, 6), device="cuda:0")
    env.step(probs=probs, greedy=True)
    m = _bits(env.masks.cpu().numpy(), env.A, 6).astype(bool)
    pr = probs.cpu().numpy()
    expect = np.where(m, pr, -1.0).argmax(-1)
    ok = m.any(-1)
    assert np.array_equal(env.actions.cpu().numpy()[ok], expect[ok])
    # sampled actions are always valid, and cover more than one action overall
    env.reset(np.arange(1, B + 1))
    env.step(probs=probs, greedy=False)
    a = env.actions.cpu().numpy()
    m = _bits(env.masks.cpu().numpy(), env.A, 6).astype(bool)
    assert np.all(np.take_along_axis(m, np.maximum(a, 0)[..., None], -1)[..., 0] | (a < 0))
    assert len(np.unique(a)) >= 4


def test_partition_invariance_and_determinism():
    """Shard equivalence: B envs on one handle == the same episodes split over two."""
    import torch

    params = load_kats()["synthetic50"]["params"]
    eps = np.arange(1, 257)
    full = _env(params, eps, "tma")
    lo = _env(params, eps[:100], "tma")
    hi = _env(params, eps[100:], "direct")
    again = _env(params, eps, "direct")
    for t in range(full.T):
        for e in (full, lo, hi, again):
            e.step()
    torch.cuda.synchronize()
    for name in ("_local", "_glob", "reward_rel", "reward_abs", "actions"):
        a = getattr(full, name)
        assert torch.equal(a, torch.cat([getattr(lo, name), getattr(hi, name)])), name
        assert torch.equal(a, getattr(again, name)), name
    assert torch.equal(full.positions, torch.cat([lo.positions, hi.positions], dim=1))


def test_full_size_properties():
    """BASELINE config sizes (8192 envs x 4 UAVs): properties that do not need the oracle."""
    import torch

    params = load_kats()["synthetic50"]["params"]
    B = 8192
    env = _env(params, np.arange(1, B + 1))
    total_rel = torch.zeros(B, device="cuda:0", dtype=torch.float64)
    for t in range(env.T):
        rel, ab, _ = env.step()
        assert torch.isfinite(rel).all() and torch.isfinite(ab).all()
        total_rel += rel.double()
        lm, gm = env.local_maps, env.global_map
        assert float(lm.min()) > 0.0 and float(lm.max()) < 1.0
        assert float(gm.min()) > 0.0 and float(gm.max()) < 1.0
    # positions stay on the lattice and inside the environment (agent/agent.py:106-117)
    p = env.positions.cpu().numpy()
    assert p[..., :2].min() >= 0 and p[..., :2].max() <= 50 and set(np.unique(p[..., 2])) <= {5, 10, 15}
    assert (p % 5 == 0).all()
    # a random subsample of envs is bit-exact against the kernel model
    from oracle import kernel_model as km

    rng = np.random.default_rng(8192)
    pick = np.unique(np.concatenate([[0, 17, 4095, 8191], rng.choice(B, 64, replace=False)]))  # >= 64 random envs
    model = km.KernelModelEnv(params, pick + 1)
    for t in range(env.T):
        model.step()
    assert np.array_equal(env.global_map[pick].cpu().numpy(), model.glob)
    assert np.array_equal(env.local_maps[pick].cpu().numpy(), model.local)
    # information is gained on average: the mean episode return of the relative reward is positive
    assert float(total_rel.mean()) > 0.0


@pytest.mark.gpu
def test_step_host_matches_device_step():
    """ipp_step_host (host policy in pinned memory -> one C call) gives exactly what step(probs=...) gives."""
    import torch
    from ipp_marl_b200 import BatchedIPPEnv

    params = load_kats()["synthetic50"]["params"]
    B = 96
    a = BatchedIPPEnv(params, B, device="cuda:0")
    b = BatchedIPPEnv(params, B, device="cuda:0")
    a.reset()
    b.reset()
    g = torch.Generator().manual_seed(1)
    packed = b.host_results()  # one pinned block -> single device->host copy
    loose = (torch.empty((B,), dtype=torch.float32).pin_memory(), torch.empty((B,), dtype=torch.float32).pin_memory(),
             torch.empty((B, a.A), dtype=torch.int32).pin_memory())
    for t in range(a.T):
        rel_h, abs_h, act_h = packed if t % 3 else loose
        probs = torch.rand((B, a.A, 6), generator=g, dtype=torch.float32).pin_memory()
        rel, ab, _ = a.step(probs=probs.cuda())
        if t % 2 == 0:
            b.step_host(probs, None, rel_h, abs_h, act_h)
        else:  # injected actions from the host: replay what the other env just chose
            b.step_host(None, a.actions.cpu().contiguous(), rel_h, abs_h, act_h)
        torch.cuda.synchronize()
        assert torch.equal(rel.cpu(), rel_h) and torch.equal(ab.cpu(), abs_h)
        assert torch.equal(a.actions.cpu(), act_h)
        assert torch.equal(a.pos.cpu(), b.pos.cpu())
    assert torch.equal(a.local_odds, b.local_odds) and torch.equal(a.global_odds, b.global_odds)
    with pytest.raises(Exception):
        b.step_host(probs, None, rel_h, abs_h, act_h)  # episode finished


def test_run_steps_graph_equals_eager_steps():
    """ipp_run_steps (one CUDA-graph launch per episode, captured once and replayed) == the same ipp_reset / ipp_step
    calls one by one: bit-identical maps, positions, per-step rewards, actions and masks — on the first (capturing)
    call, on the replay of a second episode, and for a partial run from the middle of an episode."""
    import torch
    from ipp_marl_b200 import BatchedIPPEnv

    params = load_kats()["synthetic50"]["params"]
    params["experiment"]["missions"]["n_agents"] = 3
    B = 37
    a = BatchedIPPEnv(params, B, device="cuda:0")
    b = BatchedIPPEnv(params, B, device="cuda:0")
    for ep0 in (1, 500):
        eps = torch.arange(B) + ep0
        a.reset(eps)
        rel, ab, act, msk = [], [], [], []
        for t in range(a.T):
            r, q, _ = a.step()
            rel.append(r.clone()); ab.append(q.clone()); act.append(a.actions.clone()); msk.append(a.masks.clone())
        assert b.run_steps(reset=True, episodes=eps) is True
        torch.cuda.synchronize()
        assert torch.equal(a._local, b._local) and torch.equal(a._glob, b._glob) and torch.equal(a._flags, b._flags)
        assert torch.equal(a.positions, b.positions)
        assert torch.equal(torch.stack(rel), b.reward_hist[:, 0]) and torch.equal(torch.stack(ab), b.reward_hist[:, 1])
        assert torch.equal(torch.stack(act), b.action_hist) and torch.equal(torch.stack(msk), b.mask_hist)
    # partial: eager reset + 4 steps, then the rest of the episode as one launch
    eps = torch.arange(B) + 77
    a.reset(eps)
    b.reset(eps)
    for t in range(a.T):
        a.step()
    for t in range(4):
        b.step()
    assert b.run_steps() is True
    torch.cuda.synchronize()
    assert torch.equal(a._local, b._local) and torch.equal(a._glob, b._glob) and torch.equal(a.positions, b.positions)
    with pytest.raises(Exception):
        b.run_steps()  # episode finished

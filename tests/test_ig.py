"""Information-gain greedy planner + evaluation metrics.

CPU: oracle/numpy_ig.py reproduces the unmodified reference IG_baseline (fixtures tests/golden/ig_*.npz written by
oracle/make_golden.py) bit for bit: gains, utilities, argmax actions, entropy and F1 curves.
GPU: ipp_ig_plan / ipp_eval_metrics through BatchedIPPEnv vs the same fixtures.
"""
import numpy as np
import pytest

from tests.helpers import golden_episodes, load_episode

CASES = golden_episodes("ig_*.npz")
SMALL = [p for p in CASES if "g493" not in p]


def _id(p):
    return p.split("ig_")[1].replace(".npz", "")


def test_ig_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[_id(p) for p in CASES])
def test_numpy_ig_bit_exact(path):
    from oracle import numpy_ig as ni

    g = load_episode(path)
    if "g493" in path:  # reference-sized grid: the first planning steps only (CPU time)
        from oracle import numpy_oracle as no

        env = no.OracleEnv(g["params"], g["episode"])
        for t in range(2):
            env.observe()
            masks, gains, util, acts = ni.plan(env.geo, env.pos, env.local, bool(g["communication"]))
            assert np.array_equal(np.array(env.pos), g["pos"][t])
            assert np.array_equal(gains, g["gains"][t])
            assert np.array_equal(util, g["util"][t])
            assert np.array_equal(acts, g["action"][t])
            env.act(acts)
        return
    rec = ni.run_ig_episode(g["params"], g["episode"], communication=bool(g["communication"]))
    for key in ("pos", "gains", "util", "action", "entropy", "f1"):
        assert np.array_equal(rec[key], g[key]), key


def test_cell_utilities_sequential_semantics():
    """IG_baseline.py:297-322: agent 1 is discounted with agent 0's ALREADY discounted value; the last matching
    candidate wins (the products are not cumulative)."""
    from oracle import numpy_ig as ni

    p = np.array([5, 5, 5])
    q = np.array([10, 5, 5])
    pos = [[p, q], [p, 0], [q, p]]
    rel = [[0.6, 0.4], [1.0, 0], [0.5, 0.5]]
    out = ni.cell_utilities(pos, [list(r) for r in rel])
    a00 = 0.6 * (1 - 0.5)            # last match: agent 2's candidate 1 (value 0.5)
    a01 = 0.4 * (1 - 0.5)
    a10 = 1.0 * (1 - 0.5)            # agent 0 first (-> 1 * (1 - a00)), then overwritten by agent 2's match
    a20 = 0.5 * (1 - a01)
    a21 = 0.5 * (1 - a10)            # agent 0's value first, then agent 1's (already discounted) value
    assert np.allclose(out[0], [a00, a01]) and np.isclose(out[1][0], a10) and np.allclose(out[2], [a20, a21])


def _f1_bounds(global_map, gt, eps=2e-5):
    """F1 of class 1 with the cells within eps of the 0.5 threshold counted either way (cancelling evidence puts
    cells at 0.5 +- rounding; which side they fall on is rounding noise in the reference too)."""
    sure1 = global_map > 0.5 + eps
    edge = np.abs(global_map - 0.5) <= eps
    unseen = global_map == 0.5
    vals = []
    for fill in (False, True):
        pred = sure1 | (edge & ~unseen & fill)
        tp = np.sum(pred & (gt == 1)); fp = np.sum(pred & (gt == 0)); fn = np.sum(~pred & (gt == 1))
        vals.append(2 * tp / max(2 * tp + fp + fn, 1))
    # either subset of the edge cells may flip: bound loosely by the two extremes +- the edge count
    n_edge = int((edge & ~unseen).sum())
    slack = 2.0 * n_edge / max(int((gt == 1).sum()), 1)
    return min(vals) - slack, max(vals) + slack


@pytest.mark.gpu
@pytest.mark.parametrize("path", SMALL, ids=[_id(p) for p in SMALL])
def test_gpu_ig_vs_reference_golden(path):
    """Planner outputs per step against the reference's; the episode is driven with the reference's actions so
    that a near-tie argmax flip cannot desynchronise the trajectories (flips are counted and must be ties)."""
    import torch
    from ipp_marl_b200 import BatchedIPPEnv
    from oracle import numpy_ig as ni

    g = load_episode(path)
    comm = bool(g["communication"])
    orc = ni.run_ig_episode(g["params"], g["episode"], communication=comm, record_maps=True)
    env = BatchedIPPEnv(g["params"], 1, device="cuda:0")
    env.reset([g["episode"]])
    ent, f1 = env.eval_metrics()
    assert abs(float(ent[0]) - g["entropy"][0]) < 1e-5 and abs(float(f1[0]) - g["f1"][0]) < 1e-6
    T = g["action"].shape[0]
    flips = 0
    for t in range(T):
        env.observe()
        actions, gains, util = env.ig_plan(communication=comm, return_scores=True)
        torch.cuda.synchronize()
        assert np.array_equal(env.pos[0].cpu().numpy(), g["pos"][t])
        ga, ua = gains[0].cpu().numpy().astype(np.float64), util[0].cpu().numpy().astype(np.float64)
        assert np.allclose(ga, g["gains"][t], rtol=2e-4, atol=1e-7), (t, np.abs(ga - g["gains"][t]).max())
        assert np.allclose(ua, g["util"][t], rtol=2e-4, atol=1e-6), (t, np.abs(ua - g["util"][t]).max())
        act = actions[0].cpu().numpy()
        for a in range(act.shape[0]):
            if act[a] != g["action"][t][a]:
                u = g["util"][t][a]
                assert abs(u[act[a]] - u[g["action"][t][a]]) <= 2e-4 * abs(u.max()), (t, a, u, act[a])
                flips += 1
        env.act(actions=g["action"][t][None])
    # metrics of the final map (the reference evaluates after folding the last measurements)
    env.observe(final=True)
    ent, f1 = env.eval_metrics()
    torch.cuda.synchronize()
    assert abs(float(ent[0]) - g["entropy"][-1]) < 2e-5 * max(1.0, g["entropy"][-1])
    lo, hi = _f1_bounds(orc["global"][-1], env.ground_truth[0].cpu().numpy())
    assert lo - 1e-6 <= float(f1[0]) <= hi + 1e-6, (float(f1[0]), lo, hi, g["f1"][-1])
    assert flips <= 1


@pytest.mark.gpu
def test_gpu_ig_episode_metric_curves():
    """Full IG-greedy episodes on the GPU (own argmax): the entropy / F1 curves follow the reference's curve of the
    same episode (identical while the actions agree, and the final values stay close either way)."""
    import torch
    from ipp_marl_b200 import BatchedIPPEnv

    path = [p for p in SMALL if "g50_a4" in p][0]
    g = load_episode(path)
    env = BatchedIPPEnv(g["params"], 4, device="cuda:0")
    env.reset([g["episode"]] * 4)
    ents, f1s = [], []
    e, f = env.eval_metrics()
    ents.append(e.cpu().numpy().copy()); f1s.append(f.cpu().numpy().copy())
    same = True
    T = g["action"].shape[0]
    for t in range(T):
        env.observe()
        actions = env.ig_plan(communication=bool(g["communication"]))
        same = same and np.array_equal(actions[0].cpu().numpy(), g["action"][t])
        env.act(actions=actions)
        if t + 1 < T:
            continue
    env.observe(final=True)
    e, f = env.eval_metrics()
    torch.cuda.synchronize()
    e, f = e.cpu().numpy(), f.cpu().numpy()
    assert np.all(e == e[0]) and np.all(f == f[0])  # identical envs -> identical results
    if same:
        assert abs(e[0] - g["entropy"][-1]) < 1e-4 and abs(f[0] - g["f1"][-1]) < 0.01
    else:
        assert abs(e[0] - g["entropy"][-1]) < 0.05 and abs(f[0] - g["f1"][-1]) < 0.05

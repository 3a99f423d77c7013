"""CPU restatement of the reference's information-gain greedy planner and of its evaluation metrics
(SURVEY.md section 8f-3 / 8f-4).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, never by the product path.

Follows IG_baseline.py:
  * get_individual_ig     :222-285  per valid action: footprint of the candidate position, expected entropy
                                    reduction of the agent's local map over it, / 1000
  * get_relative_ig       :287-295  per agent: gains / sum of gains
  * get_cell_utilities    :297-322  sequential discount of candidates that another agent can also reach
  * select_action         :324-325  argmax
  * execute               :56-220   observe (measure / communicate / fuse) -> plan -> simultaneous moves ->
                                    measure -> global fuse -> metrics
and utils/utils.py:43-76 (``get_wrmse`` returns the F1 score of class 1) +
utils/state.py:53-121 "eval" branch (ground-truth weighted entropy) for the per-step metrics.

Parity pinned by tests/golden/ig_*.npz (the unmodified reference run through oracle/ref_harness.run_reference_ig).
"""
import numpy as np

from . import numpy_oracle as no


def update_cells_pyfloat(geo, x, y):
    """mapping/mappings.py:109-124 with a python-float measurement (clamps ``x`` in place)."""
    x[0.9999 < x] = 0.9999
    x[0.0001 > x] = 0.0001
    l_x = np.log(x / (1 - x))
    l_y = np.log(y / (1 - y))
    l_p = np.log(geo.prior / (1 - geo.prior))
    return 1 - (1 / (1 + np.exp(l_x + l_y - l_p)))


def individual_ig(geo, position, mask, map_state):
    """IG_baseline.py:222-285.  Returns (candidate positions or 0, gains)."""
    positions, gains = [], []
    for action in range(len(mask)):
        if mask[action] == 0:
            positions.append(0)
            gains.append(0)
            continue
        new_position = no.move(geo, position, action)
        fp = no.footprint(geo, new_position)[1]
        section = map_state[fp[2]:fp[3], fp[0]:fp[1]].copy()
        noise = no.NOISE_BY_ALTITUDE.get(int(new_position[2]), 0)  # sensors/models/sensor_models.py
        w1 = update_cells_pyfloat(geo, section.copy(), 1 - noise)
        w2 = update_cells_pyfloat(geo, section.copy(), noise)
        for w in (w1, w2):
            w[w > 0.501] = 1
            w[w < 0.499] = 0
        # operand order of the reference expression: the in-place clamps of get_shannon_entropy /
        # update_cells act on ``section`` before it is multiplied
        h0 = no.shannon_entropy(section)
        h1 = no.shannon_entropy(update_cells_pyfloat(geo, section, 1 - noise))
        term1 = section * (h0 - h1) * w1
        h0b = no.shannon_entropy(section)
        h2 = no.shannon_entropy(update_cells_pyfloat(geo, section, noise))
        ig = term1 + (1 - section) * (h0b - h2) * w2
        positions.append(new_position)
        gains.append(np.sum(ig) / 1000)
    return positions, gains


def relative_ig(gain_lists):
    """IG_baseline.py:287-295 (in place)."""
    for a in range(len(gain_lists)):
        total = sum(gain_lists[a])
        for k in range(len(gain_lists[a])):
            gain_lists[a][k] = gain_lists[a][k] / total
    return gain_lists


def cell_utilities(position_lists, rel):
    """IG_baseline.py:297-322 (in place, sequential: later agents see earlier agents' discounted values)."""
    for a in range(len(position_lists)):
        for k1 in range(len(position_lists[a])):
            p1 = position_lists[a][k1]
            r1 = rel[a][k1]
            for b in range(len(position_lists)):
                if b == a:
                    continue
                for k2 in range(len(position_lists[b])):
                    p2 = position_lists[b][k2]
                    r2 = rel[b][k2]
                    if np.array_equal(p1, p2) and type(p1) is np.ndarray:
                        rel[a][k1] = r1 * (1 - r2)
    return rel


def plan(geo, positions, local_maps, communication=True):
    """One planning step: masks (vs the CURRENT positions of lower-id agents, IG_baseline.py:136-153),
    gains, utilities, argmax.  Returns (masks [A,6], gains [A,6], utilities [A,6], actions [A])."""
    pos_lists, gain_lists, masks, seen = [], [], [], []
    for a in range(geo.n_agents):
        mask = no.collision_mask(geo, positions[a], no.action_mask(geo, positions[a]), seen)
        p, g = individual_ig(geo, positions[a], mask, local_maps[a])
        pos_lists.append(p)
        gain_lists.append(g)
        masks.append(mask.copy())
        seen.append(positions[a])
    gains = np.array([[float(v) for v in g] for g in gain_lists])
    rel = relative_ig(gain_lists)
    util = cell_utilities(pos_lists, rel) if communication else rel
    util = np.array([[float(v) for v in u] for u in util])
    return np.array(masks), gains, util, np.argmax(util, axis=1).astype(np.int64)


def eval_metrics(global_map, gt):
    """(masked entropy, F1 of class 1): IG_baseline.py:81-100 / 191-210, utils/utils.py:43-76."""
    g = np.array(global_map, copy=True)
    h = no.shannon_entropy(g)  # utils/state.py "eval": weights = ground truth (0 / 1), then masked by it again
    _, counts = np.unique(gt, return_counts=True)
    target = counts[-1]
    ent = np.sum(np.where(gt == 0, 0, h * (gt != 0))) / target
    pred = np.asarray(global_map) > 0.5
    tp = np.sum(pred & (gt == 1))
    fp = np.sum(pred & (gt == 0))
    fn = np.sum(~pred & (gt == 1))
    f1 = 2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else 0.0
    return float(ent), float(f1)


def fuse_list(geo, own, others):
    """mapping/mappings.py:99-102, the list branch of the global fusion: unlike the dict branch the other
    maps are NOT cast to float32 (they carry the dtype of the sender's local map)."""
    fused = np.float32(own.copy())
    for other in others:
        fused = no.bayes_pass(fused, other, geo.prior)
    return fused


def run_ig_episode(params, episode, communication=True, actions=None, record_maps=False):
    """IG_baseline.execute in oracle terms.  ``actions`` [T, A] replays given moves instead of the argmax."""
    env = no.OracleEnv(params, episode)
    g = env.geo
    rec = {"gains": [], "util": [], "action": [], "mask": [], "pos": [], "entropy": [], "f1": [], "global": []}
    e, f = eval_metrics(env.global_map, env.gt)
    rec["entropy"].append(e)
    rec["f1"].append(f)
    ig_global = None
    for t in range(g.budget + 1):
        env.observe()
        if t == 0:
            ig_global = env.global_map  # IG_baseline.py:121-125 (dict branch, same as the COMA loop)
        rec["pos"].append(np.array(env.pos, dtype=np.int64))
        masks, gains, util, acts = plan(g, env.pos, env.local, communication)
        if actions is not None:
            acts = np.asarray(actions[t], dtype=np.int64)
        env.act(acts)
        ig_global = fuse_list(g, ig_global, list(env.m2c))  # IG_baseline.py:172-175
        e, f = eval_metrics(ig_global, env.gt)
        if record_maps:
            rec["global"].append(np.asarray(ig_global, dtype=np.float64).copy())
        for k, v in (("gains", gains), ("util", util), ("action", acts), ("mask", masks), ("entropy", e), ("f1", f)):
            rec[k].append(v)
    return {k: np.array(v) for k, v in rec.items()}

"""Write tests/golden/* from the UNMODIFIED reference imported from /root/reference.

Run in the build container only:  ``python -m oracle.make_golden``
(the GPU box has no /root/reference; it uses the committed fixtures).

The reference holds no tests / golden vectors of its own, so
these files — outputs of the reference itself with the noise / action / comm-failure
draws replaced by ``oracle.noise`` (see oracle/ref_harness.py for exactly what is
patched) — are what pins the oracle and, through it, the CUDA path.
numpy/torch versions are recorded in each file: the reference's dtype flow depends
on them (SURVEY.md section 7 "dtype drift").
"""
import json
import os

import numpy as np

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _versions():
    import torch

    return {"numpy": np.__version__, "torch": torch.__version__}


def _pack_episode(rec, keep_steps=None, maps=True):
    out = {"gt": rec["gt"].astype(np.uint8)}
    T = len(rec["steps"])
    for key in ("pos", "comm", "mask", "action", "pos_next"):
        out[key] = np.stack([s[key] for s in rec["steps"]])
    out["reward_rel"] = np.array([s["reward_rel"] for s in rec["steps"]])
    out["reward_abs"] = np.array([s["reward_abs"] for s in rec["steps"]])
    steps = list(range(T)) if keep_steps is None else [t for t in keep_steps if t < T]
    out["map_steps"] = np.array(steps, dtype=np.int64)
    if maps:
        for key in ("global", "local_fused", "local_after_move"):
            out[key] = np.stack([rec["steps"][t][key] for t in steps])
    if "obs" in rec["steps"][0]:  # network-input features, every step
        out["obs"] = np.stack([s["obs"] for s in rec["steps"]])      # [T, A, P, P, 7] float64
        out["state"] = np.stack([s["state"] for s in rec["steps"]])  # [T, A, P, P, 12] float32
    # checksums of every step (cheap, lets big configs be pinned without storing maps)
    for key in ("global", "local_fused", "local_after_move"):
        out[key + "_sum"] = np.array([rec["steps"][t][key].sum() for t in range(T)])
        out[key + "_sumsq"] = np.array([(rec["steps"][t][key] ** 2).sum() for t in range(T)])
    return out


def x_of(params):
    return params["environment"]["x_dim"]


def episodes(only=None, default=True):
    cases = [
        # name, params, episodes, keep_steps
        ("g50_a4", rh.synthetic_params(50, 4), (1, 2, 5), None),
        ("g50_a2", rh.synthetic_params(50, 2), (1, 3), None),
        ("g50_a4_comm15_fail30", rh.synthetic_params(50, 4, comm_range=15, failure_rate=0.3), (4,), None),
        ("g50_a3_prior40", rh.synthetic_params(50, 3, comm_range=100, prior=0.4), (2,), (0, 1, 7, 14)),
        ("g100_a8", rh.synthetic_params(100, 8), (1,), (0, 1, 7, 14)),
        # fix_range False (agent/communication_log.py:22-31): episodes 1 / 2 / 5 draw the ranges 15 / 0 / 100 m
        ("g50_a4_randrange", rh.synthetic_params(50, 4, fix_range=False), (1, 2, 5), (0, 1, 7, 14)),
    ]
    if only is not None:
        cases = [c for c in cases if c[0] in only]
    for name, params, eps, keep in cases:
        for ep in eps:
            rec = rh.run_reference_episode(params, ep, features=(x_of(params) == 50))
            out = _pack_episode(rec, keep)
            out["params_json"] = np.array(json.dumps(params))
            out["episode"] = np.array(ep)
            out["versions_json"] = np.array(json.dumps(_versions()))
            path = os.path.join(OUT, "episode_%s_ep%d.npz" % (name, ep))
            np.savez_compressed(path, **out)
            print("wrote", path, os.path.getsize(path))
    if not default:
        return
    # reference default (G = 493): rewards, moves, checksums and the final global map only
    params = rh.default_params()
    rec = rh.run_reference_episode(params, 1, features=True)
    out = _pack_episode(rec, keep_steps=(14,), maps=False)
    out["global_final_f32"] = rec["steps"][-1]["global"].astype(np.float32)
    out["params_json"] = np.array(json.dumps(params))
    out["episode"] = np.array(1)
    out["versions_json"] = np.array(json.dumps(_versions()))
    path = os.path.join(OUT, "episode_default_g493_a4_ep1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


def kats():
    """Known-answer values of the individual reference functions."""
    ns = rh.load()
    rh.install_noise_patch()
    out = {"versions": _versions()}
    for tag, params in (("default", rh.default_params()), ("synthetic50", rh.synthetic_params(50, 4)),
                        ("synthetic100", rh.synthetic_params(100, 8))):
        gm = ns.GridMap(params)
        cam = ns.Camera(params, ns.AltitudeSensorModel(params), gm)
        ass = ns.AgentStateSpace(params)
        aas = ns.AgentActionSpace(params)
        entry = {
            "params": params,
            "res_x": gm.res_x,
            "res_y": gm.res_y,
            "gx": gm.x_dim,
            "gy": gm.y_dim,
            "lattice": [int(v) for v in ass.space_dim],
        }
        fov = []
        xmax = params["environment"]["x_dim"]
        for x in range(0, xmax + 1, 5):
            for y in range(0, xmax + 1, 5):
                for z in (5, 10, 15):
                    raw, clipped = cam.project_field_of_view(np.array([x, y, z]), gm.resolution_x, gm.resolution_y)
                    fov.append([[x, y, z], raw, clipped])
        entry["fov"] = fov
        entry["start"] = [
            [ep, a, [int(v) for v in ass.get_random_agent_state(a, ep)]] for ep in range(1, 40) for a in range(8)
        ]
        gts = []
        for ep in range(1, 40):
            sensor = ns.Sensor(ns.AltitudeSensorModel(params), gm)
            if tag == "default" and ep > 3:
                break
            m = ns.Mapping(gm, sensor, params, ep)
            f = m.simulated_map
            gts.append([ep, int(f.sum()), int(f[0, 0]), int(f[-1, 0]), int(f[0, -1]), int(f[-1, -1]),
                        int(f[:, 0].sum()), int(f[0, :].sum())])
        entry["gt"] = gts
        masks = []
        rng = np.random.RandomState(7)
        P = xmax // 5 + 1
        for _ in range(300):
            pos = np.array([5 * rng.randint(P), 5 * rng.randint(P), 5 * rng.randint(1, 4)])
            others = [np.array([pos[0] + 5 * rng.randint(-1, 2), pos[1] + 5 * rng.randint(-1, 2), 5 * rng.randint(1, 4)])
                      for _ in range(rng.randint(0, 4))]
            others = [o for o in others if 0 <= o[0] <= xmax and 0 <= o[1] <= xmax]
            m0, _ = aas.get_action_mask(pos)
            m1 = aas.apply_collision_mask(pos, m0.copy(), others, ass)
            masks.append([pos.tolist(), [o.tolist() for o in others], m0.tolist(), m1.tolist()])
        entry["masks"] = masks
        out[tag] = entry
    # Bayes update KAT (mapping/mappings.py:109-124)
    params = rh.synthetic_params(50, 4)
    gm = ns.GridMap(params)
    m = ns.Mapping(gm, ns.Sensor(ns.AltitudeSensorModel(params), gm), params, 1)
    xs = [0.5, 0.625, 0.99, 0.9999, 0.99999, 0.0001, 0.3, 0.0]
    upd = {}
    for y in (0.99, 0.01, 0.735, 0.265, 0.625, 0.375, 0.5):
        x = np.array(xs, dtype=np.float32)
        r = m.apply_update(x, np.float32(y), None)
        upd[str(y)] = {"out": [float(v) for v in r], "dtype": str(r.dtype), "x_after": [float(v) for v in x]}
    out["apply_update"] = {"x": xs, "y": upd}
    # python-float measurement (IG_baseline.py:240-245)
    x = np.array(xs, dtype=np.float32)
    r = m.update_cells(x, 0.99, None)
    out["apply_update_pyfloat"] = {"out": [float(v) for v in r], "dtype": str(r.dtype)}
    # entropy KAT (utils/state.py:118-121)
    pe = np.array([0.5, 0.625, 0.99, 0.9999, 0.99999, 0.00001, 0.3], dtype=np.float64)
    out["entropy"] = {"p": pe.tolist(), "H": [float(v) for v in ns.get_shannon_entropy(pe.copy())]}
    # noiseless reward chain
    chain = {}
    for tag, params in (("default", rh.default_params()), ("synthetic50", rh.synthetic_params(50, 4))):
        gm = ns.GridMap(params)
        mp = ns.Mapping(gm, ns.Sensor(ns.AltitudeSensorModel(params), gm), params, 1)
        ass = ns.AgentStateSpace(params)
        rh.NoiseContext.noiseless = True
        glob = mp.init_priors()
        steps = []
        for poses in ([[25, 25, 15]], [[25, 25, 10], [30, 25, 15]], [[25, 25, 5], [30, 30, 5]]):
            m2cs = []
            for pz in poses:
                _, _, _, m2c, _ = mp.update_grid_map(np.array(pz), mp.init_priors(), 0, None)
                m2cs.append(m2c)
            fused = mp.fuse_map(glob, m2cs, None, "global")
            _, rel, ab = ns.get_global_reward(glob, fused, "COMA", None, mp.simulated_map, ass, None, None, 0, 14)
            steps.append({"poses": poses, "rel": float(rel), "abs": float(ab), "sum": float(fused.sum()),
                          "max": float(fused.max()), "min": float(fused.min())})
            glob = fused
        rh.NoiseContext.noiseless = False
        chain[tag] = steps
    out["reward_chain"] = chain
    path = os.path.join(OUT, "kats.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path))


def ig_episodes():
    """The reference's IG-greedy baseline (IG_baseline.py) run unmodified: per-step gains, utilities, argmax
    actions and the entropy / F1 curves."""
    cases = [
        ("g50_a4", rh.synthetic_params(50, 4), 2),
        ("g50_a3_comm15_fail30", rh.synthetic_params(50, 3, comm_range=15, failure_rate=0.3), 5),
        ("g50_a2", rh.synthetic_params(50, 2), 7),
        ("g100_a8", rh.synthetic_params(100, 8), 1),
        ("default_g493_a4", rh.default_params(), 1),
    ]
    for name, params, ep in cases:
        rec = rh.run_reference_ig(params, ep)
        out = {k: rec[k] for k in ("pos", "gains", "util", "action", "entropy", "f1")}
        out["gt_sum"] = np.array(int(rec["gt"].sum()))
        out["communication"] = np.array(bool(params["experiment"]["baselines"]["information_gain"]["communication"]))
        out["params_json"] = np.array(json.dumps(params))
        out["episode"] = np.array(ep)
        out["versions_json"] = np.array(json.dumps(_versions()))
        path = os.path.join(OUT, "ig_%s_ep%d.npz" % (name, ep))
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


def lawn_episodes():
    """The reference's coverage baseline (lawn_mower.py:38-315, eight hard-coded paths) run unmodified through
    tests/ref_callers.py: entropy / F1 curves and the final map (strided sample + checksums for the 493x493 grid)."""
    import subprocess
    import sys
    import tempfile

    root = os.path.dirname(os.path.dirname(OUT))
    for name, params, ep in (("g50_a8", rh.synthetic_params(50, 8), 2), ("default_g493_a8", dict(rh.default_params()), 1)):
        params = json.loads(json.dumps(params))
        params["experiment"]["missions"]["n_agents"] = 8  # lawn_mower.py hard-codes eight agents
        with tempfile.TemporaryDirectory() as tmp:
            pj, o = os.path.join(tmp, "p.json"), os.path.join(tmp, "o.npz")
            with open(pj, "w") as f:
                json.dump(params, f)
            subprocess.run([sys.executable, os.path.join(root, "tests", "ref_callers.py"), "lawn", pj, str(ep), o, "ref"],
                           check=True)
            z = np.load(o)
            out = {"entropy": z["entropy"], "f1": z["f1"], "update_calls": z["update_calls"],
                   "map_sum": np.array(z["map"].sum()), "map_sumsq": np.array((z["map"] ** 2).sum()),
                   "map_sample": z["map"][::7, ::7].copy() if z["map"].shape[0] > 100 else z["map"].copy(),
                   "map_sample_stride": np.array(7 if z["map"].shape[0] > 100 else 1)}
        out["params_json"] = np.array(json.dumps(params))
        out["episode"] = np.array(ep)
        out["versions_json"] = np.array(json.dumps(_versions()))
        path = os.path.join(OUT, "lawn_%s_ep%d.npz" % (name, ep))
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import sys

    if "ig" in sys.argv[1:]:
        ig_episodes()
    elif "lawn" in sys.argv[1:]:
        lawn_episodes()
    elif "randrange" in sys.argv[1:]:
        episodes(only=("g50_a4_randrange",), default=False)
    else:
        kats()
        episodes()
        ig_episodes()
        lawn_episodes()

"""Shared helpers for the parity tests."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_episodes(pattern="episode_*.npz"):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load_episode(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["params"] = json.loads(str(d.pop("params_json")))
    d["versions"] = json.loads(str(d.pop("versions_json")))
    d["episode"] = int(d["episode"])
    return d


def load_kats():
    with open(os.path.join(GOLDEN, "kats.json")) as f:
        return json.load(f)


def gate_stats(ref, got, rtol=1e-5, atol=1e-5):
    """SURVEY.md section 8d parity gate + the extra figures it asks to report."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    d = np.abs(ref - got)
    return {
        "max_abs": float(d.max()) if d.size else 0.0,
        "fail_gate": int((d > atol + rtol * np.abs(ref)).sum()),
        "fail_pure_rtol": int((d > rtol * np.abs(ref)).sum()),
        "rel_l2": float(np.sqrt((d ** 2).sum() / max((ref ** 2).sum(), 1e-300))),
        "n": int(d.size),
    }

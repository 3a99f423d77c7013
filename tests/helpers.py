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


def f1_bounds(map_state, gt, eps=2e-5, seen=None):
    """F1 of class 1 (utils/utils.py:43-76 thresholds at > 0.5) with the cells within eps of 0.5 counted either way:
    cancelling evidence (a cell seen as 1 and as 0 from the same altitude) leaves p = 0.5 +- 1e-8, and which side it
    lands on is the last bit of numpy's float32 log in the reference.  -> (lowest, highest) attainable F1."""
    map_state = np.asarray(map_state, dtype=np.float64)
    gt = np.asarray(gt)
    sure1 = map_state > 0.5 + eps
    edge = np.abs(map_state - 0.5) <= eps
    # never-observed cells are exactly 0.5 on both sides; observed ones can cancel to exactly 0.5 on one side only
    edge &= (map_state != 0.5) if seen is None else np.asarray(seen, dtype=bool)
    tp = np.sum(sure1 & (gt == 1))
    fp = np.sum(sure1 & (gt == 0))
    fn = np.sum(~sure1 & (gt == 1))
    e1 = int(np.sum(edge & (gt == 1)))
    e0 = int(np.sum(edge & (gt == 0)))
    f1 = lambda tp_, fp_, fn_: 2.0 * tp_ / max(2 * tp_ + fp_ + fn_, 1)
    return f1(tp, fp + e0, fn), f1(tp + e1, fp, fn - e1)

"""The reference's UNCHANGED caller files on top of the CUDA-backed facade (north star: "IG_baseline.py, lawn_mower.py
and coma_wrapper.py drop in unchanged").

tests/ref_callers.py executes ``EpisodeGenerator.execute`` (-> ``COMAWrapper.build_observations`` / ``.steps``,
coma_wrapper.py:37-183), ``IG_baseline.execute`` (IG_baseline.py:56-220) and ``LawnMower.execute``
(lawn_mower.py:38-315) from the verbatim reference install ``baseline/_ref`` (scripts/install_ref.py; it travels to the
GPU box with the snapshot) and dumps what they compute.  GPU tests run them with every environment import resolved to
``ipp_marl_b200.facade`` and compare with the golden outputs of the same callers on the reference's own modules; the
CPU test re-checks those goldens against the live reference.  Gate: SURVEY.md section 8d, allclose(1e-5, 1e-5).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import GOLDEN, gate_stats, load_episode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "ref_callers.py")


def _have_reference():
    from oracle import ref_harness as rh

    return rh.available()


def _run(mode, params, episode, side, tmp_path):
    pj, out = str(tmp_path / "p.json"), str(tmp_path / ("%s_%s.npz" % (mode, side)))
    with open(pj, "w") as f:
        json.dump(params, f)
    res = subprocess.run([sys.executable, SCRIPT, mode, pj, str(episode), out, side], capture_output=True, text=True,
                         timeout=1500)
    assert res.returncode == 0, res.stderr[-3000:]
    return np.load(out)


def _gate(ref, got, what):
    s = gate_stats(ref, got)
    assert s["fail_gate"] == 0, (what, s)
    return s


def _check_coma(r, g, maps=True):
    assert np.array_equal(r["gt"], g["gt"])
    assert np.array_equal(r["pos"][:-1], g["pos"]) and np.array_equal(r["pos"][1:], g["pos_next"])
    assert np.array_equal(r["action"], g["action"])
    assert np.allclose(r["reward_rel"], g["reward_rel"], rtol=1e-5, atol=1e-5)
    assert np.allclose(r["reward_abs"], g["reward_abs"], rtol=1e-5, atol=1e-5)
    if maps:
        _gate(g["global"], r["global"][g["map_steps"]], "global map")
    else:
        assert np.allclose(r["global"].sum(axis=(1, 2)), g["global_sum"], rtol=1e-7)
        _gate(g["global_final_f32"], r["global"][-1], "final global map")


def _check_f1(r, f1_ref):
    """F1 thresholds the map at > 0.5 (utils/utils.py:66-68); cells where evidence cancels sit at 0.5 +- 1e-8 and
    their side is decided by the last bit of numpy's float32 log in the reference.  The reference's F1 must lie in
    the range the facade's map allows when exactly those cells are counted either way, and our own F1 inside it too."""
    assert len(r["f1"]) == len(f1_ref) == len(r["f1_lo"])
    assert np.all(r["f1_lo"] - 1e-12 <= f1_ref) and np.all(f1_ref <= r["f1_hi"] + 1e-12), (r["f1_lo"], f1_ref, r["f1_hi"])
    assert np.all(r["f1_lo"] - 1e-12 <= r["f1"]) and np.all(r["f1"] <= r["f1_hi"] + 1e-12)


def test_unchanged_callers_reproduce_the_goldens_on_the_reference_itself(tmp_path):
    """Pins the runner: coma_wrapper / episode_generator driven by tests/ref_callers.py on the reference's own
    modules give the committed golden episode bit for bit (incl. the observation / critic-state tensors)."""
    if not _have_reference():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    g = load_episode(os.path.join(GOLDEN, "episode_g50_a4_ep2.npz"))
    r = _run("coma", g["params"], g["episode"], "ref", tmp_path)
    _check_coma(r, g)
    assert np.array_equal(r["global"], g["global"]) and np.array_equal(r["reward_rel"], g["reward_rel"])
    assert np.array_equal(r["obs"], g["obs"]) and np.array_equal(r["state"], g["state"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g50_a4_ep2", "g50_a4_comm15_fail30_ep4", "g50_a2_ep3", "default_g493_a4_ep1"])
def test_coma_wrapper_unchanged_on_facade(tmp_path, name):
    if not _have_reference():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    g = load_episode(os.path.join(GOLDEN, "episode_%s.npz" % name))
    r = _run("coma", g["params"], g["episode"], "facade", tmp_path)
    _check_coma(r, g, maps="global" in g)
    # network inputs built by the reference's actor/critic transformations from facade outputs
    # The weighted-entropy channels (3: local, 8: global) multiply H by w = 1 / 0 / 0.5 chosen by comparing the POOLED
    # probability (channels 5 / 9) with 0.501 / 0.499 (utils/state.py:67-73): a pooled cell within 1e-6 of a threshold
    # may take either weight (its probability itself agrees to 1e-7), so those cells are exempt in channels 3 / 8.
    ref_o, got_o = np.array(g["obs"], dtype=np.float64), np.array(r["obs"], dtype=np.float64)
    ref_s, got_s = np.array(g["state"], dtype=np.float64), np.array(r["state"], dtype=np.float64)

    def edge(prob):
        return (np.abs(prob - 0.501) < 1e-6) | (np.abs(prob - 0.499) < 1e-6)

    e_loc, e_glob = edge(ref_o[..., 5]), edge(ref_s[..., 9])
    got_o[..., 3] = np.where(e_loc, ref_o[..., 3], got_o[..., 3])
    got_s[..., 3] = np.where(e_loc, ref_s[..., 3], got_s[..., 3])
    got_s[..., 8] = np.where(e_glob, ref_s[..., 8], got_s[..., 8])
    assert e_loc.mean() < 0.01 and e_glob.mean() < 0.01  # a pooled cell on a threshold stays there while unobserved
    so = gate_stats(ref_o, got_o, atol=2e-5)
    ss = gate_stats(ref_s, got_s, atol=2e-5)
    assert so["fail_gate"] == 0 and ss["fail_gate"] == 0, (so, ss)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g50_a4_ep2", "g50_a3_comm15_fail30_ep5", "default_g493_a4_ep1"])
def test_ig_baseline_unchanged_on_facade(tmp_path, name):
    if not _have_reference():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    z = np.load(os.path.join(GOLDEN, "ig_%s.npz" % name))
    params, ep = json.loads(str(z["params_json"])), int(z["episode"])
    r = _run("ig", params, ep, "facade", tmp_path)
    assert np.array_equal(r["action"], z["action"])
    assert np.allclose(r["gains"], z["gains"], rtol=1e-5, atol=1e-7)
    assert np.allclose(r["entropy"], z["entropy"], rtol=1e-5, atol=1e-6)
    _check_f1(r, z["f1"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g50_a8_ep2", "default_g493_a8_ep1"])
def test_lawn_mower_unchanged_on_facade(tmp_path, name):
    if not _have_reference():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    z = np.load(os.path.join(GOLDEN, "lawn_%s.npz" % name))
    params, ep = json.loads(str(z["params_json"])), int(z["episode"])
    r = _run("lawn", params, ep, "facade", tmp_path)
    assert int(r["update_calls"]) == int(z["update_calls"])
    assert np.allclose(r["entropy"], z["entropy"], rtol=1e-5, atol=1e-6)
    _check_f1(r, z["f1"])
    k = int(z["map_sample_stride"])
    _gate(z["map_sample"], r["map"][::k, ::k], "final map")
    assert abs(float(r["map"].sum()) - float(z["map_sum"])) <= 1e-6 * float(z["map_sum"])

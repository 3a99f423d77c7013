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
    so = gate_stats(g["obs"], r["obs"], atol=2e-5)
    ss = gate_stats(g["state"], r["state"], atol=2e-5)
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
    assert np.allclose(r["f1"], z["f1"], rtol=0, atol=1e-12)


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
    assert np.allclose(r["f1"], z["f1"], rtol=0, atol=1e-12)
    k = int(z["map_sample_stride"])
    _gate(z["map_sample"], r["map"][::k, ::k], "final map")
    assert abs(float(r["map"].sum()) - float(z["map_sum"])) <= 1e-6 * float(z["map_sum"])

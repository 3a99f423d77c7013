"""The CPU arm of bench.py (`--impl reference`, `cpu_baseline`): the verbatim reference install, the env-only loop built
from the reference's own objects (oracle/ref_timing.py) and the contract figures bench.py divides by.  CPU only."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402
from scripts import install_ref  # noqa: E402


def test_contract_bytes_per_env_step():
    """G^2 * (2 * 4 * (A + 1) + 1) — the figures the roofline line is built on."""
    assert bench.bytes_per_env_step(50, 2) == 62_500
    assert bench.bytes_per_env_step(50, 4) == 102_500
    assert bench.bytes_per_env_step(100, 8) == 730_000
    assert bench.bytes_per_env_step(493, 4) == 9_965_009


def test_bench_parameter_trees():
    from ipp_marl_b200.geometry import HostTables

    for grid, agents, g in ((50, 4, 50), (100, 8, 100), (493, 2, 493)):
        t = HostTables(bench.kat_params(agents, grid))
        assert (t.gx, t.gy, t.n_agents) == (g, g, agents)


def test_reference_install_is_verbatim():
    if not install_ref.installed():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    assert install_ref.verify()  # every installed file still has the digest recorded at install time
    src = install_ref.source_root()
    if os.path.isdir(os.path.join(src, "marl_framework")):  # build container: byte-identical to the source tree
        for rel in ("marl_framework/coma_wrapper.py", "marl_framework/mapping/mappings.py", "marl_framework/params.yaml"):
            with open(os.path.join(src, rel), "rb") as a, open(os.path.join(install_ref.DEST, rel), "rb") as b:
                assert a.read() == b.read(), rel


def test_reference_env_loop_runs_one_episode():
    """oracle/ref_timing.env_loop_episode: one episode of the reference's own objects in coma_wrapper's call order
    (what `bench.py --impl reference` times) — budget + 1 env-steps, every agent inside the map at the end."""
    from oracle import ref_timing

    if not ref_timing.available():
        pytest.skip("reference tree not installed (scripts/install_ref.py)")
    ns = ref_timing._load()
    params = bench.kat_params(2, 50)
    steps, gt_seconds = ref_timing.env_loop_episode(ns, params, 3)
    assert steps == params["experiment"]["constraints"]["budget"] + 1 and gt_seconds >= 0.0
    r = ref_timing.run(params, 0.2, 1)
    assert r["steps"] >= steps and r["value"] > 0 and 0.0 <= r["ground_truth_share"] <= 1.0

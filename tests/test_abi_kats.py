"""Known-answer tests of the single-map C-ABI entry points against the values the UNMODIFIED reference produced
(tests/golden/kats.json, written by oracle/make_golden.py): ipp_update_cells (array and
Python-float measurement, in-place clamp), ipp_shannon_entropy (in-place clamp), ipp_project_fov (every lattice
position of three geometries), ipp_measure (values and noise bits), and ipp_fuse_map + ipp_utility_reward through the
reward-chain KAT on both grids.  Every call goes through ctypes on the shared library, nothing else."""
import ctypes as C

import numpy as np
import pytest

from tests.helpers import load_kats

pytestmark = pytest.mark.gpu


def _rt(params):
    from ipp_marl_b200.facade import _runtime as R

    return R, R.runtime(params)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_update_cells_kat_float32_and_python_float():
    k = load_kats()
    R, rt = _rt(k["synthetic50"]["params"])
    xs = k["apply_update"]["x"]
    worst = 0.0
    for y_str, ent in k["apply_update"]["y"].items():
        assert ent["dtype"] == "float64"  # the reference returns float64 under numpy >= 2
        x = np.array(xs, dtype=np.float32)
        y = np.full(1, float(y_str), dtype=np.float32)  # make_golden passed np.float32(y): float32 logit
        out = np.empty(x.size, dtype=np.float64)
        rt.check(rt.lib.ipp_update_cells(rt.h, _ptr(x), 0, _ptr(y), 0, 1, x.size, _ptr(out)), "ipp_update_cells")
        ref = np.array(ent["out"])
        assert np.allclose(out, ref, rtol=1e-5, atol=1e-5), (y_str, out, ref)
        worst = max(worst, float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300))))
        assert np.array_equal(x, np.array(ent["x_after"], dtype=np.float32)), "in-place clamp (mappings.py:110-111)"
        # the same measurement as a per-cell float32 array
        x2 = np.array(xs, dtype=np.float32)
        ya = np.full(x2.size, float(y_str), dtype=np.float32)
        out2 = np.empty_like(out)
        rt.check(rt.lib.ipp_update_cells(rt.h, _ptr(x2), 0, _ptr(ya), 0, 0, x2.size, _ptr(out2)), "ipp_update_cells")
        assert np.array_equal(out2, out)
    assert worst < 2e-6, worst  # far inside the gate: only the last bit of the float32 logs can differ
    # Python-float measurement (IG_baseline.py:240-245): float64 logit of y
    x = np.array(xs, dtype=np.float32)
    y = np.full(1, 0.99, dtype=np.float64)
    out = np.empty(x.size, dtype=np.float64)
    rt.check(rt.lib.ipp_update_cells(rt.h, _ptr(x), 0, _ptr(y), 1, 1, x.size, _ptr(out)), "ipp_update_cells")
    ref = np.array(k["apply_update_pyfloat"]["out"])
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-5)
    assert float(np.max(np.abs(out - ref) / np.abs(ref))) < 2e-6
    # float64 map section (a local map that came out of a fuse): clamp and logit in float64
    x = np.array([0.5, 0.99995, 0.3, 1e-5], dtype=np.float64)
    y = np.full(4, 0.735, dtype=np.float32)
    out = np.empty(4, dtype=np.float64)
    rt.check(rt.lib.ipp_update_cells(rt.h, _ptr(x), 1, _ptr(y), 0, 0, 4, _ptr(out)), "ipp_update_cells")
    xc = np.clip(np.array([0.5, 0.99995, 0.3, 1e-5]), 0.0001, 0.9999)
    assert np.array_equal(x, xc)
    ly = np.log(y / (1 - y))
    ref = 1 - 1 / (1 + np.exp(np.log(xc / (1 - xc)) + ly - 0.0))
    assert np.allclose(out, ref, rtol=1e-6, atol=0)  # float32 log of y: numpy's SIMD logf vs a correctly rounded one


def test_shannon_entropy_kat_and_in_place_clamp():
    k = load_kats()
    R, rt = _rt(k["synthetic50"]["params"])
    p = np.array(k["entropy"]["p"], dtype=np.float64)
    out = np.empty_like(p)
    rt.check(rt.lib.ipp_shannon_entropy(rt.h, _ptr(p), 1, p.size, _ptr(out)), "ipp_shannon_entropy")
    assert np.allclose(out, np.array(k["entropy"]["H"]), rtol=1e-12, atol=1e-15)
    assert np.array_equal(p, np.clip(np.array(k["entropy"]["p"]), 0.0001, 0.9999))  # utils/state.py:119-120
    p32 = np.array(k["entropy"]["p"], dtype=np.float32)
    o32 = np.empty_like(p32)
    rt.check(rt.lib.ipp_shannon_entropy(rt.h, _ptr(p32), 0, p32.size, _ptr(o32)), "ipp_shannon_entropy")
    pc = np.clip(np.array(k["entropy"]["p"], dtype=np.float32), np.float32(0.0001), np.float32(0.9999))
    assert np.array_equal(p32, pc)
    assert np.allclose(o32, -pc * np.log2(pc) - (1 - pc) * np.log2(1 - pc), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["default", "synthetic50", "synthetic100"])
def test_project_fov_every_lattice_position(tag):
    k = load_kats()[tag]
    R, rt = _rt(k["params"])
    raw, clipped = (C.c_int32 * 4)(), (C.c_int32 * 4)()
    assert len(k["fov"]) >= 3 * 11 * 11
    for pos, r_ref, c_ref in k["fov"]:
        p = (C.c_int32 * 3)(*pos)
        rt.check(rt.lib.ipp_project_fov(rt.h, p, raw, clipped), "ipp_project_fov")
        assert list(raw) == r_ref and list(clipped) == c_ref, (pos, list(raw), r_ref, list(clipped), c_ref)


@pytest.mark.parametrize("altitude", [5, 10, 15])
def test_measure_values_and_noise_bits(altitude):
    """Simulation.get_measurement (mapping/simulations.py:42-65): values in {1-noise, noise} rounded to 3 digits as
    float32, wrong where the stream's noise word falls below the altitude's flip threshold."""
    from oracle import noise as hn

    k = load_kats()["synthetic50"]
    params = k["params"]
    R, rt = _rt(params)
    G = k["gx"]
    rng = np.random.RandomState(altitude)
    gt = (rng.rand(G, G) < 0.4).astype(np.uint8)
    yu, yd, xl, xr = 3, 41, 7, 49
    noise = {5: 0.01, 10: 0.265, 15: 0.375}[altitude]
    y_hi, y_lo = np.float32(np.round(1 - noise, 3)), np.float32(np.round(noise, 3))
    out = np.empty((xr - xl, yd - yu), dtype=np.float32)
    rect = (C.c_int32 * 4)(yu, yd, xl, xr)
    ep, agent, index = 7, 2, 5
    rt.check(rt.lib.ipp_measure(rt.h, _ptr(gt), rect, altitude, ep, agent, index, float(y_hi), float(y_lo), _ptr(out)),
             "ipp_measure")
    cells = np.arange(xl, xr)[:, None] * G + np.arange(yu, yd)[None, :]
    key = hn.stream_key(params["environment"]["seed"], ep, agent, index, hn.PURPOSE_NOISE)
    wrong = hn.noise_word(key, cells) < hn.flip_threshold(noise)
    seen = (gt[xl:xr, yu:yd] != 0) != wrong
    assert np.array_equal(out, np.where(seen, y_hi, y_lo).astype(np.float32))
    assert set(np.unique(out)) <= {y_hi, y_lo}
    frac = wrong.mean()
    assert abs(frac - noise) < 0.04


@pytest.mark.parametrize("tag", ["synthetic50", "default"])
def test_reward_chain_kat_through_fuse_and_utility(tag):
    """SURVEY.md section 8c reward chain (noiseless): update_grid_map on fresh prior maps -> fuse_map(global, [m2c...])
    -> get_global_reward, three steps, on the synthetic 50x50 and the default 493x493 grid."""
    k = load_kats()
    params = k[tag]["params"]
    R, rt = _rt(params)
    from ipp_marl_b200.geometry import HostTables

    tb = HostTables(params)
    G = tb.gx
    # ground truth of episode 1 (mapping/ground_truths.py:42-56): split_idx 1, pct 41 -> last rows = 1
    gt = np.zeros((G, G))
    gt[int(G * (1 - 41) / 100):, :] = 1
    glob = np.full((G, G), 0.5, dtype=np.float32)
    raw, clipped = (C.c_int32 * 4)(), (C.c_int32 * 4)()
    for step in k["reward_chain"][tag]:
        m2cs = []
        for pos in step["poses"]:
            rt.check(rt.lib.ipp_project_fov(rt.h, (C.c_int32 * 3)(*pos), raw, clipped), "ipp_project_fov")
            yu, yd, xl, xr = list(clipped)
            iz = tb.altitudes.index(pos[2])
            m2c = np.full((G, G), 0.5, dtype=np.float32)
            m2c[xl:xr, yu:yd] = np.where(gt[xl:xr, yu:yd] == 1, tb.y_hi[iz], tb.y_lo[iz])
            m2cs.append(m2c)
        stack = np.ascontiguousarray(np.stack(m2cs))
        own = np.ascontiguousarray(glob, dtype=np.float32)
        fused = np.empty((G, G), dtype=np.float64)
        rt.check(rt.lib.ipp_fuse_map(rt.h, _ptr(own), _ptr(stack), len(m2cs), own.size, _ptr(fused)), "ipp_fuse_map")
        out = np.zeros(2)
        last = np.ascontiguousarray(glob)
        rt.check(rt.lib.ipp_utility_reward(rt.h, _ptr(last), 1 if last.dtype == np.float64 else 0, _ptr(fused), 1,
                                           last.size, _ptr(out)), "ipp_utility_reward")
        rel, ab = 22 * out[1] - 0.5, 10 * out[0] - 0.17
        assert abs(rel - step["rel"]) <= 1e-5 + 1e-5 * abs(step["rel"]), (rel, step["rel"])
        assert abs(ab - step["abs"]) <= 1e-5 + 1e-5 * abs(step["abs"]), (ab, step["abs"])
        assert abs(fused.sum() - step["sum"]) <= 1e-7 * step["sum"]
        assert abs(fused.max() - step["max"]) <= 1e-6 and abs(fused.min() - step["min"]) <= 1e-6 * 1e-2 + 1e-9
        glob = fused

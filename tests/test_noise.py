"""CPU: the counter-based hash specification (oracle/noise.py) — determinism, marginals, independence."""
import numpy as np

from oracle import noise as hn


def test_known_answers():
    # pins the bit pattern of the hash: the CUDA side (ipp_device.cuh) implements the same functions
    assert int(hn.mix32(0)) == 0 and int(hn.mix32(1)) == 0x86D2FA73
    k = hn.stream_key(3, 1, 2, 5, hn.PURPOSE_NOISE)
    assert int(k) == 0x79CE12EF
    assert [int(v) for v in hn.noise_word(k, np.arange(5))] == [0x2158E4FF, 0xDD7C7C1C, 0xF6B8CFB2, 0x339B8980,
                                                                0x3B5B165B]
    w = hn.noise_word(k, np.arange(8))
    assert w.dtype == np.uint32 and len(set(w.tolist())) == 8
    again = hn.noise_word(hn.stream_key(3, 1, 2, 5, hn.PURPOSE_NOISE), np.arange(8))
    assert np.array_equal(w, again)
    assert not np.array_equal(w, hn.noise_word(hn.stream_key(3, 1, 2, 6, hn.PURPOSE_NOISE), np.arange(8)))


def test_flip_rates_and_independence_within_quad():
    n = 2_000_000
    cells = np.arange(4 * n, dtype=np.uint32)
    key = hn.stream_key(3, 77, 1, 9, hn.PURPOSE_NOISE)
    w = hn.noise_word(key, cells).reshape(n, 4)
    for noise in (0.01, 0.265, 0.375):
        hit = w < hn.flip_threshold(noise)
        p = hit.mean(0)
        assert np.all(np.abs(p - noise) < 5 * np.sqrt(noise * (1 - noise) / n)), (noise, p)
        for a in range(4):
            for b in range(a + 1, 4):
                joint = (hit[:, a] & hit[:, b]).mean()
                sd = np.sqrt(noise * noise * (1 - noise * noise) / n)
                assert abs(joint - noise * noise) < 5 * sd, (noise, a, b, joint)


def test_uniform01_range():
    u = hn.uniform01(hn.cell_hash(hn.stream_key(1, 2, 3, 4, hn.PURPOSE_ACTION), np.arange(1000)))
    assert u.dtype == np.float32 and float(u.min()) >= 0.0 and float(u.max()) < 1.0

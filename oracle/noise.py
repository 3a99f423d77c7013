"""Counter-based hash shared by the oracle and the CUDA kernels (numpy, uint32).

The reference draws measurement noise from torch's global RNG
(mapping/simulations.py:56-58), actions from torch.multinomial
(actor/network.py:94) and message failures from numpy's global RNG
(agent/communication_log.py:46), so RNG parity is impossible by construction
(SURVEY.md section 7).  Both sides therefore take their random bits from this
stateless hash keyed on (seed, episode, agent, index, purpose, cell); it is the
*specification* of ``ipp_marl_b200/csrc/ipp_hash.cuh``.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np

_M1 = np.uint32(0x21F0AAAD)
_M2 = np.uint32(0x735A2D97)
_GOLD = np.uint32(0x9E3779B9)
_CELL = np.uint32(0x9E3779B1)

PURPOSE_NOISE = 0
PURPOSE_ACTION = 1
PURPOSE_COMM = 2


def mix32(x):
    """Bijective 32-bit finaliser (two multiply / xor-shift rounds)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= _M1
        x ^= x >> np.uint32(15)
        x *= _M2
        x ^= x >> np.uint32(15)
    return x


def stream_key(seed, episode, agent, index, purpose):
    """Key of one random stream: (seed, episode, agent, index, purpose) -> u32."""
    with np.errstate(over="ignore"):
        k = mix32(np.uint32(seed & 0xFFFFFFFF) + _GOLD)
        k = mix32(k ^ np.asarray(episode).astype(np.uint32))
        tag = (
            (np.uint32(purpose) << np.uint32(24))
            | (np.asarray(agent).astype(np.uint32) << np.uint32(16))
            | np.asarray(index).astype(np.uint32)
        )
        k = mix32(k + tag)
    return k


def cell_hash(key, cell):
    """u32 hash of element ``cell`` of stream ``key``."""
    with np.errstate(over="ignore"):
        c = np.asarray(cell).astype(np.uint32) * _CELL
    return mix32(np.asarray(key, dtype=np.uint32) ^ c)


_KC = np.array([0x00000000, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F], dtype=np.uint32)
_MC = np.array([0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F], dtype=np.uint32)  # all odd


def noise_word(key, cell):
    """u32 noise word of flat cell index ``cell`` in measurement stream ``key``.

    One strong hash per QUAD (4 consecutive cells, index cell >> 2), then one xor + one odd
    multiply per cell: 4x fewer mixing rounds than hashing every cell, and the four words of a
    quad are statistically independent for threshold tests (checked in tests/test_noise.py).
    """
    cell = np.asarray(cell).astype(np.uint32)
    h = cell_hash(key, cell >> np.uint32(2))
    c = (cell & np.uint32(3)).astype(np.intp)
    with np.errstate(over="ignore"):
        return (h ^ _KC[c]) * _MC[c]


def flip_threshold(noise):
    """A cell is measured wrongly iff noise_word < flip_threshold(noise)."""
    return np.uint32(int(np.floor(float(noise) * 4294967296.0)))


def uniform01(h):
    """24-bit uniform in [0, 1) from a hash word (exact in float32)."""
    return (np.asarray(h, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(
        1.0 / 16777216.0
    )

"""Ground-truth facade (reference: mapping/ground_truths.py:16-176).

The reference computes a Gaussian random field (an O(G^2) Python loop + two FFTs) and then
discards it, returning a half-plane field chosen by ``np.random.seed(episode)`` (:42-56,176).  Only
the returned field and the global-RNG side effect (seed + two randint draws) are reproduced.
"""
import numpy as np


def gaussian_random_field(pk, x_dim: int, y_dim: int, episode: int) -> np.array:
    field = np.zeros((y_dim, x_dim))
    np.random.seed(episode)
    split_idx = np.random.randint(4)
    percentage_idx = np.random.randint(30, 61)
    if split_idx == 0:
        field[: int((y_dim * percentage_idx) / 100), :] = 1
    elif split_idx == 1:
        field[int((y_dim * (1 - percentage_idx)) / 100):, :] = 1
    elif split_idx == 2:
        field[:, : int((x_dim * percentage_idx) / 100)] = 1
    elif split_idx == 3:
        field[:, int((x_dim * (1 - percentage_idx)) / 100):] = 1
    return field

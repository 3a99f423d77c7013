"""ipp_marl_b200 — B200-native batched multi-UAV informative-path-planning environment.

The per-timestep environment path of dmar-bonn/ipp-marl (footprint projection, noisy
measurement, Bayesian occupancy update, local/global map fusion, information-gain reward,
transition + collision masks) as sm_100a CUDA kernels behind a C ABI (include/ipp_b200.h).
"""
from .geometry import HostTables, make_config  # noqa: F401
from .env import BatchedIPPEnv  # noqa: F401

__all__ = ["HostTables", "make_config", "BatchedIPPEnv"]

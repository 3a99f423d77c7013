"""Shared plumbing of the facade: one C-ABI handle per parameter set, numpy <-> pointer helpers,
and the noise-stream context (which counter-based stream the next measurement draws from)."""
import ctypes as C
import json

import numpy as np

from ipp_marl_b200 import _native as N
from ipp_marl_b200.geometry import HostTables, make_config

_handles = {}


class Runtime:
    def __init__(self, params):
        import torch

        if not torch.cuda.is_available():
            raise N.IppError("the ipp_marl_b200 facade needs a CUDA device (no CPU fallback exists)")
        self.lib = N.load()
        self.tables = HostTables(params)
        self.cfg = make_config(self.tables, 1)
        self.h = C.c_void_p()
        N.check(self.lib, None, self.lib.ipp_create(C.byref(self.cfg), C.byref(self.h)), "ipp_create")

    def check(self, rc, what):
        N.check(self.lib, self.h, rc, what)


def runtime(params):
    key = json.dumps(params, sort_keys=True, default=str)
    rt = _handles.get(key)
    if rt is None:
        rt = _handles[key] = Runtime(params)
    return rt


def f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def native(a):
    """C-contiguous copy of ``a`` in the dtype the reference would compute in: float64 stays float64, everything
    else is float32 (measurement arrays and prior maps are float32 in the reference).  -> (array, is_f64)."""
    a = np.asarray(a)
    if a.dtype == np.float64:
        return np.array(a, dtype=np.float64, order="C", copy=True), 1
    return np.array(a, dtype=np.float32, order="C", copy=True), 0


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class NoiseContext:
    """Random stream of the next measurement: (episode, agent, index) -> oracle/noise.py key.

    The reference draws noise from torch's global RNG (mapping/simulations.py:56-58).  Here every
    measurement uses a counter-based stream; callers that know which agent / measurement index they
    are at (the Agent facade, parity tests) set ``agent`` and ``index``; anonymous callers
    (IG_baseline.py, lawn_mower.py call Mapping.update_grid_map directly) get a fresh stream per call
    from the running ``counter`` under the pseudo agent 255.
    """

    agent = None
    index = 0
    counter = 0

    @classmethod
    def next_stream(cls):
        if cls.agent is not None:
            return int(cls.agent), int(cls.index)
        cls.counter = (cls.counter + 1) & 0xFFFF
        return 255, cls.counter

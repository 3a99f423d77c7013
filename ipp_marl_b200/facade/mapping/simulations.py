"""Simulation facade (reference: mapping/simulations.py:16-65): ground truth + noisy measurement."""
import ctypes as C
from typing import Dict

import numpy as np

from ipp_marl_b200.facade import _runtime as R
from mapping import ground_truths
from mapping.grid_maps import GridMap


class Simulation:
    def __init__(self, params: Dict, sensor, episode: int, sensor_model):
        self.params = params
        self.sensor = sensor
        self.cluster_radius = self.params["sensor"]["simulation"]["cluster_radius"]
        self.seed = params["environment"]["seed"]
        self.grid_map = GridMap(self.params)
        self.x_dim_pixel = self.grid_map.x_dim
        self.y_dim_pixel = self.grid_map.y_dim
        self.episode = episode
        self.simulated_map = self.simulate_map(episode)
        self.sensor_model = sensor_model

    def simulate_map(self, episode: int):
        return ground_truths.gaussian_random_field(
            lambda k: k ** (-self.cluster_radius), self.y_dim_pixel, self.x_dim_pixel, episode
        )

    def get_measurement(self, altitude, footprint, mode):
        """footprint = clipped [yu, yd, xl, xr]; returns float32 [xr-xl, yd-yu] (simulations.py:42-51)."""
        rt = R.runtime(self.params)
        yu, yd, xl, xr = (int(v) for v in footprint)
        out = np.empty((max(xr - xl, 0), max(yd - yu, 0)), dtype=np.float32)
        if out.size == 0:
            return out
        noise = self.sensor_model.get_noise_variance(altitude)
        acc = 1 - noise
        y_hi = np.float32(np.round(acc, 3))
        y_lo = np.float32(np.round(1 - acc, 3))
        gt = np.ascontiguousarray(self.simulated_map != 0, dtype=np.uint8)
        rect = (C.c_int32 * 4)(yu, yd, xl, xr)
        agent, index = R.NoiseContext.next_stream()
        rc = rt.lib.ipp_measure(rt.h, R.ptr(gt), rect, int(altitude), int(self.episode) & 0xFFFFFFFF, agent, index,
                                float(y_hi), float(y_lo), R.ptr(out))
        rt.check(rc, "ipp_measure")
        return out

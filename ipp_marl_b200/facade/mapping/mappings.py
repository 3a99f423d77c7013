"""Mapping facade (reference: mapping/mappings.py:19-132): Bayesian occupancy update + map fusion."""
from typing import Dict

import numpy as np

from agent.state_space import AgentStateSpace
from ipp_marl_b200.facade import _runtime as R
from mapping.grid_maps import GridMap
from mapping.simulations import Simulation
from sensors.cameras import Camera
from sensors.models.sensor_models import AltitudeSensorModel


def _fixed_footprint(footprint, clipped):
    """Where the clipped measurement sits inside the raw footprint image (utils/utils.py:79-98)."""
    h, w = footprint[1] - footprint[0], footprint[3] - footprint[2]
    yu, yd, xl, xr = 0, h, 0, w
    if clipped[0] > footprint[0]:
        yu = h - (clipped[1] - clipped[0])
    if clipped[1] < footprint[1]:
        yd = clipped[1] - clipped[0]
    if clipped[3] < footprint[3]:
        xr = clipped[3] - clipped[2]
    if clipped[2] > footprint[2]:
        xl = w - (clipped[3] - clipped[2])
    return int(yu), int(yd), int(xl), int(xr)


class Mapping:
    def __init__(self, grid_map: GridMap, sensor, params: Dict, episode: int):
        self.params = params
        self.grid_map = grid_map
        self.sensor = sensor
        self.sensor_model = AltitudeSensorModel(self.params)
        self.agent_state_space = AgentStateSpace(self.params)
        self.simulation = Simulation(self.params, self.sensor, episode, self.sensor_model)
        self.simulated_map = self.simulation.simulated_map
        self.prior = self.params["mapping"]["prior"]
        from utils import state as _state

        _state.bind_params(self.params)  # get_shannon_entropy has no params argument in the reference

    def update_grid_map(self, position, map_state, t, mode):
        """mappings.py:32-78: mutates AND returns map_state; 5-tuple like the reference."""
        camera = Camera(self.params, self.sensor_model, self.grid_map)
        footprint, clipped = camera.project_field_of_view(
            position, self.grid_map.resolution_x, self.grid_map.resolution_y
        )
        footprint_img = np.ones((footprint[1] - footprint[0], footprint[3] - footprint[2])) * 0.5
        section = map_state[clipped[2]:clipped[3], clipped[0]:clipped[1]]
        measurement = self.simulation.get_measurement(position[2], clipped, mode)
        cell_update = self.update_cells(section, measurement, mode)
        map_state[clipped[2]:clipped[3], clipped[0]:clipped[1]] = cell_update
        map2communicate = np.ones_like(map_state) * 0.5
        map2communicate[clipped[2]:clipped[3], clipped[0]:clipped[1]] = measurement
        fixed = _fixed_footprint(footprint, clipped)
        footprint_img[fixed[2]:fixed[3], fixed[0]:fixed[1]] = measurement
        return map_state, cell_update, clipped, map2communicate, footprint_img

    def fuse_map(self, own_map_state, other_map_states, agent_id, fusion_mode):
        """mappings.py:80-104: never mutates its inputs; successive whole-map passes.  Like the reference under
        numpy >= 2 the result is float64 as soon as one peer was fused (a float32 copy otherwise)."""
        if fusion_mode == "local":
            others = [other_map_states[k]["map2communicate"] for k in other_map_states if k != agent_id]
        elif isinstance(other_map_states, dict):
            others = [other_map_states[k]["map2communicate"] for k in other_map_states]
        else:
            others = list(other_map_states)
        own = np.array(own_map_state, dtype=np.float32, order="C", copy=True)  # np.float32(own.copy())
        if not others:
            return own
        rt = R.runtime(self.params)
        stack = np.ascontiguousarray(np.stack([R.f32c(o) for o in others]), dtype=np.float32)
        out = np.empty(own.shape, dtype=np.float64)
        rc = rt.lib.ipp_fuse_map(rt.h, R.ptr(own), R.ptr(stack), len(others), own.size, R.ptr(out))
        rt.check(rc, "ipp_fuse_map")
        return out

    def update_cells(self, map_section, measurement, mode):
        return self.apply_update(map_section, measurement, mode)

    def apply_update(self, x, y, mode):
        """mappings.py:109-119: x is clamped IN PLACE (in its own dtype); y is an array (float32 measurement) or a
        Python float (IG_baseline.py:240-245: float64 scalar); returns float64 like the reference under numpy >= 2."""
        rt = R.runtime(self.params)
        xc, x64 = R.native(x)
        scalar = np.ndim(y) == 0
        if scalar:
            y64 = 0 if isinstance(y, np.float32) else 1
            yc = np.full(1, y, dtype=np.float64 if y64 else np.float32)
        else:
            yc, y64 = R.native(np.broadcast_to(y, np.shape(x)))
        out = np.empty(xc.shape, dtype=np.float64)
        rc = rt.lib.ipp_update_cells(rt.h, R.ptr(xc), x64, R.ptr(yc), y64, 1 if scalar else 0, xc.size, R.ptr(out))
        rt.check(rc, "ipp_update_cells")
        if isinstance(x, np.ndarray):
            x[...] = xc  # the reference's in-place clamp (mappings.py:110-111)
        return out

    @staticmethod
    def get_update(l_):
        return 1 - (1 / (1 + np.exp(l_)))

    def init_priors(self):
        return np.full((int(self.grid_map.x_dim), int(self.grid_map.y_dim)), self.prior, dtype="float32")

"""Camera facade (reference: sensors/cameras.py:12-79): footprint projection."""
import ctypes as C
from typing import Dict, Tuple

import numpy as np

from ipp_marl_b200.facade import _runtime as R
from mapping.grid_maps import GridMap
from sensors import Sensor


class Camera(Sensor):
    def __init__(self, params: Dict, sensor_model, grid_map: GridMap):
        super().__init__(sensor_model, grid_map)
        self.params = params
        self.grid_map = GridMap(self.params)

    @property
    def angle_x(self) -> float:
        return self.params["sensor"]["field_of_view"]["angle_x"]

    @property
    def angle_y(self) -> float:
        return self.params["sensor"]["field_of_view"]["angle_y"]

    def field_of_view_range(self, height: float) -> Tuple[float, float]:
        return (2 * height * np.tan(0.5 * np.radians(self.angle_x)),
                2 * height * np.tan(0.5 * np.radians(self.angle_y)))

    def project_field_of_view(self, position: np.array, res_x, res_y):
        """-> ([yu, yd, xl, xr] raw, same clipped to the grid): cameras.py:46-79 via the host tables
        (positions must be lattice points, as everywhere in the reference)."""
        rt = R.runtime(self.params)
        pos = (C.c_int32 * 3)(int(position[0]), int(position[1]), int(position[2]))
        raw = (C.c_int32 * 4)()
        clipped = (C.c_int32 * 4)()
        rc = rt.lib.ipp_project_fov(rt.h, pos, raw, clipped)
        rt.check(rc, "ipp_project_fov(%s)" % list(position))
        return list(raw), list(clipped)

    def take_measurement(self, position, verbose: bool = True):
        pass

    def get_resolution_factor(self, position) -> float:
        pass

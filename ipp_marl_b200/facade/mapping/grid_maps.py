"""GridMap facade (reference: mapping/grid_maps.py:8-70)."""
import logging
import math
from typing import Dict

logger = logging.getLogger(__name__)


class GridMap:
    def __init__(self, params: Dict):
        self.params = params
        self.mean = None
        self.resolution_x = self.res_x
        self.resolution_y = self.res_y
        self.occupancy_matrix = None

    def _dim(self, key, res):
        # same error behaviour as grid_maps.py:19-27: ValueError on a missing specification
        if "environment" not in self.params.keys():
            logger.error("Cannot find environment specification in config file!")
            raise ValueError
        if key not in self.params["environment"].keys():
            logger.error("Cannot find environment's %s specification in config file!" % key)
            raise ValueError
        return int(self.params["environment"][key] / res)

    @property
    def x_dim(self) -> int:
        return self._dim("x_dim", self.resolution_x)

    @property
    def y_dim(self) -> int:
        return self._dim("y_dim", self.resolution_y)

    def _res(self, axis):
        alt = self.params["experiment"]["constraints"]["min_altitude"]
        angle = self.params["sensor"]["field_of_view"]["angle_" + axis]
        number = self.params["sensor"]["pixel"]["number_" + axis]
        return (2 * alt * math.tan(math.radians(angle) * 0.5)) / number

    @property
    def res_x(self):
        return self._res("x")

    @property
    def res_y(self):
        return self._res("y")

    @property
    def num_grid_cells(self):
        return self.x_dim * self.y_dim

"""AgentStateSpace facade (reference: agent/state_space.py:9-67): lattice geometry (host integers)."""
from typing import Dict

import numpy as np


class AgentStateSpace:
    def __init__(self, params: Dict):
        self.params = params
        self.seed = params["environment"]["seed"]
        con = params["experiment"]["constraints"]
        self.spacing = con["spacing"]
        self.min_altitude = con["min_altitude"]
        self.max_altitude = con["max_altitude"]
        self.space_x_dim = params["environment"]["x_dim"] // self.spacing + 1
        self.space_y_dim = params["environment"]["y_dim"] // self.spacing + 1
        self.space_z_dim = (self.max_altitude - self.min_altitude) // self.spacing + 1
        self.space_dim = np.array([self.space_x_dim, self.space_y_dim, self.space_z_dim])
        self.class_weighting = params["experiment"]["missions"]["class_weighting"]
        self.planning_uncertainty = params["experiment"]["missions"]["planning_uncertainty"]

    def get_random_agent_state(self, agent_id, episode):
        r = np.random.RandomState(seed=self.seed * episode * agent_id)
        x = self.spacing * r.randint(0, self.space_x_dim)
        y = self.spacing * r.randint(0, self.space_y_dim)
        return np.array([x, y, 15])

    def position_to_index(self, position):
        return np.array([position[0] // self.spacing, position[1] // self.spacing,
                         (position[2] // self.spacing) - 1])

    def index_to_position(self, state):
        return np.array([state[0] * self.spacing, state[1] * self.spacing, self.spacing + state[2] * self.spacing])

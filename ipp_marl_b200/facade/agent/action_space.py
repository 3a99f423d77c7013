"""AgentActionSpace facade (reference: agent/action_space.py:9-589), 6-action space of params.yaml.

Masks are float64 ``np.ones(6)`` arrays mutated in place by apply_collision_mask, as in the
reference (:57,70,331).  The 4/9/27-action variants are not part of the GPU path.
"""
from typing import Dict

import numpy as np

_OFFSETS = {0: (0, 0, 1), 1: (-1, 0, 0), 2: (0, -1, 0), 3: (0, 1, 0), 4: (1, 0, 0), 5: (0, 0, -1)}


class AgentActionSpace:
    def __init__(self, params: Dict):
        self.params = params
        con = params["experiment"]["constraints"]
        self.spacing = con["spacing"]
        self.min_altitude = con["min_altitude"]
        self.max_altitude = con["max_altitude"]
        self.space_x_dim = 3
        self.space_y_dim = 3
        self.space_z_dim = (self.max_altitude - self.min_altitude) // self.spacing + 1
        self.num_actions = con["num_actions"]
        if self.num_actions != 6:
            raise NotImplementedError("only num_actions == 6 is implemented (DESIGN.md section 7)")
        self.environment_x_dim = params["environment"]["x_dim"]
        self.environment_y_dim = params["environment"]["y_dim"]
        self.space_dim = np.array([self.space_x_dim, self.space_y_dim, self.space_z_dim])

    def get_action_mask(self, position):
        mask = np.ones(6)
        if position[2] == self.max_altitude:
            mask[0] = 0
        if position[2] == self.min_altitude:
            mask[5] = 0
        if position[1] == 0:
            mask[2] = 0
        if position[1] == self.environment_y_dim:
            mask[3] = 0
        if position[0] == 0:
            mask[1] = 0
        if position[0] == self.environment_x_dim:
            mask[4] = 0
        return mask, mask

    def action_to_position(self, position, action_index: int):
        off = _OFFSETS.get(int(action_index), (0, 0, 0))
        return position + np.array([self.spacing * off[0], self.spacing * off[1], self.spacing * off[2]])

    def apply_collision_mask(self, position, mask, next_other_positions, agent_state_space):
        for other in next_other_positions:
            rel = agent_state_space.position_to_index(other) - agent_state_space.position_to_index(position)
            rules = (((0, 0), (0, 5)), ((-1, 0), (1,)), ((0, -1), (2,)), ((0, 1), (3,)), ((1, 0), (4,)))
            for (dx, dy), idxs in rules:
                if rel[0] == dx and rel[1] == dy and np.sum(mask) > 1:
                    for i in idxs:
                        mask[i] = 0
        return mask

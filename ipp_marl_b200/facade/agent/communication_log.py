"""CommunicationLog facade (reference: agent/communication_log.py:12-65)."""
from typing import Dict

import numpy as np


class CommunicationLog:
    def __init__(self, params: Dict, num_episode: int):
        self.params = params
        uav = self.params["experiment"]["uav"]
        self.communication_range = uav["communication_range"]
        self.fix_range = uav["fix_range"]
        self.failure_rate = uav["failure_rate"]
        self.n_agents = self.params["experiment"]["missions"]["n_agents"]
        self.global_log = dict()
        if not self.fix_range:
            np.random.seed(num_episode)
            self.communication_range = {0: 0, 1: 15, 2: 25, 3: 100}[int(np.random.randint(4))]

    def store_agent_message(self, message: Dict, agent_id: int):
        self.global_log[agent_id] = message
        return self.global_log

    def get_messages(self, agent_id: int):
        own = self.global_log[agent_id]["position"]
        local_log = dict()
        for other_id in self.global_log.keys():
            other = self.global_log[other_id]["position"]
            r = np.random.random_sample()  # one draw per ordered pair, used or not (:46)
            d = np.linalg.norm(np.asarray(own) - np.asarray(other), ord=2)
            if d < 0.001 or (0.001 <= d <= self.communication_range and r >= self.failure_rate):
                local_log[other_id] = self.global_log[other_id]
        return local_log

    def get_global_positions(self):
        return [[self.global_log[a]["position"]] for a in self.global_log]

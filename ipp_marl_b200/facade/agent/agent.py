"""Agent facade (reference: agent/agent.py:13-117): per-UAV orchestration of mask / act / move /
measure; same attributes (position, local_map, map_footprint, map2communicate, footprint_img)."""
from typing import Dict

from agent.action_space import AgentActionSpace
from ipp_marl_b200.facade._runtime import NoiseContext


class Agent:
    def __init__(self, actor_network, params: Dict, mapping, agent_id: int, agent_state_space):
        self.params = params
        self.agent_id = agent_id
        self.mission_type = self.params["experiment"]["missions"]["type"]
        self.n_actions = self.params["experiment"]["constraints"]["num_actions"]
        self.v_max = self.params["experiment"]["uav"]["max_v"]
        self.a_max = self.params["experiment"]["uav"]["max_a"]
        self.x_dim = params["environment"]["x_dim"]
        self.y_dim = params["environment"]["y_dim"]
        self.mapping = mapping
        self.local_map = mapping.init_priors()
        self.agent_state_space = agent_state_space
        self.action_space = AgentActionSpace(self.params)
        self.actor_network = actor_network
        self.agent_info = dict()
        self.position = None
        self.map_footprint = None
        self.map2communicate = None
        self.footprint_img = None

    def _measure(self, t_index, t, mode):
        NoiseContext.agent, NoiseContext.index = self.agent_id, t_index
        try:
            out = self.mapping.update_grid_map(self.position, self.local_map, t, mode)
        finally:
            NoiseContext.agent = None
        return out

    def communicate(self, t, num_episode, communication_log, mode):
        if t == 0:
            self.position = self.agent_state_space.get_random_agent_state(self.agent_id, num_episode)
            (self.local_map, self.map_footprint, _, self.map2communicate, self.footprint_img) = self._measure(0, t, mode)
        agent_info = {
            "local_map": self.local_map,
            "position": self.position,
            "map_footprint": self.map_footprint,
            "map2communicate": self.map2communicate,
            "footprint_img": self.footprint_img,
        }
        global_log = communication_log.store_agent_message(agent_info, self.agent_id)
        return global_log, self.local_map, self.position

    def receive_messages(self, communication_log, agent_id, t):
        received = communication_log.get_messages(self.agent_id)
        if len(received) > 0:
            self.local_map = self.mapping.fuse_map(self.local_map, received, agent_id, "local")
        return received, self.local_map

    def step(self, agent_id, t, num_episode, batch_memory, mode, next_other_positions):
        mask, _ = self.action_space.get_action_mask(self.position)
        mask = self.action_space.apply_collision_mask(self.position, mask, next_other_positions,
                                                      self.agent_state_space)
        probs, action, mask_out, eps = self.actor_network.get_action_index(
            batch_memory, mask, self.agent_id, t, num_episode, mode
        )
        self.position = self.action_space.action_to_position(self.position, action)
        if not self.is_in_map(self.position):
            print("OUT OF MAP")
        (self.local_map, self.map_footprint, footprint_idx, self.map2communicate,
         self.footprint_img) = self._measure(t + 1, t, mode)
        batch_memory.insert(-1, agent_id, action=action, mask=mask_out)
        return self.local_map, self.position, eps, action, footprint_idx, self.map2communicate

    def is_in_map(self, position):
        return bool(0 <= position[0] <= self.x_dim and 0 <= position[1] <= self.y_dim and 5 <= position[2] <= 15)

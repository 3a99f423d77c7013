"""Run one episode through the drop-in facade classes in the reference's call order
(coma_wrapper.py:37-183 / missions/episode_generator.py:38-56) and dump what the golden fixture holds.

Executed as a subprocess by tests/test_facade*.py (the facade rewires sys.path / sys.modules).
usage: python tests/facade_episode.py <golden.npz> <out.npz> [--host-only]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    golden, out = sys.argv[1], sys.argv[2]
    host_only = "--host-only" in sys.argv
    z = np.load(golden, allow_pickle=False)
    params = json.loads(str(z["params_json"]))
    episode = int(z["episode"])
    from ipp_marl_b200 import facade

    facade.install(None)
    from oracle import noise as hn

    from mapping.grid_maps import GridMap
    from mapping.mappings import Mapping
    from sensors import Sensor
    from sensors.models.sensor_models import AltitudeSensorModel
    from agent.agent import Agent
    from agent.state_space import AgentStateSpace
    from agent.communication_log import CommunicationLog
    from utils.reward import get_global_reward
    import marl_framework.mapping.mappings as twin

    assert twin.Mapping is Mapping, "marl_framework.* twin must be the same module object"
    seed = params["environment"]["seed"]
    A = params["experiment"]["missions"]["n_agents"]
    budget = params["experiment"]["constraints"]["budget"]
    grid_map = GridMap(params)
    ass = AgentStateSpace(params)
    res = {"gx": grid_map.x_dim, "gy": grid_map.y_dim, "lattice": ass.space_dim}
    if host_only:
        # host logic only (no GPU): geometry, start positions, masks, ground truth
        from agent.action_space import AgentActionSpace
        from mapping import ground_truths

        aas = AgentActionSpace(params)
        res["start"] = np.array([ass.get_random_agent_state(a, episode) for a in range(A)])
        res["gt"] = ground_truths.gaussian_random_field(None, grid_map.y_dim, grid_map.x_dim, episode).astype(np.uint8)
        masks = []
        for t in range(len(z["pos"])):
            moved = []
            row = []
            for a in range(A):
                m, _ = aas.get_action_mask(z["pos"][t][a])
                m = aas.apply_collision_mask(z["pos"][t][a], m, moved, ass)
                moved.append(aas.action_to_position(z["pos"][t][a], int(z["action"][t][a])))
                row.append(m)
            masks.append(row)
            assert np.array_equal(np.array(moved), z["pos_next"][t])
        res["mask"] = np.array(masks)
        np.savez(out, **res)
        return

    class Actor:
        def get_action_index(self, batch_memory, mask, agent_id, t, num_episode, mode):
            self.mask = np.array(mask, dtype=np.float64).copy()
            return None, self.action, mask, 0.0

    class Memory:
        def insert(self, *a, **k):
            pass

    actor, memory = Actor(), Memory()
    sensor = Sensor(AltitudeSensorModel(params), grid_map)
    mapping = Mapping(grid_map, sensor, params, episode)
    agents = [Agent(actor, params, mapping, a, ass) for a in range(A)]
    global_map = agents[0].local_map.copy()
    T = len(z["reward_rel"])
    rec = {k: [] for k in ("pos", "comm", "mask", "pos_next", "reward_rel", "reward_abs", "global", "local_fused",
                           "local_after_move")}
    for t in range(T):
        log = CommunicationLog(params, episode)
        info = {}
        for a in range(A):
            info, _, _ = agents[a].communicate(t, episode, log, None)
        rec["pos"].append(np.array([agents[a].position for a in range(A)]))
        comm = np.zeros((A, A), dtype=np.uint8)
        fused = []
        for a in range(A):
            key = hn.stream_key(seed, episode, a, t, hn.PURPOSE_COMM)
            draws = iter(float(hn.uniform01(hn.cell_hash(key, j))) for j in range(A))
            saved = np.random.random_sample
            np.random.random_sample = lambda *x, **y: next(draws)
            try:
                received, fm = agents[a].receive_messages(log, a, t)
            finally:
                np.random.random_sample = saved
            for j in received:
                comm[a, j] = 1
            fused.append(np.array(fm, dtype=np.float64))
        rec["comm"].append(comm)
        rec["local_fused"].append(np.array(fused))
        next_global = mapping.fuse_map(global_map, info, None, "global")
        moved, masks = [], []
        for a in range(A):
            actor.action = int(z["action"][t][a])
            _, pos, _, _, _, _ = agents[a].step(a, t, episode, memory, None, moved)
            moved.append(pos)
            masks.append(actor.mask)
        _, rel, ab = get_global_reward(global_map, next_global, "COMA", None, mapping.simulated_map, ass, None, None,
                                       t, budget)
        rec["mask"].append(np.array(masks))
        rec["pos_next"].append(np.array(moved))
        rec["reward_rel"].append(rel)
        rec["reward_abs"].append(ab)
        rec["global"].append(np.array(next_global, dtype=np.float64))
        rec["local_after_move"].append(np.array([np.array(agents[a].local_map, dtype=np.float64) for a in range(A)]))
        global_map = next_global
    res["gt"] = np.array(mapping.simulated_map != 0, dtype=np.uint8)
    for k, v in rec.items():
        res[k] = np.array(v)
    np.savez(out, **res)


if __name__ == "__main__":
    main()

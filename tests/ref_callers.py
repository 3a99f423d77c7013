"""Run the reference's UNCHANGED caller files — ``missions/episode_generator.py`` + ``coma_wrapper.py``
(``COMAWrapper.build_observations`` / ``.steps``, coma_wrapper.py:37-183), ``IG_baseline.py`` (``IG_baseline.execute``,
:56-220) and ``lawn_mower.py`` (``LawnMower.execute``, :38-315) — either on the reference's own environment modules
(``ref``) or on top of ``ipp_marl_b200.facade`` (``facade``: every ``mapping.* / sensors.* / agent.* / utils.reward /
utils.state`` import of those files resolves to the CUDA-backed modules), and dump what they compute.

Executed as a subprocess (the facade rewires sys.path / sys.modules) by tests/test_ref_callers.py, which compares the
two dumps with each other and with the committed golden fixtures.  The caller files come from ``baseline/_ref`` (the
verbatim install of scripts/install_ref.py) or /root/reference and are not touched; what is injected — identically in
both modes, through attribute assignment on the live objects — is exactly the randomness the two sides cannot share
otherwise (SURVEY.md section 7 "RNG parity is impossible by construction"):
  * the stream of every measurement: ``Mapping.update_grid_map`` is wrapped to set (agent, index) = (n mod S, n div S)
    for its n-th call (S = n_agents; 8 for the lawn mower's eight chained updates);
  * the message-failure draws: ``Agent.receive_messages`` runs with ``np.random.random_sample`` replaced by the hash
    uniform of (seed, episode, agent, t);
  * the policy: ``COMAWrapper.actor_network.get_action_index`` returns the hash-uniform masked action
    (``coma`` mode; IG_baseline and LawnMower choose their own actions).

usage: python tests/ref_callers.py {coma|ig|lawn} <params.json> <episode> <out.npz> {ref|facade}
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, params_path, episode, out, side = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    with open(params_path) as f:
        params = json.load(f)
    from oracle import noise as hn
    from oracle import ref_harness as rh

    root = rh.reference_root()
    if not os.path.isdir(os.path.join(root, "marl_framework")):
        raise SystemExit("no reference tree (run scripts/install_ref.py in the build container)")
    seed = params["environment"]["seed"]
    A = params["experiment"]["missions"]["n_agents"]

    if side == "facade":
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d",
                     "seaborn", "cma"):
            sys.modules.setdefault(name, rh._Stub(name))
        from ipp_marl_b200 import facade

        facade.install(root)
        from ipp_marl_b200.facade._runtime import NoiseContext as NC
        import mapping.mappings as m1

        mapping_modules = (m1,)
    else:
        ns = rh.load()
        rh.install_noise_patch()
        NC = rh.NoiseContext
        NC.seed, NC.episode, NC.noiseless = seed, episode, False
        mapping_modules = ns.mapping_modules

    # ---- measurement streams by call count (class-level wrapper: the callers build their own Mapping objects) ----
    n_streams = 8 if mode == "lawn" else A
    calls = {"n": 0, "seen": None}  # seen: union of all footprints so far (for the F1 knife-edge bounds)
    for mod in mapping_modules:
        real = mod.Mapping.update_grid_map

        def update_grid_map(self, *a, _real=real, **k):
            NC.agent, NC.index = calls["n"] % n_streams, calls["n"] // n_streams
            calls["n"] += 1
            out = _real(self, *a, **k)
            yu, yd, xl, xr = (int(v) for v in out[2])
            if calls["seen"] is None:
                calls["seen"] = np.zeros(np.shape(out[0]), dtype=bool)
            calls["seen"][xl:xr, yu:yd] = True
            return out

        mod.Mapping.update_grid_map = update_grid_map

    # ---- message-failure draws ----
    import agent.agent as bare_agent
    import marl_framework.agent.agent as mf_agent

    for mod in {id(bare_agent): bare_agent, id(mf_agent): mf_agent}.values():
        real_recv = mod.Agent.receive_messages

        def receive_messages(self, log, agent_id, t, _real=real_recv):
            with rh._comm_draws(seed, episode, agent_id, t):
                return _real(self, log, agent_id, t)

        mod.Agent.receive_messages = receive_messages

    writer = rh._Stub("writer")
    res = {}
    f1_lo, f1_hi = [], []

    def wrap_wrmse(mod):
        """Record, next to every F1 the caller computes (utils/utils.py:43-76), the range it can take when the cells
        within 2e-5 of the 0.5 threshold that some footprint has reached are counted either way
        (tests/helpers.py::f1_bounds)."""
        from tests.helpers import f1_bounds

        real = mod.get_wrmse

        def get_wrmse(map_state, map_simulation):
            lo, hi = f1_bounds(map_state, map_simulation, seen=calls["seen"])
            if os.environ.get("IPP_DUMP_WRMSE"):
                res.setdefault("_maps", []).append(np.array(map_state, dtype=np.float64))
            f1_lo.append(lo)
            f1_hi.append(hi)
            return real(map_state, map_simulation)

        mod.get_wrmse = get_wrmse

    if mode == "coma":
        from marl_framework.batch_memory import BatchMemory
        from marl_framework.coma_wrapper import COMAWrapper
        from marl_framework.mapping.grid_maps import GridMap
        from marl_framework.missions.episode_generator import EpisodeGenerator
        from marl_framework.sensors import Sensor
        from marl_framework.sensors.models.sensor_models import AltitudeSensorModel

        import torch

        wrapper = COMAWrapper(params, writer)

        class Policy:  # replaces the actor CNN (actor/network.py:41-68): hash-uniform masked action
            def get_action_index(self, batch_memory, mask, agent_id, t, num_episode, mode_):
                m = np.array(mask, dtype=np.float64)
                return None, torch.tensor(rh.uniform_policy_action(m, seed, num_episode, agent_id, t)), mask, 0.0

        wrapper.actor_network = Policy()
        globals_, rewards = [], []
        real_steps = wrapper.steps

        def steps(*a, **k):
            r = real_steps(*a, **k)
            rewards.append((float(r[1]), float(r[2])))
            globals_.append(np.array(r[8], dtype=np.float64))
            return r

        wrapper.steps = steps
        grid_map = GridMap(params)
        sensor = Sensor(AltitudeSensorModel(params), grid_map)
        gen = EpisodeGenerator(params, writer, grid_map, sensor)
        memory = BatchMemory(params, wrapper)
        ret = gen.execute(episode, memory, wrapper, "train")
        T = ret[6] + 1
        res["episode_return"] = np.array(float(ret[0]))
        res["pos"] = np.array([[np.array(p) for p in step] for step in ret[5]], dtype=np.int64)  # [T+1, A, 3]
        res["action"] = np.array([[int(x) for x in step] for step in ret[8]], dtype=np.int64)
        res["reward_rel"] = np.array([r[0] for r in rewards])
        res["reward_abs"] = np.array([r[1] for r in rewards])
        res["global"] = np.array(globals_)
        res["gt"] = np.array(np.asarray(ret[3]) != 0, dtype=np.uint8)
        res["obs"] = np.array([[memory.get(t, a, "observation").numpy() for a in range(A)] for t in range(T)])
        res["state"] = np.array([[memory.get(t, a, "state").numpy() for a in range(A)] for t in range(T)])
    elif mode == "ig":
        import IG_baseline as ig_mod

        wrap_wrmse(ig_mod)
        base = ig_mod.IG_baseline(params, writer, episode)
        rec = {"action": [], "gains": []}
        real_ind, real_sel = base.get_individual_ig, base.select_action

        def get_individual_ig(position, mask, map_state):
            o = real_ind(position, mask, map_state)
            rec["gains"].append([float(v) for v in o[1]])
            return o

        def select_action(util):
            a = real_sel(util)
            rec["action"].append(int(a))
            return a

        base.get_individual_ig, base.select_action = get_individual_ig, select_action
        rel_sum, abs_sum, altitudes, entropies, f1s = base.execute()
        res["action"] = np.array(rec["action"], dtype=np.int64).reshape(-1, A)
        res["gains"] = np.array(rec["gains"]).reshape(-1, A, 6)
        res["entropy"] = np.array([float(v) for v in entropies])
        res["f1"] = np.array([float(v) for v in f1s])
        res["reward_rel_sum"] = np.array(float(rel_sum))
        res["reward_abs_sum"] = np.array(float(abs_sum))
    elif mode == "lawn":
        import lawn_mower as lm_mod

        wrap_wrmse(lm_mod)
        lm = lm_mod.LawnMower(params, writer, episode)
        _, entropies, f1s = lm.execute()
        res["entropy"] = np.array([float(v) for v in entropies])
        res["f1"] = np.array([float(v) for v in f1s])
        res["map"] = np.array(lm.map, dtype=np.float64)
    else:
        raise SystemExit("unknown mode " + mode)
    if "_maps" in res:
        res["wrmse_maps"] = np.array(res.pop("_maps"))
    res["update_calls"] = np.array(calls["n"])
    res["f1_lo"], res["f1_hi"] = np.array(f1_lo), np.array(f1_hi)
    np.savez(out, **res)


if __name__ == "__main__":
    main()

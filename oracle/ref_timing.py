"""Time the UNMODIFIED reference on the host cores (bench.py's `--impl reference` arm and `cpu_baseline` leg).

TEST / MEASUREMENT INFRASTRUCTURE (never imported by the product package).  Nothing of the reference is patched here
except the plotting modules that this image lacks (matplotlib / seaborn / cma are stubbed in ``sys.modules``): the
measurement noise comes from ``torch.multinomial`` on the global RNG exactly as in mapping/simulations.py:53-65 and the
message-failure draws from ``np.random.random_sample`` (agent/communication_log.py:46).

Two workloads:
  * ``env_loop``: the reference's per-timestep environment path of SURVEY.md section 3.2, built from the reference's own
    objects in coma_wrapper's call order — ``Mapping(...)`` (ground truth), ``Agent.communicate`` /
    ``receive_messages`` (comm-range local fusion), ``Mapping.fuse_map(... "global")``, ``Agent.step`` (masks, move,
    ``update_grid_map``) and ``get_global_reward`` — with a uniform random masked policy instead of the actor CNN.
    One env-step = one timestep in which all A UAVs act.
  * ``episode_generator``: ``EpisodeGenerator.execute`` with ``COMAWrapper`` (missions/episode_generator.py:38-88),
    i.e. the same loop plus the observation / critic-state builders and the actor CNN forward on the CPU.
"""
import os
import time

import numpy as np

EP_LEN_KEY = ("experiment", "constraints", "budget")


class _RandomActor:
    """Stands in for ActorNetwork in the env-only loop: uniform over the unmasked actions (agent/agent.py:82-84 only
    calls ``get_action_index``)."""

    def get_action_index(self, batch_memory, action_mask_1d, agent_id, t, num_episode, mode):
        valid = np.flatnonzero(np.asarray(action_mask_1d) > 0)
        act = int(valid[np.random.randint(valid.size)]) if valid.size else -1
        return None, act, action_mask_1d, 0.0


class _NullMemory:
    def insert(self, *a, **k):
        pass

    def add(self, *a, **k):
        pass


def _load():
    from . import ref_harness as rh

    return rh.load()


def env_loop_episode(ns, params, episode):
    """One episode of the env-only loop; returns (steps, seconds spent in Mapping.__init__ = ground truth)."""
    n_agents = params["experiment"]["missions"]["n_agents"]
    budget = params["experiment"]["constraints"]["budget"]
    g0 = time.perf_counter()
    grid_map = ns.GridMap(params)
    sensor = ns.Sensor(ns.AltitudeSensorModel(params), grid_map)
    mapping = ns.Mapping(grid_map, sensor, params, episode)
    gen = time.perf_counter() - g0
    ass = ns.AgentStateSpace(params)
    actor, memory = _RandomActor(), _NullMemory()
    agents = [ns.Agent(actor, params, mapping, a, ass) for a in range(n_agents)]
    global_map = agents[0].local_map.copy()
    for t in range(budget + 1):
        log = ns.CommunicationLog(params, episode)
        info = {}
        for a in range(n_agents):
            info, _, _ = agents[a].communicate(t, episode, log, None)
        for a in range(n_agents):
            agents[a].receive_messages(log, a, t)
        next_global = mapping.fuse_map(global_map, info, None, "global")  # coma_wrapper.py:93-95 (once, not twice)
        moved, acts = [], []
        for a in range(n_agents):
            _, pos, _, action, _, _ = agents[a].step(a, t, episode, memory, None, moved)
            moved.append(pos)
            acts.append(action)
        ns.get_global_reward(global_map, next_global, "COMA", None, mapping.simulated_map, ass, acts, None, t, budget)
        global_map = next_global
    return budget + 1, gen


def env_loop_worker(args):
    """Pool worker: run env-loop episodes for ``seconds``; -> (env_steps, wall_s, ground_truth_s)."""
    params, first_episode, seconds = args
    import torch

    torch.set_num_threads(1)  # one single-env reference process per core
    ns = _load()
    t0 = time.perf_counter()
    steps, gen, ep = 0, 0.0, first_episode
    while True:
        s, g = env_loop_episode(ns, params, ep)
        steps += s
        gen += g
        ep += 1
        if time.perf_counter() - t0 >= seconds:
            break
    return steps, time.perf_counter() - t0, gen


def episode_generator_worker(args):
    """Pool worker: ``EpisodeGenerator.execute`` episodes for ``seconds``; -> (env_steps, wall_s, 0.0)."""
    params, first_episode, seconds = args
    import torch

    torch.set_num_threads(1)
    ns = _load()
    from . import ref_harness as rh
    from marl_framework.batch_memory import BatchMemory
    from marl_framework.coma_wrapper import COMAWrapper
    from marl_framework.missions.episode_generator import EpisodeGenerator

    writer = rh._Stub("writer")
    wrapper = COMAWrapper(params, writer)
    grid_map = ns.GridMap(params)
    sensor = ns.Sensor(ns.AltitudeSensorModel(params), grid_map)
    gen = EpisodeGenerator(params, writer, grid_map, sensor)
    t0 = time.perf_counter()
    steps, ep = 0, first_episode
    while True:
        memory = BatchMemory(params, wrapper)
        out = gen.execute(ep, memory, wrapper, "train")
        steps += out[6] + 1
        ep += 1
        if time.perf_counter() - t0 >= seconds:
            break
    return steps, time.perf_counter() - t0, 0.0


def run(params, seconds, procs, workload="env_loop", pool=None):
    """env-steps/s of ``procs`` independent single-env reference processes (the reference is single-threaded numpy,
    so "all cores" = one process per core)."""
    import multiprocessing as mp

    fn = env_loop_worker if workload == "env_loop" else episode_generator_worker
    jobs = [(params, 1 + 1000 * i, seconds) for i in range(procs)]
    t0 = time.perf_counter()
    if pool is None:
        with mp.get_context("fork").Pool(procs) as own:
            res = own.map(fn, jobs, chunksize=1)
    else:
        res = pool.map(fn, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return {
        "value": sum(r[0] / r[1] for r in res),
        "steps": sum(r[0] for r in res),
        "wall_s": wall,
        "best_single_process": max(r[0] / r[1] for r in res),
        "ground_truth_share": sum(r[2] for r in res) / max(sum(r[1] for r in res), 1e-9),
    }


def available():
    from . import ref_harness as rh

    return rh.available()


def versions():
    import numpy
    import torch

    return {"numpy": numpy.__version__, "torch": torch.__version__, "cores": os.cpu_count()}

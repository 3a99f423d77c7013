"""Drive the UNMODIFIED reference (dmar-bonn/ipp-marl).

TEST INFRASTRUCTURE.  The reference tree is looked up, in this order, at ``$IPP_REFERENCE_ROOT``,
``baseline/_ref`` (the verbatim install made by ``scripts/install_ref.py`` during ``build()``; git-ignored,
it travels to the GPU box with the snapshot) and ``/root/reference`` (build container only).  Used by
``oracle/make_golden.py`` to write the fixtures in ``tests/golden/``, by the tests that re-check the numpy
restatement / the CUDA path against the live reference, and by ``bench.py``'s CPU-baseline legs
(``oracle/ref_timing.py``).  Never imported by the product package.

What is patched (and why) — nothing else of the reference is touched:
  * ``matplotlib`` / ``seaborn`` / ``cma`` are stubbed in ``sys.modules`` (absent
    in this image; imported at module top by mapping/mappings.py:12,
    utils/reward.py:2, utils/state.py:11, utils/utils.py:6 ...).
  * ``Simulation.get_measurement`` (mapping/simulations.py:42-51) keeps its body
    but takes the per-cell "measured correctly" bits from ``oracle.noise``
    instead of ``torch.multinomial`` on the global RNG (simulations.py:56-58),
    so CPU and GPU share noise bits.  Both module aliases are patched
    (SURVEY.md section 8c "harness trap").
  * during ``CommunicationLog.get_messages`` the global ``np.random.random_sample``
    (agent/communication_log.py:46) is swapped for the hash uniform, one draw per
    ordered agent pair, so message failures are reproducible on the GPU.
  * the actor network is replaced by a stub that returns injected actions
    (agent/agent.py:82-84 only needs ``get_action_index``).
"""
import copy
import os
import sys
import types

import numpy as np

from . import noise as hn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    env = os.environ.get("IPP_REFERENCE_ROOT")
    if env:
        return env
    local = os.path.join(_REPO, "baseline", "_ref")
    if os.path.isdir(os.path.join(local, "marl_framework")):
        return local
    return "/root/reference"


REFERENCE_ROOT = reference_root()
_FRAMEWORK = os.path.join(REFERENCE_ROOT, "marl_framework")


def available():
    return os.path.isdir(_FRAMEWORK)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Stub("call")


_loaded = {}


def load():
    """Import the reference modules; returns a namespace of the ones we drive."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.cm",
        "mpl_toolkits",
        "mpl_toolkits.mplot3d",
        "seaborn",
        "cma",
    ):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    for p in (REFERENCE_ROOT, _FRAMEWORK):
        if p not in sys.path:
            sys.path.insert(0, p)

    import marl_framework.mapping.mappings as mf_mappings  # noqa: E402
    import marl_framework.mapping.simulations as mf_sim  # noqa: E402
    import mapping.simulations as bare_sim  # noqa: E402
    import mapping.mappings as bare_mappings  # noqa: E402
    import marl_framework.mapping.grid_maps as mf_grid  # noqa: E402
    import marl_framework.sensors as mf_sensors  # noqa: E402
    import marl_framework.sensors.cameras as mf_cameras  # noqa: E402
    import marl_framework.sensors.models.sensor_models as mf_sm  # noqa: E402
    import marl_framework.agent.agent as mf_agent  # noqa: E402
    import marl_framework.agent.state_space as mf_ss  # noqa: E402
    import marl_framework.agent.action_space as mf_as  # noqa: E402
    import marl_framework.agent.communication_log as mf_cl  # noqa: E402
    import marl_framework.utils.reward as mf_reward  # noqa: E402
    import marl_framework.utils.state as mf_state  # noqa: E402
    import marl_framework.params as mf_params  # noqa: E402

    ns = types.SimpleNamespace(
        Mapping=mf_mappings.Mapping,
        GridMap=mf_grid.GridMap,
        Sensor=mf_sensors.Sensor,
        Camera=mf_cameras.Camera,
        AltitudeSensorModel=mf_sm.AltitudeSensorModel,
        Agent=mf_agent.Agent,
        AgentStateSpace=mf_ss.AgentStateSpace,
        AgentActionSpace=mf_as.AgentActionSpace,
        CommunicationLog=mf_cl.CommunicationLog,
        get_global_reward=mf_reward.get_global_reward,
        get_shannon_entropy=mf_state.get_shannon_entropy,
        load_params=mf_params.load_params,
        sim_modules=(mf_sim, bare_sim),
        mapping_modules=(mf_mappings, bare_mappings),
    )
    _loaded["ns"] = ns
    return ns


def default_params():
    ns = load()
    return ns.load_params(os.path.join(_FRAMEWORK, "params.yaml"))


def synthetic_params(x_dim=50, n_agents=4, comm_range=25, failure_rate=0, seed=3, budget=14, prior=0.5,
                     fix_range=True):
    """SURVEY.md section 8d synthetic family: FoV 90/90, 10x10 px -> 1 cell = 1 m."""
    p = copy.deepcopy(default_params())
    p["environment"]["x_dim"] = x_dim
    p["environment"]["y_dim"] = x_dim
    p["environment"]["seed"] = seed
    p["sensor"]["field_of_view"]["angle_x"] = 90
    p["sensor"]["field_of_view"]["angle_y"] = 90
    p["sensor"]["pixel"]["number_x"] = 10
    p["sensor"]["pixel"]["number_y"] = 10
    p["experiment"]["missions"]["n_agents"] = n_agents
    p["experiment"]["uav"]["communication_range"] = comm_range
    p["experiment"]["uav"]["failure_rate"] = failure_rate
    p["experiment"]["uav"]["fix_range"] = fix_range
    p["experiment"]["constraints"]["budget"] = budget
    p["mapping"]["prior"] = prior
    return p


# ----------------------------------------------------------------------------------------------
# noise injection
# ----------------------------------------------------------------------------------------------
class NoiseContext:
    """Which random stream the next reference measurement draws from."""

    seed = 0
    episode = 0
    agent = 0
    index = 0
    noiseless = False


def _patched_get_measurement(self, altitude, footprint, mode):
    # Same data flow as mapping/simulations.py:42-65; only the source of the
    # "correctness" bits differs (hash instead of torch.multinomial).
    section = self.simulated_map[footprint[2] : footprint[3], footprint[0] : footprint[1]].copy()
    sensor_noise = self.sensor_model.get_noise_variance(altitude)
    gy = self.simulated_map.shape[1]
    xs = np.arange(footprint[2], footprint[3], dtype=np.int64)[:, None]
    ys = np.arange(footprint[0], footprint[1], dtype=np.int64)[None, :]
    cells = xs * gy + ys
    if NoiseContext.noiseless:
        correctness = np.ones(section.shape, dtype=np.int64)
    else:
        key = hn.stream_key(
            NoiseContext.seed, NoiseContext.episode, NoiseContext.agent, NoiseContext.index, hn.PURPOSE_NOISE
        )
        h = hn.noise_word(key, cells)
        correctness = (h >= hn.flip_threshold(sensor_noise)).astype(np.int64)
    accuracy = 1 - sensor_noise
    value = section.copy()
    value = np.where(correctness == 0, abs(value - 1), value)
    value = accuracy * value
    np.putmask(value, (1 - accuracy) > value, 1 - accuracy)
    return np.float32(np.round(value, 3))


def install_noise_patch():
    ns = load()
    for mod in ns.sim_modules:
        mod.Simulation.get_measurement = _patched_get_measurement


class _comm_draws:
    """Replace the one ``np.random.random_sample()`` per ordered pair of
    agent/communication_log.py:46 by the hash uniform of (seed, episode, i, t, j)."""

    def __init__(self, seed, episode, agent, t):
        self.key = hn.stream_key(seed, episode, agent, t, hn.PURPOSE_COMM)
        self.j = 0

    def _draw(self, *a, **k):
        r = float(hn.uniform01(hn.cell_hash(self.key, self.j)))
        self.j += 1
        return r

    def __enter__(self):
        self.saved = np.random.random_sample
        np.random.random_sample = self._draw

    def __exit__(self, *exc):
        np.random.random_sample = self.saved
        return False


# ----------------------------------------------------------------------------------------------
# env-only episode driven from reference calls, in the order of coma_wrapper.py:37-183
# ----------------------------------------------------------------------------------------------
class _StubActor:
    def __init__(self):
        self.next_action = 0
        self.last_mask = None

    def get_action_index(self, batch_memory, action_mask_1d, agent_id, t, num_episode, mode):
        self.last_mask = np.array(action_mask_1d, dtype=np.float64).copy()
        act = self.policy(self.last_mask, agent_id, t)
        return None, act, action_mask_1d, 0.0


class _StubMemory:
    def insert(self, *a, **k):
        pass

    def add(self, *a, **k):
        pass


def uniform_policy_action(mask, seed, episode, agent, t):
    """Uniform over unmasked actions, keyed (seed, episode, agent, t): SURVEY.md section 8d.

    All-zero mask (SURVEY.md section 8a10 corner case; the reference raises in
    torch.multinomial): defined as "stay in place", action -1.
    """
    valid = np.flatnonzero(np.asarray(mask) > 0)
    if valid.size == 0:
        return -1
    key = hn.stream_key(seed, episode, agent, t, hn.PURPOSE_ACTION)
    u = hn.uniform01(hn.cell_hash(key, 0))
    k = min(int(np.float32(u) * np.float32(valid.size)), valid.size - 1)
    return int(valid[k])


class _ListMemory:
    """Minimal stand-in for BatchMemory (batch_memory.py:25-115): the feature builders only use
    add(observation=...), insert(-1, state=... / action=...), get(-1, agent, "observation" | "action")."""

    def __init__(self, n_agents):
        self.rows = {a: [] for a in range(n_agents)}

    def add(self, agent_id, **kw):
        self.rows[agent_id].append(dict(kw))

    def insert(self, t, agent_id, **kw):
        self.rows[agent_id][t].update({k: v for k, v in kw.items() if v is not None})

    def get(self, t, agent_id, name):
        return self.rows[agent_id][t].get(name)


def run_reference_episode(params, episode, actions=None, noiseless=False, n_steps=None, record_maps=True,
                          features=False):
    """One episode of the reference env loop; returns per-step records.

    ``features``: also call the reference's observation / critic-state builders
    (actor/transformations.py:14-59, critic/transformations.py:17-67) exactly where coma_wrapper.py:57-68
    and :135-144 call them and record their outputs ("obs" [A,P,P,7] float64, "state" [A,P,P,12] float32).

    ``actions``: optional int array [T, A]; default = uniform_policy_action.
    Records (per step t): positions before the moves, comm matrix, fused local
    maps, global map, rewards, masks, actions, positions after the moves, local
    maps after the post-move measurement.
    """
    ns = load()
    install_noise_patch()
    seed = params["environment"]["seed"]
    n_agents = params["experiment"]["missions"]["n_agents"]
    budget = params["experiment"]["constraints"]["budget"]
    T = budget + 1 if n_steps is None else n_steps
    NoiseContext.seed = seed
    NoiseContext.episode = episode
    NoiseContext.noiseless = noiseless

    grid_map = ns.GridMap(params)
    sensor = ns.Sensor(ns.AltitudeSensorModel(params), grid_map)
    mapping = ns.Mapping(grid_map, sensor, params, episode)
    ass = ns.AgentStateSpace(params)
    actor = _StubActor()
    memory = _ListMemory(n_agents) if features else _StubMemory()
    if features:
        import torch  # noqa: F401  (the builders return torch tensors)
        from actor.transformations import get_network_input as get_actor_input
        from critic.transformations import get_network_input as get_critic_input
    agents = [ns.Agent(actor, params, mapping, a, ass) for a in range(n_agents)]
    global_map = agents[0].local_map.copy()
    rec = {
        "gt": mapping.simulated_map.copy(),
        "steps": [],
    }
    for t in range(T):
        step = {}
        log = ns.CommunicationLog(params, episode)
        info = {}
        for a in range(n_agents):
            NoiseContext.agent, NoiseContext.index = a, 0
            info, _, _ = agents[a].communicate(t, episode, log, None)
        step["pos"] = np.array([np.array(agents[a].position) for a in range(n_agents)], dtype=np.int64)
        if t == 0 and record_maps:
            step["local_after_init"] = np.array([np.asarray(agents[a].local_map, dtype=np.float64) for a in range(n_agents)])
        comm = np.zeros((n_agents, n_agents), dtype=np.uint8)
        fused_local = []
        obs = []
        for a in range(n_agents):
            with _comm_draws(seed, episode, a, t):
                received, fused = agents[a].receive_messages(log, a, t)
            for j in received:
                comm[a, j] = 1
            fused_local.append(np.asarray(fused, dtype=np.float64).copy())
            if features:  # coma_wrapper.py:57-68
                o = get_actor_input(received, fused, mapping.simulated_map, a, t, params, memory, ass)
                memory.add(a, observation=o)
                obs.append(o.numpy().copy())
        if features:
            step["obs"] = np.array(obs)
        step["comm"] = comm
        if record_maps:
            step["local_fused"] = np.array(fused_local)
        # coma_wrapper.py:93-95 / 145-147 (computed twice, identically, by the reference)
        next_global = mapping.fuse_map(global_map, info, None, "global")
        moved = []
        masks, acts = [], []
        for a in range(n_agents):
            if actions is not None:
                actor.policy = lambda m, aid, tt, _a=int(actions[t][a]): _a
            else:
                actor.policy = lambda m, aid, tt: uniform_policy_action(m, seed, episode, aid, tt)
            NoiseContext.agent, NoiseContext.index = a, t + 1
            # all-zero masks: reference would raise inside torch.multinomial; our stub returns -1
            # and action_to_position(pos, -1) leaves the offset at [0,0,0] (action_space.py:199-223).
            if features:
                import torch

                _pol = actor.policy
                actor.policy = lambda m, aid, tt, _p=_pol: torch.tensor(_p(m, aid, tt))  # .item() is called on it
            _, pos, _, action, _, _ = agents[a].step(a, t, episode, memory, None, moved)
            moved.append(pos)
            masks.append(actor.last_mask.copy())
            acts.append(int(action))
        if features:  # coma_wrapper.py:135-144 (critic_map_knowledge = this step's fused global map)
            step["state"] = np.array([
                get_critic_input(t, info, next_global, memory, a, mapping.simulated_map, params).numpy().copy()
                for a in range(n_agents)
            ])
        _, rel, ab = ns.get_global_reward(
            global_map, next_global, "COMA", None, mapping.simulated_map, ass, acts, None, t, budget
        )
        step["mask"] = np.array(masks)
        step["action"] = np.array(acts, dtype=np.int64)
        step["reward_rel"] = float(rel)
        step["reward_abs"] = float(ab)
        step["pos_next"] = np.array([np.array(p) for p in moved], dtype=np.int64)
        if record_maps:
            step["global"] = np.asarray(next_global, dtype=np.float64).copy()
            step["local_after_move"] = np.array(
                [np.asarray(agents[a].local_map, dtype=np.float64) for a in range(n_agents)]
            )
        global_map = next_global
        rec["steps"].append(step)
    return rec


# ----------------------------------------------------------------------------------------------
# the reference's information-gain greedy baseline (IG_baseline.py), unmodified, with the same noise /
# communication streams as run_reference_episode
# ----------------------------------------------------------------------------------------------
def run_reference_ig(params, episode):
    """Run ``IG_baseline(params, writer, episode).execute()`` and record what it computes per step:
    individual gains, final utilities, chosen actions, and the entropy / F1 curves it returns."""
    ns = load()
    install_noise_patch()
    import marl_framework.IG_baseline as ig_mod  # noqa: E402

    agent_cls = ig_mod.Agent  # the bare-alias class IG_baseline.py:15 imports (SURVEY.md 8c "harness trap")

    seed = params["environment"]["seed"]
    n_agents = params["experiment"]["missions"]["n_agents"]
    NoiseContext.seed = seed
    NoiseContext.episode = episode
    NoiseContext.noiseless = False

    base = ig_mod.IG_baseline(params, _Stub("writer"), episode)
    rec = {"gains": [], "util": [], "action": [], "pos": []}

    # every measurement goes through mapping.update_grid_map, in agent order: A calls at t = 0 from
    # Agent.communicate (stream index 0), then A calls per planning step t (stream index t + 1)
    calls = {"n": 0}
    real_update = base.mapping.update_grid_map

    def update_grid_map(*a, **k):
        NoiseContext.agent = calls["n"] % n_agents
        NoiseContext.index = calls["n"] // n_agents
        calls["n"] += 1
        return real_update(*a, **k)

    base.mapping.update_grid_map = update_grid_map

    real_receive = agent_cls.receive_messages

    def receive_messages(self, log, agent_id, t):
        with _comm_draws(seed, episode, agent_id, t):
            return real_receive(self, log, agent_id, t)

    real_ind, real_util, real_sel = base.get_individual_ig, base.get_cell_utilities, base.select_action
    step = {}

    def get_individual_ig(position, mask, map_state):
        out = real_ind(position, mask, map_state)
        step.setdefault("pos", []).append(np.array(position))
        step.setdefault("gains", []).append([float(v) for v in out[1]])
        return out

    def get_cell_utilities(plist, rel):
        out = real_util(plist, rel)
        return out

    def select_action(util):
        a = real_sel(util)
        step.setdefault("util", []).append([float(v) for v in util])
        step.setdefault("action", []).append(int(a))
        if len(step["action"]) == n_agents:
            for k in ("gains", "util", "action", "pos"):
                rec[k].append(np.array(step[k]))
            step.clear()
        return a

    base.get_individual_ig, base.get_cell_utilities, base.select_action = get_individual_ig, get_cell_utilities, select_action
    agent_cls.receive_messages = receive_messages
    try:
        rel_sum, abs_sum, altitudes, entropies, f1s = base.execute()
    finally:
        agent_cls.receive_messages = real_receive
    out = {k: np.array(v) for k, v in rec.items()}
    out["entropy"] = np.array([float(v) for v in entropies])
    out["f1"] = np.array([float(v) for v in f1s])
    out["reward_rel_sum"] = float(rel_sum)
    out["reward_abs_sum"] = float(abs_sum)
    out["gt"] = np.asarray(base.mapping.simulated_map).copy()
    return out

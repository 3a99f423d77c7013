"""numpy restatement of the reference's per-timestep environment path ("exact" mode).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for the CUDA path and
the timed CPU baseline of bench.py; never imported by ``ipp_marl_b200``.

Every function cites the reference file:line it follows (paths relative to
/root/reference/marl_framework).  The arithmetic mirrors the reference under the
installed numpy (>= 2, NEP 50): float32 logit on the first peer pass, float64
sigmoid, float32 state at fuse entry (SURVEY.md section 7 "dtype drift").  It is
pinned bit-for-bit against the live reference by tests/test_oracle_vs_reference.py
and against the committed fixtures in tests/golden/ (written by oracle/make_golden.py).

Randomness (measurement noise, sampled actions, message failures) comes from
``oracle.noise`` on both sides, because the reference's global-RNG draws cannot
be reproduced (SURVEY.md section 7 "RNG parity").
"""
import math

import numpy as np

from . import noise as hn

NOISE_BY_ALTITUDE = {5: 0.01, 10: 0.265, 15: 0.375}  # sensors/models/sensor_models.py:13-22


class Geometry:
    """Grid / lattice constants.  mapping/grid_maps.py:16-66, agent/state_space.py:10-21."""

    def __init__(self, params):
        env = params["environment"]
        sen = params["sensor"]
        con = params["experiment"]["constraints"]
        self.params = params
        self.seed = env["seed"]
        self.x_dim_m = env["x_dim"]
        self.y_dim_m = env["y_dim"]
        self.spacing = con["spacing"]
        self.min_altitude = con["min_altitude"]
        self.max_altitude = con["max_altitude"]
        self.budget = con["budget"]
        self.n_actions = con["num_actions"]
        self.n_agents = params["experiment"]["missions"]["n_agents"]
        self.prior = params["mapping"]["prior"]
        self.comm_range = params["experiment"]["uav"]["communication_range"]
        self.fix_range = params["experiment"]["uav"].get("fix_range", True)
        self.failure_rate = params["experiment"]["uav"]["failure_rate"]
        self.angle_x = sen["field_of_view"]["angle_x"]
        self.angle_y = sen["field_of_view"]["angle_y"]
        # grid_maps.py:52-66
        self.res_x = (2 * self.min_altitude * math.tan(math.radians(self.angle_x) * 0.5)) / sen["pixel"]["number_x"]
        self.res_y = (2 * self.min_altitude * math.tan(math.radians(self.angle_y) * 0.5)) / sen["pixel"]["number_y"]
        # grid_maps.py:29-32, 46-49
        self.gx = int(self.x_dim_m / self.res_x)
        self.gy = int(self.y_dim_m / self.res_y)
        # state_space.py:16-18
        self.px = self.x_dim_m // self.spacing + 1
        self.py = self.y_dim_m // self.spacing + 1
        self.pz = (self.max_altitude - self.min_altitude) // self.spacing + 1
        self.altitudes = [self.min_altitude + i * self.spacing for i in range(self.pz)]


def footprint(geo, position):
    """sensors/cameras.py:31-79 -> (raw [yu,yd,xl,xr], clipped [yu,yd,xl,xr])."""
    position = np.asarray(position)
    x_range_m = 2 * position[2] * np.tan(0.5 * np.radians(geo.angle_x))
    y_range_m = 2 * position[2] * np.tan(0.5 * np.radians(geo.angle_y))
    cells = np.array([np.floor(x_range_m / geo.res_x), np.floor(y_range_m / geo.res_y)])
    centre = np.floor(position[:2] / geo.res_x)  # x resolution for both axes (cameras.py:66)
    radius = np.floor(0.5 * cells)
    xl, yu = centre - radius
    xr, yd = centre + radius
    raw = [int(yu), int(yd), int(xl), int(xr)]
    xl, xr = np.clip(np.array([xl, xr]), 0, geo.gx - 1)
    yu, yd = np.clip(np.array([yu, yd]), 0, geo.gy - 1)
    return raw, [int(yu), int(yd), int(xl), int(xr)]


def ground_truth(geo, episode):
    """mapping/ground_truths.py:42-56,176: half-plane field; the GRF computed before it is discarded."""
    field = np.zeros((geo.gx, geo.gy))
    rng = np.random.RandomState(episode)  # same MT19937 stream as np.random.seed(episode)
    split = rng.randint(4)
    pct = rng.randint(30, 61)
    rows, cols = field.shape
    if split == 0:
        field[: int((rows * pct) / 100), :] = 1
    elif split == 1:
        field[int((rows * (1 - pct)) / 100) :, :] = 1  # negative start index, as in the reference
    elif split == 2:
        field[:, : int((cols * pct) / 100)] = 1
    else:
        field[:, int((cols * (1 - pct)) / 100) :] = 1
    return field


def start_position(geo, agent_id, episode):
    """agent/state_space.py:28-51."""
    rng = np.random.RandomState(seed=geo.seed * episode * agent_id)
    x = geo.spacing * rng.randint(0, geo.px)
    y = geo.spacing * rng.randint(0, geo.py)
    return np.array([x, y, 15])


def measurement(geo, gt, altitude, rect, key, noiseless=False):
    """mapping/simulations.py:42-65 with hash noise: float32 [xr-xl, yd-yu] in {acc, 1-acc}."""
    yu, yd, xl, xr = rect
    section = gt[xl:xr, yu:yd]
    sensor_noise = NOISE_BY_ALTITUDE.get(int(altitude), 0)
    xs = np.arange(xl, xr, dtype=np.int64)[:, None]
    ys = np.arange(yu, yd, dtype=np.int64)[None, :]
    if noiseless:
        wrong = np.zeros(section.shape, dtype=bool)
    else:
        wrong = hn.noise_word(key, xs * geo.gy + ys) < hn.flip_threshold(sensor_noise)
    accuracy = 1 - sensor_noise
    seen = np.where(wrong, 1 - section, section)
    value = np.where(seen == 1, accuracy, 1 - accuracy)
    # np.putmask(value, (1-acc) > acc*v, 1-acc) of the reference yields exactly these two levels
    return np.float32(np.round(value, 3))


def bayes_pass(x, y, prior):
    """mapping/mappings.py:109-124: clamp x IN PLACE, log-odds sum, 1 - 1/(1+exp)."""
    x[0.9999 < x] = 0.9999
    x[0.0001 > x] = 0.0001
    l_x = np.log(x / (1 - x))
    l_y = np.log(y / (1 - y))
    l_xy = l_x + l_y
    l_p = np.log(prior / (1 - prior))
    return 1 - (1 / (1 + np.exp(l_xy - l_p)))


def update_grid_map(geo, gt, position, map_state, key, noiseless=False):
    """mapping/mappings.py:32-78 -> (map_state (mutated), rect, measurement, map2communicate)."""
    _, rect = footprint(geo, position)
    yu, yd, xl, xr = rect
    meas = measurement(geo, gt, position[2], rect, key, noiseless)
    section = map_state[xl:xr, yu:yd]
    map_state[xl:xr, yu:yd] = bayes_pass(section, meas, geo.prior)
    m2c = np.ones_like(map_state) * 0.5
    m2c[xl:xr, yu:yd] = meas
    return map_state, rect, meas, m2c


def fuse(geo, own, others):
    """mapping/mappings.py:80-104: successive whole-map passes; float32 cast once at entry."""
    fused = np.float32(own.copy())
    for other in others:
        fused = bayes_pass(fused, np.float32(other), geo.prior)
    return fused


def comm_matrix(geo, positions, episode, t):
    """agent/communication_log.py:39-58: row i = agents whose message i receives (incl. itself)."""
    n = len(positions)
    out = np.zeros((n, n), dtype=np.uint8)
    comm_range = geo.comm_range
    if not geo.fix_range:  # communication_log.py:22-31: re-drawn (identically) by every CommunicationLog(params, episode)
        comm_range = (0, 15, 25, 100)[np.random.RandomState(episode).randint(4)]
    for i in range(n):
        key = hn.stream_key(geo.seed, episode, i, t, hn.PURPOSE_COMM)
        for j in range(n):
            d = np.linalg.norm(np.asarray(positions[i]) - np.asarray(positions[j]), ord=2)
            r = float(hn.uniform01(hn.cell_hash(key, j)))
            ok = d < 0.001
            if 0.001 <= d <= comm_range and r >= geo.failure_rate:
                ok = True
            out[i, j] = ok
    return out


def shannon_entropy(p):
    """utils/state.py:118-121 (clamps its argument in place)."""
    p[0.0001 > p] = 0.0001
    p[0.9999 < p] = 0.9999
    return -p * np.log2(p) - (1 - p) * np.log2(1 - p)


def _weights_and_entropy(m):
    """utils/state.py:53-76, "reward" branch: weights from the thresholded map itself."""
    grid = m.copy()
    target = grid.copy()
    target[target > 0.501] = 1
    target[target < 0.499] = 0
    w = target.copy()
    w[np.round(w, 2) == 0] = 0
    w[np.round(w, 2) == 1] = 1
    w[np.round(w, 2) == 0.5] = 0.5
    return w, shannon_entropy(grid)


def global_reward(last_map, next_map):
    """utils/reward.py:11-53,68-82 -> (relative_reward, absolute_reward)."""
    _, h_before = _weights_and_entropy(last_map)
    w, h_after = _weights_and_entropy(next_map)
    reduction = h_before - h_after
    absolute = np.mean(w * reduction)
    relative = absolute / np.mean(w * h_before)
    return 22 * relative - 0.5, 10 * absolute - 0.17


def action_mask(geo, position):
    """agent/action_space.py:56-70 (6 actions)."""
    mask = np.ones(6)
    if position[2] == geo.max_altitude:
        mask[0] = 0
    if position[2] == geo.min_altitude:
        mask[5] = 0
    if position[1] == 0:
        mask[2] = 0
    if position[1] == geo.y_dim_m:
        mask[3] = 0
    if position[0] == 0:
        mask[1] = 0
    if position[0] == geo.x_dim_m:
        mask[4] = 0
    return mask


def collision_mask(geo, position, mask, moved):
    """agent/action_space.py:309-344 (6 actions): order dependent, guarded by sum(mask) > 1."""
    for other in moved:
        dx = other[0] // geo.spacing - position[0] // geo.spacing
        dy = other[1] // geo.spacing - position[1] // geo.spacing
        if dx == 0 and dy == 0:
            if np.sum(mask) > 1:
                mask[0] = 0
                mask[5] = 0
        if dx == -1 and dy == 0:
            if np.sum(mask) > 1:
                mask[1] = 0
        if dx == 0 and dy == -1:
            if np.sum(mask) > 1:
                mask[2] = 0
        if dx == 0 and dy == 1:
            if np.sum(mask) > 1:
                mask[3] = 0
        if dx == 1 and dy == 0:
            if np.sum(mask) > 1:
                mask[4] = 0
    return mask


_OFFSETS = {0: (0, 0, 1), 1: (-1, 0, 0), 2: (0, -1, 0), 3: (0, 1, 0), 4: (1, 0, 0), 5: (0, 0, -1)}


def move(geo, position, action):
    """agent/action_space.py:211-223; action -1 (all-zero mask) = stay."""
    off = _OFFSETS.get(int(action), (0, 0, 0))
    return np.asarray(position) + geo.spacing * np.array(off)


def uniform_action(geo, mask, episode, agent, t):
    """Uniform over unmasked actions (SURVEY.md section 8d policy for c2/c5)."""
    valid = np.flatnonzero(mask > 0)
    if valid.size == 0:
        return -1
    key = hn.stream_key(geo.seed, episode, agent, t, hn.PURPOSE_ACTION)
    u = hn.uniform01(hn.cell_hash(key, 0))
    k = min(int(np.float32(u) * np.float32(valid.size)), valid.size - 1)
    return int(valid[k])


class OracleEnv:
    """One environment instance stepped in the reference's order of effects.

    agent/agent.py:40-104 + coma_wrapper.py:37-183 + missions/episode_generator.py:38-56.
    """

    def __init__(self, params, episode, noiseless=False):
        self.geo = Geometry(params)
        self.episode = episode
        self.noiseless = noiseless
        g = self.geo
        self.gt = ground_truth(g, episode)
        self.local = [np.full((g.gx, g.gy), g.prior, dtype="float32") for _ in range(g.n_agents)]
        self.global_map = self.local[0].copy()
        self.pos = [None] * g.n_agents
        self.m2c = [None] * g.n_agents
        self.t = 0

    def _key(self, agent, index):
        return hn.stream_key(self.geo.seed, self.episode, agent, index, hn.PURPOSE_NOISE)

    def observe(self):
        """build_observations + the global fuse and reward of steps (everything before the moves)."""
        g = self.geo
        if self.t == 0:
            for a in range(g.n_agents):
                self.pos[a] = start_position(g, a, self.episode)
                self.local[a], _, _, self.m2c[a] = update_grid_map(
                    g, self.gt, self.pos[a], self.local[a], self._key(a, 0), self.noiseless
                )
        comm = comm_matrix(g, self.pos, self.episode, self.t)
        snapshot = list(self.m2c)
        for a in range(g.n_agents):
            peers = [snapshot[j] for j in range(g.n_agents) if comm[a, j] and j != a]
            self.local[a] = fuse(g, self.local[a], peers)
        next_global = fuse(g, self.global_map, snapshot)
        rel, ab = global_reward(self.global_map, next_global)
        self.global_map = next_global
        return comm, rel, ab

    def act(self, actions=None):
        """Sequential mask / choose / move / measure (agent/agent.py:73-104)."""
        g = self.geo
        moved, masks, acts = [], [], []
        for a in range(g.n_agents):
            mask = collision_mask(g, self.pos[a], action_mask(g, self.pos[a]), moved)
            if actions is None:
                act = uniform_action(g, mask, self.episode, a, self.t)
            else:
                act = int(actions[a])
            self.pos[a] = move(g, self.pos[a], act)
            self.local[a], _, _, self.m2c[a] = update_grid_map(
                g, self.gt, self.pos[a], self.local[a], self._key(a, self.t + 1), self.noiseless
            )
            moved.append(self.pos[a])
            masks.append(mask)
            acts.append(act)
        self.t += 1
        return np.array(masks), np.array(acts, dtype=np.int64)


def run_episode(params, episode, actions=None, noiseless=False, n_steps=None, record_maps=True):
    """Same record layout as oracle.ref_harness.run_reference_episode."""
    env = OracleEnv(params, episode, noiseless)
    g = env.geo
    T = g.budget + 1 if n_steps is None else n_steps
    rec = {"gt": env.gt.copy(), "steps": []}
    for t in range(T):
        step = {}
        comm, rel, ab = env.observe()
        step["pos"] = np.array(env.pos, dtype=np.int64)
        step["comm"] = comm
        if record_maps:
            step["local_fused"] = np.array([np.asarray(m, dtype=np.float64) for m in env.local])
            step["global"] = np.asarray(env.global_map, dtype=np.float64).copy()
        masks, acts = env.act(None if actions is None else actions[t])
        step["mask"] = masks
        step["action"] = acts
        step["reward_rel"] = float(rel)
        step["reward_abs"] = float(ab)
        step["pos_next"] = np.array(env.pos, dtype=np.int64)
        if record_maps:
            step["local_after_move"] = np.array([np.asarray(m, dtype=np.float64) for m in env.local])
        rec["steps"].append(step)
    return rec

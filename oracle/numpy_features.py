"""numpy restatement of the reference's observation / critic-state builders.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows actor/transformations.py:14-176,
critic/transformations.py:17-132 and utils/state.py:14-121 ("actor"/"global" branches), including
cv2.resize(INTER_AREA) for the lattice-sized area pooling (cv2 is the oracle there: SURVEY.md
Appendix A), and is pinned bit-for-bit against the live reference's outputs (tests/golden features).
"""
import cv2
import numpy as np

from . import numpy_oracle as no


def _area(m, geo):
    """cv2.resize(m, (P_y, P_x), INTER_AREA): utils/state.py:22-41, transformations.py:77-81."""
    return cv2.resize(m, (int(geo.py), int(geo.px)), interpolation=cv2.INTER_AREA)


def _weights(target):
    """utils/state.py:67-76 on an already down-sampled map."""
    target = target.copy()
    target[target > 0.501] = 1
    target[target < 0.499] = 0
    w = target.copy()
    w[np.round(w, 2) == 0] = 0
    w[np.round(w, 2) == 1] = 1
    w[np.round(w, 2) == 0.5] = 0.5
    return w


def w_entropy_lattice(m, geo):
    """(w*H, prob_map) of area(m): utils/state.py:21-26,46-115."""
    grid = _area(m, geo)
    w = _weights(grid)
    se = no.shannon_entropy(grid)  # clamps `grid` in place, like the reference: prob_map is the clamped map
    return w * se, grid


def footprint_img(geo, position, meas):
    """mapping/mappings.py:41-43,72-76 + utils/utils.py:79-98: raw-footprint-sized image, 0.5 outside the map."""
    raw, clipped = no.footprint(geo, position)
    img = np.ones((raw[1] - raw[0], raw[3] - raw[2])) * 0.5
    h, w = raw[1] - raw[0], raw[3] - raw[2]
    yu, yd, xl, xr = 0, h, 0, w
    if clipped[0] > raw[0]:
        yu = h - (clipped[1] - clipped[0])
    if clipped[1] < raw[1]:
        yd = clipped[1] - clipped[0]
    if clipped[3] < raw[3]:
        xr = clipped[3] - clipped[2]
    if clipped[2] > raw[2]:
        xl = w - (clipped[3] - clipped[2])
    img[xl:xr, yu:yd] = meas
    return img


def position_map_ego(geo, positions, received, agent_id):
    """actor/transformations.py:110-176 (window hard-coded around lattice index 5, like the reference)."""
    pm = np.ones((geo.px, geo.py))
    own = np.array([positions[agent_id][0] // geo.spacing, positions[agent_id][1] // geo.spacing,
                    positions[agent_id][2] // geo.spacing - 1])
    if own[0] < 5:
        pm[0:5 - own[0], :] = 0
    if own[1] < 5:
        pm[:, 0:5 - own[1]] = 0
    if own[0] > 5:
        pm[geo.px - 1 - (own[0] - 6):, :] = 0
    if own[1] > 5:
        pm[:, geo.py - 1 - (own[1] - 6):] = 0
    rel = [[5, 5, (own[2] + 1) / (geo.pz + 1)]]
    for j in received:
        if j == agent_id:
            continue
        o = np.array([positions[j][0] // geo.spacing, positions[j][1] // geo.spacing, positions[j][2] // geo.spacing - 1])
        rel.append([o[0] - own[0] + 5, o[1] - own[1] + 5, (o[2] + 1) / (geo.pz + 1)])
    for r in rel:
        if 0 <= r[0] < geo.px and 0 <= r[1] < geo.px:  # the y bound uses space_x_dim in the reference (:170)
            pm[int(r[0]), int(r[1])] = r[2]
    return pm


def footprint_ownership(geo, m2c, received, agent_id):
    """actor/transformations.py:62-83: 1 own footprint, 0 received peers' footprints, own overrides."""
    fm = m2c[agent_id].copy()
    fm[fm < 0.49] = 1
    fm[fm > 0.51] = 1
    for j in received:
        if j == agent_id:
            continue
        other = m2c[j]
        fm[other < 0.49] = 0
        fm[other > 0.51] = 0
    own = m2c[agent_id]
    fm[own < 0.49] = 1
    fm[own > 0.51] = 1
    return _area(fm, geo)


def actor_observation(geo, t, agent_id, positions, received, fused_local, m2c, fp_imgs):
    """actor/transformations.py:14-59 -> [P, P, 7] float64."""
    pm = position_map_ego(geo, positions, received, agent_id)
    w_entropy, prob = w_entropy_lattice(np.asarray(fused_local), geo)
    grid_fp = _area(fp_imgs[agent_id], geo)
    wf = _weights(grid_fp)
    local_w_entropy = wf * no.shannon_entropy(grid_fp)
    budget = np.ones_like(pm) * ((geo.budget - t) / geo.budget)
    agent = np.ones_like(pm) * ((agent_id + 1) / geo.n_agents)
    fo = footprint_ownership(geo, m2c, received, agent_id)
    return np.dstack([budget, agent, pm, w_entropy, local_w_entropy, prob, fo])


def critic_state(geo, agent_id, observation, positions, global_map, m2c, actions):
    """critic/transformations.py:17-132 -> [P, P, 12] float32."""
    pos_map = np.zeros((geo.px, geo.py))
    for j in range(geo.n_agents):
        p = positions[j]
        pos_map[p[0] // geo.spacing, p[1] // geo.spacing] = (p[2] // geo.spacing - 1 + 1) / geo.pz
    w_entropy, prob = w_entropy_lattice(np.asarray(global_map), geo)
    fm = m2c[0].copy()
    fm[fm < 0.49] = 1
    fm[fm > 0.51] = 1
    for j in range(1, geo.n_agents):
        fm[m2c[j] < 0.49] = 1
        fm[m2c[j] > 0.51] = 1
    fm = _area(fm, geo)
    act_map = np.zeros((geo.px, geo.py))
    for j in range(geo.n_agents):
        if j != agent_id:
            p = positions[j]
            act_map[p[0] // geo.spacing, p[1] // geo.spacing] = (int(actions[j]) + 1) / geo.n_actions
    return np.dstack((observation, pos_map[:, :, None], w_entropy, prob, fm[:, :, None],
                      act_map[:, :, None])).astype(np.float32)


class FeatureOracleEnv(no.OracleEnv):
    """OracleEnv that also builds the observations / critic states at the reference's call sites."""

    def __init__(self, params, episode, noiseless=False):
        super().__init__(params, episode, noiseless)
        self.fp_img = [None] * self.geo.n_agents
        self.meas = [None] * self.geo.n_agents

    def _measure(self, a, index):
        g = self.geo
        self.local[a], rect, meas, self.m2c[a] = no.update_grid_map(
            g, self.gt, self.pos[a], self.local[a], self._key(a, index), self.noiseless)
        self.fp_img[a] = footprint_img(g, self.pos[a], meas)

    def observe_features(self):
        """build_observations + global fuse + reward; returns (comm, rel, abs, obs[A,P,P,7])."""
        g = self.geo
        if self.t == 0:
            for a in range(g.n_agents):
                self.pos[a] = no.start_position(g, a, self.episode)
                self._measure(a, 0)
        comm = no.comm_matrix(g, self.pos, self.episode, self.t)
        snapshot = list(self.m2c)
        self.snap_pos = [np.array(p) for p in self.pos]
        self.snap_m2c = snapshot
        obs = []
        for a in range(g.n_agents):
            received = [j for j in range(g.n_agents) if comm[a, j]]
            peers = [snapshot[j] for j in received if j != a]
            self.local[a] = no.fuse(g, self.local[a], peers)
            obs.append(actor_observation(g, self.t, a, self.snap_pos, received, self.local[a], snapshot, self.fp_img))
        next_global = no.fuse(g, self.global_map, snapshot)
        rel, ab = no.global_reward(self.global_map, next_global)
        self.global_map = next_global
        self.obs = obs
        return comm, rel, ab, np.array(obs)

    def act_features(self, actions=None):
        g = self.geo
        moved, masks, acts = [], [], []
        for a in range(g.n_agents):
            mask = no.collision_mask(g, self.pos[a], no.action_mask(g, self.pos[a]), moved)
            act = no.uniform_action(g, mask, self.episode, a, self.t) if actions is None else int(actions[a])
            self.pos[a] = no.move(g, self.pos[a], act)
            self._measure(a, self.t + 1)
            moved.append(self.pos[a])
            masks.append(mask)
            acts.append(act)
        states = np.array([critic_state(g, a, self.obs[a], self.snap_pos, self.global_map, self.snap_m2c, acts)
                           for a in range(g.n_agents)])
        self.t += 1
        return np.array(masks), np.array(acts, dtype=np.int64), states

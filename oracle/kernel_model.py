"""Bit-level numpy model of the CUDA kernels' arithmetic ("f32 mode"), batched over envs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Two jobs:
  * it is the specification of ``ipp_marl_b200/csrc`` — the belief maps written by
    the kernels must equal this model BIT FOR BIT (only IEEE +,-,*,/,min,max in
    float32 are used on the belief path, so numpy reproduces them exactly);
  * it is checked against ``oracle.numpy_oracle`` (= the reference's arithmetic)
    with the SURVEY.md section 8d gate ``allclose(rtol=1e-5, atol=1e-5)``, at batch
    sizes the per-env oracle is too slow for.

Arithmetic (DESIGN.md "belief update in odds space"): the reference's
``1 - 1/(1 + exp(logit(x) + logit(y) - logit(prior)))`` (mapping/mappings.py:109-124)
is algebraically ``o' = o * k`` on the odds ``o = x/(1-x)`` with
``k = exp(logit(y) - logit(prior))``; ``k`` takes two values per altitude (cell seen
as 1 / as 0) plus ``k_out`` for the 0.5 "not observed" cells of map2communicate.
The clamp the reference applies before every pass is kept (in odds space).

STATE IN ODDS SPACE: the belief maps are stored as float32 odds ``o`` (``local_o`` / ``glob_o``), so a pass
is ``o = min(max(o, o_min), o_max) * k`` — one clamp and one multiply per cell, no division and no
transcendental anywhere on the belief path.  Probabilities exist only where somebody reads them
(``local`` / ``glob`` properties here, ``ipp_export_beliefs`` and the feature / planner / metric kernels on
the device): ``p = o/(1+o)`` for ``o < 1`` else ``1 - 1/(1+o)`` (symmetric form: probabilities near 1 keep
their accuracy).  float32 odds carry at least the relative precision of float32 probabilities
(``dp/p = (1-p) do/o``), so the once-per-step rounding of the reference's float32 state is bounded by
this model's rounding.
"""
import numpy as np

from . import noise as hn
from . import numpy_oracle as no

F32 = np.float32
# weight thresholds of utils/state.py:67-73 (p > 0.501 / p < 0.499) in odds space
W_HI = F32(0.501 / 0.499)
W_LO = F32(0.499 / 0.501)


class KernelTables:
    """Host-side tables handed to the kernels (computed with the reference's own expressions)."""

    def __init__(self, params):
        geo = no.Geometry(params)
        self.geo = geo
        self.gx, self.gy = geo.gx, geo.gy
        self.n_alt = geo.pz
        self.rx = np.zeros(self.n_alt, np.int32)
        self.ry = np.zeros(self.n_alt, np.int32)
        for i, z in enumerate(geo.altitudes):
            raw, _ = no.footprint(geo, np.array([0, 0, z]))
            # raw = [yu, yd, xl, xr] around cell 0: radius = upper bound
            self.ry[i] = raw[1]
            self.rx[i] = raw[3]
        # cameras.py:66: floor(position / res_x) for BOTH axes
        self.cell_x = np.array([int(np.floor(i * geo.spacing / geo.res_x)) for i in range(geo.px)], np.int32)
        self.cell_y = np.array([int(np.floor(i * geo.spacing / geo.res_x)) for i in range(geo.py)], np.int32)
        prior = geo.prior
        l_p = np.log(prior / (1 - prior))
        self.k_hi = np.zeros(self.n_alt, F32)
        self.k_lo = np.zeros(self.n_alt, F32)
        self.thresh = np.zeros(self.n_alt, np.uint32)
        for i, z in enumerate(geo.altitudes):
            noise = no.NOISE_BY_ALTITUDE.get(int(z), 0)
            acc = 1 - noise
            y_hi = np.float32(np.round(acc, 3))
            y_lo = np.float32(np.round(1 - acc, 3))
            with np.errstate(divide="ignore"):
                l_hi = np.log(y_hi / (1 - y_hi))  # float32, as mappings.py:113 evaluates it
                l_lo = np.log(y_lo / (1 - y_lo))
            self.k_hi[i] = F32(np.exp(np.float64(l_hi) - l_p))
            self.k_lo[i] = F32(np.exp(np.float64(l_lo) - l_p))
            self.thresh[i] = hn.flip_threshold(noise)
        l_half = np.log(F32(0.5) / (1 - F32(0.5)))
        self.k_out = F32(np.exp(np.float64(l_half) - l_p))
        # clamps: float32 in probability space (first pass of the reference), and the loosest
        # odds-space bounds covering both the float32 and the float64 clamp of later passes
        self.p_min = F32(0.0001)
        self.p_max = F32(0.9999)
        o_lo32 = self.p_min / (F32(1) - self.p_min)
        o_hi32 = self.p_max / (F32(1) - self.p_max)
        self.o_min = min(F32(0.0001 / 0.9999), o_lo32)
        self.o_max = max(F32(0.9999 / 0.0001), o_hi32)
        # comm range: largest integer squared distance still within range (communication_log.py:49-53)
        def d2_of(r):
            d2 = int(np.floor(r * r)) + 2
            while d2 > 0 and not (np.sqrt(np.float64(d2)) <= r):
                d2 -= 1
            return d2 if r >= 0 else -1

        self.comm_d2_max = d2_of(float(geo.comm_range))
        self.comm_d2_table = np.array([d2_of(float(r)) for r in (0, 15, 25, 100)], dtype=np.int64)
        fr = float(geo.failure_rate)
        # r >= failure_rate with r = n / 2^24  <=>  n >= ceil(fr * 2^24)
        n = int(np.ceil(fr * 16777216.0))
        while n > 0 and (n - 1) / 16777216.0 >= fr:
            n -= 1
        while n / 16777216.0 < fr:
            n += 1
        self.fail_thresh24 = n


def _rects(tab, pos):
    """Clipped footprint [yu, yd, xl, xr] of positions [..., 3] (metres) via the tables."""
    geo = tab.geo
    ix = pos[..., 0] // geo.spacing
    iy = pos[..., 1] // geo.spacing
    iz = pos[..., 2] // geo.spacing - geo.min_altitude // geo.spacing
    cx = tab.cell_x[ix]
    cy = tab.cell_y[iy]
    rx = tab.rx[iz]
    ry = tab.ry[iz]
    xl = np.clip(cx - rx, 0, tab.gx - 1)
    xr = np.clip(cx + rx, 0, tab.gx - 1)
    yu = np.clip(cy - ry, 0, tab.gy - 1)
    yd = np.clip(cy + ry, 0, tab.gy - 1)
    return np.stack([yu, yd, xl, xr], axis=-1), iz


class KernelModelEnv:
    """Batched env with the kernels' arithmetic.  State arrays mirror the device layout."""

    def __init__(self, params, episodes, noiseless=False):
        self.tab = KernelTables(params)
        geo = self.tab.geo
        self.geo = geo
        self.episodes = np.asarray(episodes, dtype=np.int64)
        B, A = len(self.episodes), geo.n_agents
        self.B, self.A = B, A
        self.noiseless = noiseless
        gx, gy = self.tab.gx, self.tab.gy
        self.gt = np.stack([no.ground_truth(geo, int(e)) for e in self.episodes]).astype(np.uint8)
        o_prior = F32(geo.prior) / (F32(1) - F32(geo.prior))
        self.local_o = np.full((B, A, gx, gy), o_prior, dtype=F32)
        self.glob_o = np.full((B, gx, gy), o_prior, dtype=F32)
        self.pos = np.stack(
            [[no.start_position(geo, a, int(e)) for a in range(A)] for e in self.episodes]
        ).astype(np.int64)
        self.t = 0
        self.flag_stuck = np.zeros(B, dtype=bool)
        xs = np.arange(gx, dtype=np.int64)[:, None]
        ys = np.arange(gy, dtype=np.int64)[None, :]
        self._xs, self._ys = xs, ys
        self._cell = (xs * gy + ys)[None]
        # initial measurement at the start positions (agent/agent.py:44-49): own-update pass only
        for a in range(A):
            inr, k = self._k_of(self.pos[:, a], a, 0)
            self.local_o[:, a] = self._apply(self.local_o[:, a], [(inr, k, False)])

    # ---- probabilities (what ipp_export_beliefs returns) -------------------------------------
    @staticmethod
    def to_p(o):
        o = np.asarray(o, dtype=F32)
        one = F32(1)
        with np.errstate(over="ignore", invalid="ignore"):
            d = one + o
            return np.where(o < one, o / d, one - one / d).astype(F32)

    @property
    def local(self):
        return self.to_p(self.local_o)

    @property
    def glob(self):
        return self.to_p(self.glob_o)

    # ---- measurement multipliers ---------------------------------------------------------
    def _k_of(self, pos, agent, index):
        """(in_rect [B,gx,gy] bool, k [B,gx,gy] f32) of the measurement (agent, index) taken at pos."""
        tab = self.tab
        rect, iz = _rects(tab, pos)
        yu, yd, xl, xr = (rect[:, i][:, None, None] for i in range(4))
        inr = (self._xs[None] >= xl) & (self._xs[None] < xr) & (self._ys[None] >= yu) & (self._ys[None] < yd)
        key = hn.stream_key(self.geo.seed, self.episodes, agent, index, hn.PURPOSE_NOISE)
        if self.noiseless:
            wrong = np.zeros(inr.shape, dtype=bool)
        else:
            h = hn.noise_word(key[:, None, None], self._cell)
            wrong = h < tab.thresh[iz][:, None, None]
        seen_one = (self.gt != 0) ^ wrong
        k = np.where(seen_one, tab.k_hi[iz][:, None, None], tab.k_lo[iz][:, None, None]).astype(F32)
        return inr, k

    # ---- one map through a list of passes ------------------------------------------------
    def _apply(self, o, passes):
        """passes: list of (in_rect, k, is_fuse[, enabled[B]]).  Odds in, odds out (float32).

        fuse pass : every cell is clamped, cells in the rect *= k, the others *= k_out
        own update: only cells in the rect are clamped and *= k  (mappings.py:46-61)
        """
        tab = self.tab
        o = o.astype(F32, copy=True)
        B = o.shape[0]
        for ps in passes:
            inr, k, is_fuse = ps[0], ps[1], ps[2]
            en = ps[3][:, None, None] if len(ps) > 3 else np.ones((B, 1, 1), dtype=bool)
            oc = np.minimum(np.maximum(o, tab.o_min), tab.o_max)
            if is_fuse:
                kk = np.where(inr, k, tab.k_out).astype(F32)
                o = np.where(en, oc * kk, o).astype(F32)
            else:
                o = np.where(en & inr, oc * k, o).astype(F32)
        return o

    # ---- comm matrix ------------------------------------------------------------------------
    def comm(self):
        tab, A = self.tab, self.A
        out = np.zeros((self.B, A, A), dtype=bool)
        d2_max = tab.comm_d2_max
        if not self.geo.fix_range:  # per-env range index = first randint(4) of the episode's MT19937 stream
            idx = np.array([np.random.RandomState(int(e)).randint(4) for e in np.asarray(self.episodes).ravel()])
            d2_max = tab.comm_d2_table[idx]
        for i in range(A):
            key = hn.stream_key(self.geo.seed, self.episodes, i, self.t, hn.PURPOSE_COMM)
            for j in range(A):
                d = self.pos[:, i] - self.pos[:, j]
                d2 = (d * d).sum(-1)
                n24 = hn.cell_hash(key, j) >> np.uint32(8)
                out[:, i, j] = (d2 == 0) | ((d2 <= d2_max) & (n24 >= tab.fail_thresh24))
        return out

    # ---- masks / moves ------------------------------------------------------------------------
    def _choose_and_move(self, actions):
        geo, A, B = self.geo, self.A, self.B
        sp = geo.spacing
        new_pos = self.pos.copy()
        masks = np.zeros((B, A, 6), dtype=np.uint8)
        acts = np.zeros((B, A), dtype=np.int64)
        for a in range(A):
            p = self.pos[:, a]
            m = np.ones((B, 6), dtype=bool)
            m[:, 0] &= p[:, 2] != geo.max_altitude
            m[:, 5] &= p[:, 2] != geo.min_altitude
            m[:, 2] &= p[:, 1] != 0
            m[:, 3] &= p[:, 1] != geo.y_dim_m
            m[:, 1] &= p[:, 0] != 0
            m[:, 4] &= p[:, 0] != geo.x_dim_m
            for j in range(a):
                q = new_pos[:, j]
                dx = q[:, 0] // sp - p[:, 0] // sp
                dy = q[:, 1] // sp - p[:, 1] // sp
                for cond, idxs in (
                    ((dx == 0) & (dy == 0), (0, 5)),
                    ((dx == -1) & (dy == 0), (1,)),
                    ((dx == 0) & (dy == -1), (2,)),
                    ((dx == 0) & (dy == 1), (3,)),
                    ((dx == 1) & (dy == 0), (4,)),
                ):
                    go = cond & (m.sum(1) > 1)
                    for ix in idxs:
                        m[go, ix] = False
            cnt = m.sum(1)
            if actions is None:
                key = hn.stream_key(geo.seed, self.episodes, a, self.t, hn.PURPOSE_ACTION)
                u = hn.uniform01(hn.cell_hash(key, 0))
                kth = np.minimum((u * cnt.astype(F32)).astype(np.int64), np.maximum(cnt - 1, 0))
                order = np.cumsum(m, axis=1) - 1
                pick = (m & (order == kth[:, None])).argmax(1)
                act = np.where(cnt > 0, pick, -1)
            else:
                act = np.asarray(actions)[:, a].astype(np.int64).copy()
                act[(act < -1) | (act >= 6)] = -1
                # an injected action that would leave the lattice becomes "stay" (kernel rule)
                bounds = np.ones((B, 6), dtype=bool)
                bounds[:, 0] &= p[:, 2] != geo.max_altitude
                bounds[:, 5] &= p[:, 2] != geo.min_altitude
                bounds[:, 2] &= p[:, 1] != 0
                bounds[:, 3] &= p[:, 1] != geo.y_dim_m
                bounds[:, 1] &= p[:, 0] != 0
                bounds[:, 4] &= p[:, 0] != geo.x_dim_m
                bad = (act >= 0) & ~np.take_along_axis(bounds, np.maximum(act, 0)[:, None], 1)[:, 0]
                act[bad] = -1
            self.flag_stuck |= cnt == 0
            off = np.zeros((B, 3), dtype=np.int64)
            off[act == 0, 2] = sp
            off[act == 1, 0] = -sp
            off[act == 2, 1] = -sp
            off[act == 3, 1] = sp
            off[act == 4, 0] = sp
            off[act == 5, 2] = -sp
            new_pos[:, a] = p + off
            masks[:, a] = m
            acts[:, a] = act
        return new_pos, masks, acts

    # ---- reward -------------------------------------------------------------------------------
    def _reward(self, last, nxt):
        """utils/reward.py:68-82 on odds maps: float32 per-cell terms and float64 sums.  H of the clamped odds
        (p = o/(1+o), q = 1/(1+o)); the weights compare the NEXT odds with 0.501/0.499 and 0.499/0.501.
        The reward is NOT part of the bit-exact specification (only the belief maps are): the kernel evaluates the
        same H as lg(1+o) - p lg(o) with the approximate MUFU reciprocal / log2 and sums per warp in float32, and is
        compared with this model — and with the reference's golden rewards — at rtol = atol = 1e-5."""
        tab = self.tab

        def H(o):
            oc = np.minimum(np.maximum(o, tab.o_min), tab.o_max).astype(F32)
            q = (F32(1) / (F32(1) + oc)).astype(F32)
            pc = (oc * q).astype(F32)
            return (-pc * np.log2(pc) - q * np.log2(q)).astype(F32)

        w = np.where(nxt > W_HI, F32(1), np.where(nxt < W_LO, F32(0), F32(0.5)))
        hl, hn_ = H(last), H(nxt)
        s1 = (w * (hl - hn_)).astype(np.float64).sum(axis=(1, 2))
        s2 = (w * hl).astype(np.float64).sum(axis=(1, 2))
        n = float(self.tab.gx * self.tab.gy)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = 22.0 * (s1 / s2) - 0.5
        ab = 10.0 * (s1 / n) - 0.17
        return rel.astype(F32), ab.astype(F32)

    # ---- one fused env step (ipp_step) ---------------------------------------------------------
    def step(self, actions=None):
        A = self.A
        comm = self.comm()
        prev = [self._k_of(self.pos[:, j], j, self.t) for j in range(A)]
        new_pos, masks, acts = self._choose_and_move(actions)
        new = [self._k_of(new_pos[:, i], i, self.t + 1) for i in range(A)]
        last = self.glob_o
        self.glob_o = self._apply(last, [(prev[j][0], prev[j][1], True) for j in range(A)])
        rel, ab = self._reward(last, self.glob_o)
        fused = np.empty_like(self.local_o)
        for i in range(A):
            passes = [(prev[j][0], prev[j][1], True, comm[:, i, j]) for j in range(A) if j != i]
            fused[:, i] = self._apply(self.local_o[:, i], passes)
            self.local_o[:, i] = self._apply(self.local_o[:, i], passes + [(new[i][0], new[i][1], False)])
        self.local_fused_model = self.to_p(fused)  # what a separate observe kernel would have stored
        self.pos = new_pos
        self.t += 1
        return dict(comm=comm, mask=masks, action=acts, reward_rel=rel, reward_abs=ab)

    # ---- the same timestep split around a policy network (ipp_observe / ipp_act) ----------------
    def observe(self):
        """Fuse + reward only (same float32 odds state, so split and fused modes agree bit for bit)."""
        A = self.A
        comm = self.comm()
        prev = [self._k_of(self.pos[:, j], j, self.t) for j in range(A)]
        last = self.glob_o
        self.glob_o = self._apply(last, [(prev[j][0], prev[j][1], True) for j in range(A)])
        rel, ab = self._reward(last, self.glob_o)
        for i in range(A):
            passes = [(prev[j][0], prev[j][1], True, comm[:, i, j]) for j in range(A) if j != i]
            self.local_o[:, i] = self._apply(self.local_o[:, i], passes)
        return dict(comm=comm, reward_rel=rel, reward_abs=ab)

    def act(self, actions=None):
        A = self.A
        new_pos, masks, acts = self._choose_and_move(actions)
        for i in range(A):
            inr, k = self._k_of(new_pos[:, i], i, self.t + 1)
            self.local_o[:, i] = self._apply(self.local_o[:, i], [(inr, k, False)])
        self.pos = new_pos
        self.t += 1
        return dict(mask=masks, action=acts)


class SparseKernelModelEnv(KernelModelEnv):
    """The kernels' footprint-sparse processing rule on top of the dense model ("Sparsity
    without changing results"): a quad (4 cells) of a local map is PROCESSED — taken through the map's whole pass
    chain — only if some enabled fuse pass or the own update has a cell in it, or if its tile (32 quads) is flagged
    "may hold odds outside [o_min, o_max]" while a fuse pass (= whole-map clamp) runs; every other quad keeps its
    bits.  ``flags[b, i, tile]`` is maintained exactly like ``ipp_state.map_flags``.  tests/test_kernel_model.py steps
    this class next to the dense model and requires bit-identical maps: the CPU-side proof that skipping is exact."""

    TILE_CELLS = 128

    def __init__(self, params, episodes, noiseless=False):
        super().__init__(params, episodes, noiseless)
        n_cells = self.tab.gx * self.tab.gy
        self.n_tiles = (n_cells + self.TILE_CELLS - 1) // self.TILE_CELLS
        self.flags = self._tile_any(self._out_of_range(self.local_o))  # reset: the t = 0 measurement is one own pass
        self.processed_pairs = 0
        self.touched_pairs = 0
        self.total_pairs = 0

    def _out_of_range(self, o):
        return (o > self.tab.o_max) | (o < self.tab.o_min)

    def _quads(self, mask):
        """[..., gx, gy] cell mask -> [..., n_quads] 'some cell of the quad'."""
        flat = mask.reshape(mask.shape[:-2] + (-1,))
        pad = (-flat.shape[-1]) % 4
        if pad:
            flat = np.concatenate([flat, np.zeros(flat.shape[:-1] + (pad,), bool)], -1)
        return flat.reshape(flat.shape[:-1] + (-1, 4)).any(-1)

    def _tile_any(self, mask):
        """[..., gx, gy] cell mask -> [..., n_tiles]."""
        flat = mask.reshape(mask.shape[:-2] + (-1,))
        pad = self.n_tiles * self.TILE_CELLS - flat.shape[-1]
        if pad:
            flat = np.concatenate([flat, np.zeros(flat.shape[:-1] + (pad,), bool)], -1)
        return flat.reshape(flat.shape[:-1] + (self.n_tiles, self.TILE_CELLS)).any(-1)

    def _cells_of_tiles(self, tmask):
        """[..., n_tiles] -> [..., gx, gy] (every cell of a selected tile)."""
        n_cells = self.tab.gx * self.tab.gy
        cells = np.repeat(tmask, self.TILE_CELLS, axis=-1)[..., :n_cells]
        return cells.reshape(cells.shape[:-1] + (self.tab.gx, self.tab.gy))

    def _tile_range(self, inr):
        """[B, gx, gy] footprint mask -> [B, n_tiles]: the tiles between the footprint's first and last cell — what the
        plan kernel's `tile_range` hands to the map kernel (a superset of the tiles the footprint really reaches)."""
        flat = inr.reshape(inr.shape[0], -1)
        some = flat.any(1)
        first = flat.argmax(1) // self.TILE_CELLS
        last = (flat.shape[1] - 1 - flat[:, ::-1].argmax(1)) // self.TILE_CELLS
        t = np.arange(self.n_tiles)[None]
        return some[:, None] & (t >= first[:, None]) & (t <= last[:, None])

    def _apply_bookkept(self, o, passes, tile_oor, own=None):
        """csrc/ipp_cell.cuh::fuse_chain + local_quad restated: the passes of the ENABLED agents in id order; a pass
        whose footprint misses a tile multiplies its cells by exactly 1 (k_out == 1) and is not executed there, and a
        clamp is executed exactly where the value may lie outside [o_min, o_max] — `tile_oor` [B, n_tiles]: the tile's
        range flag to begin with, then "a multiply happened since the last clamp".  passes: (in_rect, k, enabled[B])."""
        tab = self.tab
        kout_one = bool(tab.k_out == F32(1))
        o = o.astype(F32, copy=True)
        oor = tile_oor.copy()
        for inr, k, en in passes:
            en_t = np.broadcast_to(en[:, None], oor.shape)
            clamp_here = self._cells_of_tiles(en_t & oor)
            o = np.where(clamp_here, np.minimum(np.maximum(o, tab.o_min), tab.o_max), o).astype(F32)
            oor = oor & ~en_t
            touch = self._tile_range(inr) if kout_one else np.ones_like(oor)
            mul_here = self._cells_of_tiles(en_t & touch)
            kk = np.where(inr, k, tab.k_out).astype(F32)
            o = np.where(mul_here, o * kk, o).astype(F32)
            oor = oor | (en_t & touch)
        if own is not None:
            inr, k = own
            oc = np.minimum(np.maximum(o, tab.o_min), tab.o_max)
            o = np.where(inr, oc * k, o).astype(F32)
        return o

    def _cells_of_quads(self, qmask):
        """[..., n_quads] -> [..., gx, gy] (every cell of a selected quad)."""
        n_cells = self.tab.gx * self.tab.gy
        cells = np.repeat(qmask, 4, axis=-1)[..., :n_cells]
        return cells.reshape(cells.shape[:-1] + (self.tab.gx, self.tab.gy))

    def step(self, actions=None):
        A = self.A
        comm = self.comm()
        prev = [self._k_of(self.pos[:, j], j, self.t) for j in range(A)]
        new_pos, masks, acts = self._choose_and_move(actions)
        new = [self._k_of(new_pos[:, i], i, self.t + 1) for i in range(A)]
        last = self.glob_o
        # global map: every cell is clamped first (the reward needs the clamped odds anyway), then the bookkept chain
        all_on = np.ones(self.B, bool)
        g0 = np.minimum(np.maximum(last, self.tab.o_min), self.tab.o_max).astype(F32)
        self.glob_o = self._apply_bookkept(g0, [(prev[j][0], prev[j][1], all_on) for j in range(A)],
                                           np.zeros((self.B, self.n_tiles), bool))
        rel, ab = self._reward(last, self.glob_o)
        kout_one = bool(self.tab.k_out == F32(1))
        quads_per_tile = self.TILE_CELLS // 4
        for i in range(A):
            passes = [(prev[j][0], prev[j][1], comm[:, i, j]) for j in range(A) if j != i]
            dense = self._apply_bookkept(self.local_o[:, i], passes, self.flags[:, i], own=(new[i][0], new[i][1]))
            en = np.zeros(self.B, bool)
            touched = new[i][0].copy()
            for j in range(A):
                if j != i:
                    en |= comm[:, i, j]
                    touched |= prev[j][0] & comm[:, i, j][:, None, None]
            tq = self._quads(touched)                                            # [B, n_quads]
            all_tile = en[:, None] & (self.flags[:, i] | (not kout_one))         # [B, n_tiles]
            all_q = np.repeat(all_tile, quads_per_tile, axis=1)[:, : tq.shape[1]]
            proc_q = tq | all_q
            proc = self._cells_of_quads(proc_q)
            self.local_o[:, i] = np.where(proc, dense, self.local_o[:, i])
            bad = self._tile_any(self._out_of_range(self.local_o[:, i]) & proc)
            self.flags[:, i] = bad | (self.flags[:, i] & ~en[:, None])
            tile_proc = np.zeros((self.B, self.n_tiles), bool)
            tile_touch = np.zeros((self.B, self.n_tiles), bool)
            for tl in range(self.n_tiles):
                sl = slice(tl * quads_per_tile, (tl + 1) * quads_per_tile)
                tile_proc[:, tl] = proc_q[:, sl].any(1)
                tile_touch[:, tl] = tq[:, sl].any(1)
            self.processed_pairs += int(tile_proc.sum())
            self.touched_pairs += int(tile_touch.sum())
            self.total_pairs += tile_proc.size
        self.pos = new_pos
        self.t += 1
        return dict(comm=comm, mask=masks, action=acts, reward_rel=rel, reward_abs=ab)

    # ---- the same rule in split mode (ipp_observe / ipp_act) -----------------------------------------------------
    def observe(self):
        A = self.A
        comm = self.comm()
        prev = [self._k_of(self.pos[:, j], j, self.t) for j in range(A)]
        last = self.glob_o
        self.glob_o = self._apply(last, [(prev[j][0], prev[j][1], True) for j in range(A)])
        rel, ab = self._reward(last, self.glob_o)
        kout_one = bool(self.tab.k_out == F32(1))
        quads_per_tile = self.TILE_CELLS // 4
        for i in range(A):
            passes = [(prev[j][0], prev[j][1], True, comm[:, i, j]) for j in range(A) if j != i]
            dense = self._apply(self.local_o[:, i], passes)
            en = np.zeros(self.B, bool)
            touched = np.zeros_like(prev[0][0])
            for j in range(A):
                if j != i:
                    en |= comm[:, i, j]
                    touched |= prev[j][0] & comm[:, i, j][:, None, None]
            tq = self._quads(touched)
            all_tile = en[:, None] & (self.flags[:, i] | (not kout_one))
            proc = self._cells_of_quads(tq | np.repeat(all_tile, quads_per_tile, axis=1)[:, : tq.shape[1]])
            self.local_o[:, i] = np.where(proc, dense, self.local_o[:, i])
            bad = self._tile_any(self._out_of_range(self.local_o[:, i]) & proc)
            self.flags[:, i] = bad | (self.flags[:, i] & ~en[:, None])
        return dict(comm=comm, reward_rel=rel, reward_abs=ab)

    def act(self, actions=None):
        A = self.A
        new_pos, masks, acts = self._choose_and_move(actions)
        for i in range(A):
            inr, k = self._k_of(new_pos[:, i], i, self.t + 1)
            dense = self._apply(self.local_o[:, i], [(inr, k, False)])
            proc = self._cells_of_quads(self._quads(inr))  # own_update_kernel: the quads of the new footprint
            self.local_o[:, i] = np.where(proc, dense, self.local_o[:, i])
            self.flags[:, i] |= self._tile_any(self._out_of_range(self.local_o[:, i]) & proc)
        self.pos = new_pos
        self.t += 1
        return dict(mask=masks, action=acts)


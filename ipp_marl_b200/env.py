"""BatchedIPPEnv — B independent multi-UAV IPP environments stepped by the CUDA kernels.

Host-side mirror of the reference's per-episode objects for a whole batch:
``Mapping`` + ``Simulation`` (mapping/mappings.py:19-30), the ``Agent`` list
(missions/episode_generator.py:90-102) and the per-timestep sequence of
``COMAWrapper.build_observations`` / ``.steps`` (coma_wrapper.py:37-183), with the state kept
resident in HBM between steps.  PyTorch is only the allocator / stream provider here; every
computation goes through the C ABI (include/ipp_b200.h).  There is no CPU path: constructing
the env without a CUDA device or without the compiled library raises.
"""
import ctypes as C

import torch

from . import _native as N
from .geometry import HostTables, make_config


def default_episode_ids(n_envs, env_id_base=0):
    """Episode number of env b of a shard that starts at global env index ``env_id_base``: global index + 1
    (SURVEY.md 8d/8e).  Rank r of a sharded job passes ``env_id_base = r * n_envs``; all random streams are keyed on the
    episode number, so the results do not depend on how the batch is cut into shards."""
    return torch.arange(int(n_envs), dtype=torch.int64) + (int(env_id_base) + 1)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class BatchedIPPEnv:
    def __init__(self, params, n_envs, device=None, env_id_base=0):
        if not torch.cuda.is_available():
            raise N.IppError("BatchedIPPEnv needs a CUDA device (no CPU fallback exists)")
        self.lib = N.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.tables = HostTables(params)
        self.B = int(n_envs)
        self.A = self.tables.n_agents
        self.T = self.tables.budget + 1
        self.env_id_base = int(env_id_base)
        t = self.tables
        self.cfg = make_config(t, self.B)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib, None, self.lib.ipp_create(C.byref(self.cfg), C.byref(self._h)), "ipp_create")
        S = t.map_stride
        dev = self.device
        self._local = torch.empty((self.B, self.A, S), dtype=torch.float32, device=dev)
        self._glob = torch.empty((self.B, S), dtype=torch.float32, device=dev)
        self._gt = torch.zeros((self.B, t.gt_stride), dtype=torch.uint8, device=dev)
        self.episodes = torch.zeros((self.B,), dtype=torch.int32, device=dev)  # bit pattern = uint32
        self._codes = torch.zeros((self.B, 2, t.code_stride), dtype=torch.uint8, device=dev)
        self._flags = torch.full((self.B, int(self.cfg.n_seg), 8), -1, dtype=torch.int32, device=dev)
        self.positions = torch.zeros((self.T + 1, self.B, self.A, 3), dtype=torch.int32, device=dev)
        # rewards + chosen actions in ONE block (ipp_step_host returns them to the host with a single copy)
        self._results = torch.zeros((self.B * (2 + self.A),), dtype=torch.int32, device=dev)
        self.actions = self._results[2 * self.B:].view(self.B, self.A)
        self.masks = torch.zeros((self.B, self.A), dtype=torch.uint8, device=dev)
        self.comm = torch.zeros((self.B, self.A), dtype=torch.uint8, device=dev)
        self.reward_rel = self._results[: self.B].view(torch.float32)
        self.reward_abs = self._results[self.B: 2 * self.B].view(torch.float32)
        self.stuck = torch.zeros((self.B,), dtype=torch.uint8, device=dev)
        self._state = N.IppState(_ptr(self._local), _ptr(self._glob), _ptr(self._gt), _ptr(self.episodes),
                                 _ptr(self._codes), _ptr(self._flags))
        self.t = 0
        self._observed = False
        self._folded = False
        self._ig_actions = None
        self._ios = {}
        self._hist = None
        self._default_episodes = False

    # ---- views in the reference's array convention: [.., gx, gy], first axis = world x ----------
    def _view(self, flat):
        t = self.tables
        return flat[..., : t.n_cells].unflatten(-1, (t.gx, t.gy))

    # The resident state holds ODDS (include/ipp_b200.h); the reference-facing views below are probabilities,
    # exported by a kernel into fresh tensors on every access.
    def _export(self, local):
        out = torch.empty_like(self._local if local else self._glob)
        rc = self.lib.ipp_export_beliefs(self._h, C.byref(self._state), _ptr(out) if local else C.c_void_p(0),
                                         C.c_void_p(0) if local else _ptr(out), self._stream())
        N.check(self.lib, self._h, rc, "ipp_export_beliefs")
        return self._view(out)

    @property
    def local_maps(self):
        """Agent.local_map of every (env, agent): probabilities [B, A, gx, gy]."""
        return self._export(True)

    @property
    def global_map(self):
        """Accumulated global map of every env: probabilities [B, gx, gy]."""
        return self._export(False)

    @property
    def local_odds(self):
        """The resident state itself: odds p/(1-p) [B, A, gx, gy] (a view, no copy)."""
        return self._view(self._local)

    @property
    def global_odds(self):
        return self._view(self._glob)

    @property
    def ground_truth(self):
        return self._view(self._gt)

    @property
    def pos(self):
        return self.positions[self.t]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    VARIANTS = {"direct": 0, "tma": 1}

    def set_step_variant(self, name):
        """Pick the map-kernel implementation ("direct" | "tma"); both give identical results."""
        rc = self.lib.ipp_set_step_variant(self._h, self.VARIANTS[name])
        N.check(self.lib, self._h, rc, "ipp_set_step_variant(%s)" % name)

    @property
    def step_variant(self):
        v = self.lib.ipp_get_step_variant(self._h)
        return {0: "direct", 1: "tma"}[v]

    def close(self):
        if self._h:
            self.lib.ipp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- episode reset ------------------------------------------------------------------------
    def reset(self, episodes=None):
        """Start episode ``episodes[b]`` in env b (default: global env index + 1)."""
        self._set_episodes(episodes)
        self.t = 0
        self._observed = False
        self._folded = False
        rc = self.lib.ipp_reset(self._h, C.byref(self._state), _ptr(self.positions[0]), self._stream())
        N.check(self.lib, self._h, rc, "ipp_reset")

    def _set_episodes(self, episodes):
        """Episode number of every env (device buffer read by ipp_reset).  The default numbering (global env index + 1,
        SURVEY.md 8d/8e) is uploaded once; a repeated default reset costs no host-to-device copy."""
        if episodes is None:
            if self._default_episodes:
                return
            ep = default_episode_ids(self.B, self.env_id_base)
            self._default_episodes = True
        else:
            if torch.is_tensor(episodes) and episodes.is_cuda:
                ep = episodes.to(torch.int64).reshape(self.B)
            else:
                ep = torch.as_tensor(episodes, dtype=torch.int64).reshape(self.B)
            self._default_episodes = False
        self.episodes.copy_(torch.where(ep >= 2**31, ep - 2**32, ep).to(torch.int32), non_blocking=True)

    def _io(self, actions, probs, greedy):
        io = N.IppStepIO()
        io.pos_in = _ptr(self.positions[self.t])
        io.pos_out = _ptr(self.positions[self.t + 1]) if self.t < self.T else C.c_void_p(0)
        io.actions_in = _ptr(actions)
        io.probs_in = _ptr(probs)
        io.greedy = 1 if greedy else 0
        io.actions_out = _ptr(self.actions)
        io.mask_out = _ptr(self.masks)
        io.comm_out = _ptr(self.comm)
        io.reward_rel = _ptr(self.reward_rel)
        io.reward_abs = _ptr(self.reward_abs)
        io.stuck_out = _ptr(self.stuck)
        return io

    def _prep(self, actions, probs):
        if actions is not None:
            actions = torch.as_tensor(actions, device=self.device).to(torch.int32).reshape(self.B, self.A).contiguous()
        if probs is not None:
            probs = torch.as_tensor(probs, device=self.device).to(torch.float32).reshape(self.B, self.A, 6).contiguous()
        return actions, probs

    # ---- one fused timestep -------------------------------------------------------------------
    def step(self, actions=None, probs=None, greedy=False, _phase_hook=None):
        """Whole timestep in one pass over the maps.  Returns (reward_rel, reward_abs, done).

        ``_phase_hook(phase)`` (bench.py only) is called between the two launches of the step so
        that the map kernel can be bracketed with CUDA events."""
        if self.t >= self.T:
            raise N.IppError("episode finished: call reset()")
        if actions is None and probs is None and not greedy:
            io = self._io_cached(self.t)
        else:
            actions, probs = self._prep(actions, probs)
            io = self._io(actions, probs, greedy)
        if _phase_hook is None:
            rc = self.lib.ipp_step(self._h, C.byref(self._state), self.t, C.byref(io), self._stream())
            N.check(self.lib, self._h, rc, "ipp_step")
        else:
            for phase in (1, 2):
                _phase_hook(phase, True)
                rc = self.lib.ipp_step_phases(self._h, C.byref(self._state), self.t, C.byref(io), phase, self._stream())
                N.check(self.lib, self._h, rc, "ipp_step_phases")
                _phase_hook(phase, False)
        done = self.t == self.tables.budget  # coma_wrapper.py:163-164
        self.t += 1
        return self.reward_rel, self.reward_abs, done

    # ---- a whole episode (or any run of timesteps) as ONE launch -------------------------------
    def run_steps(self, n_steps=None, reset=False, episodes=None):
        """``n_steps`` timesteps of the uniform masked policy from the current timestep — preceded by the episode reset
        when ``reset`` — as ONE CUDA-graph launch (C ABI ``ipp_run_steps``; captured once, then replayed).  Default:
        the rest of the episode.  The results of every step are kept: ``reward_hist`` [T, 2, B] (relative, absolute),
        ``action_hist`` / ``mask_hist`` [T, B, A]; ``reward_rel`` / ``reward_abs`` / ``actions`` / ``masks`` keep their
        meaning "of the last ipp_step call" and are NOT written by this path.  Returns ``done``."""
        if reset:
            self._set_episodes(episodes)
            self.t = 0
            self._observed = False
            self._folded = False
        if n_steps is None:
            n_steps = self.T - self.t
        if n_steps < 1 or self.t + n_steps > self.T:
            raise N.IppError("run_steps: %d steps from t = %d do not fit the episode (T = %d)" % (n_steps, self.t, self.T))
        if self._hist is None:
            dev = self.device
            self.reward_hist = torch.zeros((self.T, 2, self.B), dtype=torch.float32, device=dev)
            self.action_hist = torch.zeros((self.T, self.B, self.A), dtype=torch.int32, device=dev)
            self.mask_hist = torch.zeros((self.T, self.B, self.A), dtype=torch.uint8, device=dev)
            ios = (N.IppStepIO * self.T)()
            keep, self.t = self.t, 0
            for t in range(self.T):
                self.t = t
                io = self._io(None, None, False)
                io.reward_rel = _ptr(self.reward_hist[t, 0])
                io.reward_abs = _ptr(self.reward_hist[t, 1])
                io.actions_out = _ptr(self.action_hist[t])
                io.mask_out = _ptr(self.mask_hist[t])
                ios[t] = io
            self.t = keep
            self._hist = ios
        first = C.cast(C.byref(self._hist, self.t * C.sizeof(N.IppStepIO)), C.POINTER(N.IppStepIO))
        rc = self.lib.ipp_run_steps(self._h, C.byref(self._state), 1 if reset else 0, _ptr(self.positions[0]), self.t,
                                    n_steps, first, self._stream())
        N.check(self.lib, self._h, rc, "ipp_run_steps")
        self.t += n_steps
        return self.t == self.T

    def host_results(self):
        """Pinned host buffers (reward_rel [B], reward_abs [B], actions [B, A]) carved out of ONE block, so that
        step_host brings the step's results back with a single device->host copy."""
        blk = torch.zeros((self.B * (2 + self.A),), dtype=torch.int32).pin_memory()
        return (blk[: self.B].view(torch.float32), blk[self.B: 2 * self.B].view(torch.float32),
                blk[2 * self.B:].view(self.B, self.A))

    def step_host(self, probs_host=None, actions_host=None, reward_rel_host=None, reward_abs_host=None,
                  actions_out_host=None):
        """One fused timestep for a policy that lives on the host: ONE C call copies ``probs_host`` [B, A, 6]
        float32 (or ``actions_host`` [B, A] int32) to the device, runs the step and copies rewards / chosen actions
        into the given host tensors (pinned memory keeps everything asynchronous).  Synchronise the current stream
        before reading the outputs.  Returns ``done``."""
        if self.t >= self.T:
            raise N.IppError("episode finished: call reset()")
        for x, shape, dt in ((probs_host, (self.B, self.A, 6), torch.float32), (actions_host, (self.B, self.A), torch.int32),
                             (reward_rel_host, (self.B,), torch.float32), (reward_abs_host, (self.B,), torch.float32),
                             (actions_out_host, (self.B, self.A), torch.int32)):
            if x is not None and (x.is_cuda or x.dtype != dt or tuple(x.shape) != shape or not x.is_contiguous()):
                raise N.IppError("step_host: host tensors must be contiguous CPU tensors of shape %s, %s" % (shape, dt))
        io = self._io_cached(self.t)
        rc = self.lib.ipp_step_host(self._h, C.byref(self._state), self.t, C.byref(io), _ptr(probs_host),
                                    _ptr(actions_host), _ptr(reward_rel_host), _ptr(reward_abs_host),
                                    _ptr(actions_out_host), self._stream())
        N.check(self.lib, self._h, rc, "ipp_step_host")
        done = self.t == self.tables.budget
        self.t += 1
        return done

    def _io_cached(self, t):
        """ipp_step_io of timestep t without injected inputs (built once: the pointers never change)."""
        io = self._ios.get(t)
        if io is None:
            keep, self.t = self.t, t
            io = self._io(None, None, False)
            self.t = keep
            self._ios[t] = io
        return io

    # ---- the same timestep split around a policy network --------------------------------------
    def observe(self, final=False):
        """Fuse local + global maps and compute the reward (everything before the actor forward).

        ``final=True`` is the fold of the LAST measurements after the episode's last move (IG_baseline.py:171-175
        evaluates its metrics on that map); no act() can follow it."""
        if self.t >= self.T and (not final or self._folded):
            raise N.IppError("episode finished: call reset()")
        io = self._io(None, None, False)
        rc = self.lib.ipp_observe(self._h, C.byref(self._state), self.t, C.byref(io), self._stream())
        N.check(self.lib, self._h, rc, "ipp_observe")
        self._observed = True
        if self.t == self.T:
            self._folded = True  # terminal fold done: only reset() may follow
        return self.reward_rel, self.reward_abs

    def act(self, actions=None, probs=None, greedy=False):
        """Masks, action choice, moves and the measurement at the new positions."""
        if not self._observed or self.t >= self.T:
            raise N.IppError("act() must follow observe() in the same timestep")
        actions, probs = self._prep(actions, probs)
        io = self._io(actions, probs, greedy)
        rc = self.lib.ipp_act(self._h, C.byref(self._state), self.t, C.byref(io), self._stream())
        N.check(self.lib, self._h, rc, "ipp_act")
        self._observed = False
        done = self.t == self.tables.budget
        self.t += 1
        return self.actions, done

    # ---- network-input features (split mode) --------------------------------------------------
    def features_actor(self, out=None):
        """Actor observations [B, A, px, py, 7] of the current timestep (call after observe())."""
        if not self._observed:
            raise N.IppError("features_actor() must follow observe() in the same timestep")
        t = self.tables
        if out is None:
            out = torch.empty((self.B, self.A, t.px, t.py, 7), dtype=torch.float32, device=self.device)
        io = self._io(None, None, False)
        rc = self.lib.ipp_features_actor(self._h, C.byref(self._state), self.t, C.byref(io), _ptr(out), self._stream())
        N.check(self.lib, self._h, rc, "ipp_features_actor")
        return out

    def features_critic(self, obs, out=None):
        """Critic states [B, A, px, py, 12] of the timestep that act() just finished."""
        t = self.tables
        if out is None:
            out = torch.empty((self.B, self.A, t.px, t.py, 12), dtype=torch.float32, device=self.device)
        rc = self.lib.ipp_features_critic(self._h, C.byref(self._state), self.t - 1, _ptr(self.positions[self.t - 1]),
                                          _ptr(self.actions), _ptr(obs.contiguous()), _ptr(out), self._stream())
        N.check(self.lib, self._h, rc, "ipp_features_critic")
        return out

    # ---- IG-greedy planner + evaluation metrics --------------------
    def ig_plan(self, communication=True, return_scores=False):
        """IG_baseline.py:127-135,222-325 for the whole batch (call after observe(); feed the result to
        act(actions=...)).  Returns actions [B, A] int32, with ``return_scores`` also (gains, utilities)
        [B, A, 6] float64."""
        if not self._observed or self.t >= self.T:
            raise N.IppError("ig_plan() must follow observe() in the same timestep")
        if self._ig_actions is None:
            self._ig_actions = torch.empty((self.B, self.A), dtype=torch.int32, device=self.device)
        gains = util = None
        if return_scores:
            gains = torch.empty((self.B, self.A, 6), dtype=torch.float64, device=self.device)
            util = torch.empty_like(gains)
        rc = self.lib.ipp_ig_plan(self._h, C.byref(self._state), _ptr(self.positions[self.t]),
                                  1 if communication else 0, _ptr(self._ig_actions), _ptr(self.masks), _ptr(gains),
                                  _ptr(util), self._stream())
        N.check(self.lib, self._h, rc, "ipp_ig_plan")
        return (self._ig_actions, gains, util) if return_scores else self._ig_actions

    def eval_metrics(self):
        """(masked entropy, F1 of class 1) of the accumulated global map, float64 [B] each
        (IG_baseline.py:81-100,191-210)."""
        ent = torch.empty((self.B,), dtype=torch.float64, device=self.device)
        f1 = torch.empty_like(ent)
        rc = self.lib.ipp_eval_metrics(self._h, C.byref(self._state), _ptr(ent), _ptr(f1), self._stream())
        N.check(self.lib, self._h, rc, "ipp_eval_metrics")
        return ent, f1

    # ---- sizes for the roofline (SURVEY.md section 8d contract figure) --------------------------
    def algorithmic_bytes_per_env_step(self):
        t = self.tables
        return t.n_cells * (2 * 4 * (self.A + 1) + 1)

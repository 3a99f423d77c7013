// Hot-path kernels of the batched IPP environment (sm_100a).
//
//   plan_kernel          per env: comm matrix, sequential masks / action choice / moves, and the
//                        measurement codes of the footprints at the new positions (rect-sparse)
//   step_direct_kernel   direct-load map kernel: fuse (local + global) + own update + reward sums
//   reward_finalize      per-env reward from per-chunk partial sums (only when an env spans >1 chunk)
//   own_update_kernel    own measurement update of the local maps (split observe/act mode)
//   reset_prep / reset_fill   episode reset (MT19937-compatible start positions + ground truth)
//
// Arithmetic specification: oracle/kernel_model.py (bit-exact for belief maps).
#include <cstdlib>

#include "ipp_cell.cuh"
#include "ipp_launch.h"

namespace ipp {

// =================================================================================================
// plan: agent/communication_log.py:39-58, agent/action_space.py:56-70,211-223,328-344,
//       sensors/cameras.py:46-79 + mapping/simulations.py:42-65 (measurement at the new position)
// =================================================================================================
__device__ __forceinline__ uint32_t bounds_mask(const ipp_config& c, const int32_t* p) {
  uint32_t m = 0x3Fu;
  if (p[2] == c.max_altitude) m &= ~1u;
  if (p[2] == c.min_altitude) m &= ~(1u << 5);
  if (p[1] == 0) m &= ~(1u << 2);
  if (p[1] == c.y_dim_m) m &= ~(1u << 3);
  if (p[0] == 0) m &= ~(1u << 1);
  if (p[0] == c.x_dim_m) m &= ~(1u << 4);
  return m;
}

__device__ __forceinline__ int32_t kth_set_bit(uint32_t m, int32_t k) {
  for (int32_t a = 0; a < IPP_N_ACTIONS; ++a) {
    if ((m >> a) & 1u) {
      if (k == 0) return a;
      --k;
    }
  }
  return -1;
}

// Masks / action choice / moves of one env by a GROUP of lanes, lane `a` of the group = agent a (lanes a >= A and the
// lanes of envs beyond the batch only keep the warp converged).  Everything that does not depend on the other
// agents' moves runs in parallel — comm-matrix row (agent/communication_log.py:39-58: one uniform draw per ordered
// pair, used or not), bounds mask, the agent's random number, its policy probabilities — then the agents take their
// turns in id order, because agent i sees the NEW positions of agents < i (agent/agent.py:73-104,
// coma_wrapper.py:97-104) and the collision rules are order dependent (action_space.py:328-344).
struct PlanShared {
  int32_t pos[IPP_MAX_AGENTS][3];   // positions at the start of the step
  int32_t nidx[IPP_MAX_AGENTS][2];  // lattice indices after the move
  uint32_t stuck;
};

__device__ void plan_moves_group(const ipp_config& cfg, const int32_t b, const bool env_ok, const int a,
                                 const ipp_step_io& io, const ipp_state& st, const int32_t t, const bool do_comm,
                                 const bool do_move, PlanShared& sh, int32_t (*npos)[3], uint32_t* rec,
                                 const int32_t* __restrict__ gt_params) {
  const int32_t A = cfg.n_agents;
  const bool live = env_ok && a < A;
  int32_t p[3] = {0, 0, 0};
  uint32_t ep = 0;
  if (live) {
    ep = st.episodes[b];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      p[d] = io.pos_in[((int64_t)b * A + a) * 3 + d];
      sh.pos[a][d] = p[d];
    }
    if (a == 0) sh.stuck = 0u;
  }
  __syncwarp();
  // rec (shared memory): this env's facts for the map kernels — EnvMeta field order, then the tile ranges of the
  // communicated and of the new footprints; build_item_records turns them into one ItemRec per map segment
  if (live && do_comm && io.comm_out != nullptr) {
    const uint32_t key = stream_key(cfg.seed, ep, (uint32_t)a, (uint32_t)t, PURPOSE_COMM);
    // fix_range False (communication_log.py:22-31): range index = the episode's first randint(4) = the ground
    // truth's split, stored by reset_prep_kernel
    const int32_t d2_max = cfg.fix_range ? cfg.comm_d2_max : cfg.comm_d2_table[gt_params[(int64_t)b * 4] & 3];
    uint32_t row = 0;
    for (int32_t j = 0; j < A; ++j) {
      const int32_t dx = p[0] - sh.pos[j][0], dy = p[1] - sh.pos[j][1], dz = p[2] - sh.pos[j][2];
      const int32_t d2 = dx * dx + dy * dy + dz * dz;
      const uint32_t n24 = cell_hash(key, (uint32_t)j) >> 8;
      const bool ok = (d2 == 0) || (d2 <= d2_max && n24 >= cfg.fail_thresh24);
      row |= (ok ? 1u : 0u) << j;
    }
    io.comm_out[(int64_t)b * A + a] = (uint8_t)row;
    const uint32_t en = row & ~(1u << a);  // own measurement already used
    uint32_t en4 = 0;
    for (int j = 0; j < A; ++j)
      if ((en >> j) & 1u) en4 |= 0xFu << (4 * j);
    rec[a] = en;
    rec[A + a] = en4;
    rec[2 * A + a] = lut_row(cfg, p);
    rec[4 * A + a] = tile_range(cfg, p);  // tiles the communicated measurement (taken at p) reaches
    if (!do_move) {  // ipp_observe: the map kernel fuses only
      rec[3 * A + a] = 0u;
      rec[5 * A + a] = 1u;  // empty range
    }
  }
  if (!do_move) return;

  // ---- per-agent preparation (independent of the other agents' moves) ----
  const uint32_t bounds = bounds_mask(cfg, p);
  const int32_t ix = p[0] / cfg.spacing, iy = p[1] / cfg.spacing;
  int32_t injected = -1;
  float u = 0.0f;
  float pk[IPP_N_ACTIONS];
#pragma unroll
  for (int k = 0; k < IPP_N_ACTIONS; ++k) pk[k] = 0.0f;
  if (live) {
    if (io.actions_in != nullptr) {
      injected = io.actions_in[(int64_t)b * A + a];
    } else {
      const uint32_t key = stream_key(cfg.seed, ep, (uint32_t)a, (uint32_t)t, PURPOSE_ACTION);
      u = (float)(cell_hash(key, 0u) >> 8) * (1.0f / 16777216.0f);
      if (io.probs_in != nullptr) {
#pragma unroll
        for (int k = 0; k < IPP_N_ACTIONS; ++k) pk[k] = io.probs_in[((int64_t)b * A + a) * IPP_N_ACTIONS + k];
      }
    }
  }
  // ---- the agents' turns, in id order ----
  for (int32_t turn = 0; turn < A; ++turn) {
    if (live && a == turn) {
      uint32_t m = bounds;
      // One already-moved lower-id agent j against a: every rule is guarded by "more than one action still
      // allowed", evaluated before the zeroing (action_space.py:328-344) => order dependent.
      for (int32_t j = 0; j < a; ++j) {
        const int32_t dx = sh.nidx[j][0] - ix, dy = sh.nidx[j][1] - iy;
        if (dx == 0 && dy == 0 && __popc(m) > 1) m &= ~((1u << 0) | (1u << 5));
        if (dx == -1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 1);
        if (dx == 0 && dy == -1 && __popc(m) > 1) m &= ~(1u << 2);
        if (dx == 0 && dy == 1 && __popc(m) > 1) m &= ~(1u << 3);
        if (dx == 1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 4);
      }
      const int32_t cnt = __popc(m);
      int32_t act = -1;
      uint32_t stuck = 0;
      if (io.actions_in != nullptr) {
        // injected action (the reference's policy never emits a masked one): an action that would
        // leave the lattice is turned into "stay" and flagged, so positions always index the tables
        act = injected;
        if (act < -1 || act >= IPP_N_ACTIONS) act = -1;
        if (act >= 0 && !((bounds >> act) & 1u)) {
          act = -1;
          stuck |= 2u;
        }
      } else if (cnt > 0) {
        if (io.probs_in != nullptr) {
          // actor/network.py:63-66,90-96: probs * mask, then multinomial (train) or argmax (eval)
          float w[IPP_N_ACTIONS];
          float total = 0.0f;
#pragma unroll
          for (int32_t k = 0; k < IPP_N_ACTIONS; ++k) {
            w[k] = ((m >> k) & 1u) ? fmaxf(pk[k], 0.0f) : 0.0f;
            total += w[k];
          }
          if (!(total > 0.0f)) {
            act = kth_set_bit(m, min((int32_t)(u * (float)cnt), cnt - 1));
          } else if (io.greedy) {
            float best = -1.0f;
#pragma unroll
            for (int32_t k = 0; k < IPP_N_ACTIONS; ++k)
              if (((m >> k) & 1u) && w[k] > best) { best = w[k]; act = k; }
          } else {
            const float target = u * total;
            float acc = 0.0f;
            bool done = false;
#pragma unroll
            for (int32_t k = 0; k < IPP_N_ACTIONS; ++k) {
              if (!done && w[k] > 0.0f) {
                act = k;  // last positive weight wins if rounding leaves target >= acc at the end
                acc += w[k];
                if (target < acc) done = true;
              }
            }
          }
        } else {
          act = kth_set_bit(m, min((int32_t)(u * (float)cnt), cnt - 1));
        }
      }
      if (cnt == 0) stuck |= 1u;  // the reference raises here; we stay in place and flag
      const int32_t ox = (act == 4) - (act == 1), oy = (act == 3) - (act == 2), oz = (act == 0) - (act == 5);
      sh.nidx[a][0] = ix + ox;
      sh.nidx[a][1] = iy + oy;
      int32_t np[3] = {p[0] + ox * cfg.spacing, p[1] + oy * cfg.spacing, p[2] + oz * cfg.spacing};
#pragma unroll
      for (int32_t d = 0; d < 3; ++d) {
        npos[a][d] = np[d];
        io.pos_out[((int64_t)b * A + a) * 3 + d] = np[d];
      }
      if (io.actions_out != nullptr) io.actions_out[(int64_t)b * A + a] = act;
      if (io.mask_out != nullptr) io.mask_out[(int64_t)b * A + a] = (uint8_t)m;
      if (stuck != 0u) sh.stuck |= stuck;  // only this lane of the env is active in this turn
      if (do_comm) {
        rec[3 * A + a] = lut_row(cfg, np);
        rec[5 * A + a] = tile_range(cfg, np);
      }
    }
    __syncwarp();
  }
  if (live && a == 0 && io.stuck_out != nullptr) io.stuck_out[b] = (uint8_t)sh.stuck;
}

// Write the code byte of quad q (cells 4q..4q+3, first cell in grid row x) for measurement m.
__device__ __forceinline__ void write_code_byte(const ipp_config& cfg, const Meas& m, const int a, const int ap,
                                                const uint8_t* __restrict__ gt, uint8_t* __restrict__ codes,
                                                const int32_t q, const int32_t x, const int32_t row0) {
  const int32_t c0 = q << 2;
  const int32_t y0 = c0 - row0;
  const int32_t left = cfg.gx * cfg.gy - c0;
  const uint32_t valid = left >= 4 ? 0xFu : ((1u << max(left, 0)) - 1u);
  const uint32_t in = rect_mask4(m, x, y0, min(4, cfg.gy - y0)) & valid;
  const uint32_t g4 = *reinterpret_cast<const uint32_t*>(gt + c0);
  codes[(int64_t)q * ap + a] = (uint8_t)(in | ((seen_mask4(m.key, m.thresh, c0, g4) & in) << 4));
}

// Write the code bytes of all A measurements into this env's zeroed code row.  One lane per
// (agent, grid row) pair — the rows of the footprint plus the row just above it, whose last quad may wrap
// into the footprint.  A lane walks the quads that START in its row, left to right, so every quad of a
// footprint is written by exactly one lane: plain byte stores, no atomics, no per-task division.
template <int MAXA>
__device__ __forceinline__ void write_all_codes(const ipp_config& cfg, const Meas* meas, const int A, const int ap,
                                                const uint8_t* __restrict__ gt, uint8_t* __restrict__ codes,
                                                const int lane) {
  int32_t total = 0;
  for (int a = 0; a < A; ++a) {
    const int32_t h = meas[a].xr - meas[a].xl, w = meas[a].yd - meas[a].yu;
    total += (h > 0 && w > 0) ? h + 1 : 0;
  }
  const int32_t gy = cfg.gy;
  for (int32_t idx = lane; idx < total; idx += 32) {
    int a = 0;
    int32_t rr = idx;
    for (; a < A; ++a) {  // which agent does this (agent, row) pair belong to
      const int32_t h = meas[a].xr - meas[a].xl, w = meas[a].yd - meas[a].yu;
      const int32_t n = (h > 0 && w > 0) ? h + 1 : 0;
      if (rr < n) break;
      rr -= n;
    }
    const Meas m = meas[a];
    const int32_t h = m.xr - m.xl;
    const int32_t x = m.xl - 1 + rr;  // this lane owns the quads whose first cell lies in grid row x
    if (x < 0) continue;
    const int32_t row0 = x * gy, row1 = row0 + gy;
    const int32_t q_first = (row0 + 3) >> 2, q_last = (row1 - 1) >> 2;
    int32_t qa = q_last + 1, qb = q_last - 1;  // run of quads overlapping the footprint cells of row x (empty if rr == 0)
    if (rr >= 1) {
      qa = max(q_first, (row0 + m.yu) >> 2);
      qb = min(q_last, (row0 + m.yd - 1) >> 2);
    }
    if (qa <= qb) write_code_byte(cfg, m, a, ap, gt, codes, qa, x, row0);
    for (int32_t q = qa + 1; q < qb; ++q) {  // strictly inside the run: all 4 cells are footprint cells of row x
      const uint32_t g4 = *reinterpret_cast<const uint32_t*>(gt + (q << 2));
      codes[(int64_t)q * ap + a] = (uint8_t)(0xFu | (seen_mask4(m.key, m.thresh, q << 2, g4) << 4));
    }
    if (qb > qa) write_code_byte(cfg, m, a, ap, gt, codes, qb, x, row0);
    // the last quad of the row may wrap into row x+1: footprint cells there have y < n1
    const int32_t n1 = (q_last << 2) + 4 - row1;
    if (rr < h && n1 > 0 && m.yu < n1 && qb < q_last) write_code_byte(cfg, m, a, ap, gt, codes, q_last, x, row0);
  }
}

// bits of the tiles [t0, t0 + nt) of a map segment that lie inside the packed tile range r (bit 0 = tile t0; nt <= 32)
__device__ __forceinline__ uint32_t range_bits(uint32_t r, int32_t t0, int32_t nt) {
  const int32_t lo = max((int32_t)(r & 0xFFFFu), t0) - t0, hi = min((int32_t)(r >> 16), t0 + nt - 1) - t0;
  return hi < lo ? 0u : ((2u << hi) - (1u << lo));
}

// One warp writes the ItemRec of every segment of one env (ipp_cell.cuh).  Lane i < A works on agent / local map i
// with whole-segment tile masks (a footprint's tiles are one run of bits), lane t < ITEM_TILES then assembles the
// facts of tile t from the A masks.
// rec = the env's words from plan_moves_group: comm | comm4 | lut_prev | lut_next | rng_prev | rng_next (A each).
__device__ void build_item_records(const ipp_config& cfg, const ipp_state& st, const int32_t b, const int lane,
                                   const uint32_t* rec, const bool have_next, uint32_t* __restrict__ step_meta) {
  const int A = cfg.n_agents;
  const int rw = rec_words(A);
  const bool kout_one = (cfg.k_out == 1.0f);
  const int32_t n_quads = (cfg.gx * cfg.gy + 3) >> 2;
  const bool agent = lane < A;
  const uint32_t comm_i = agent ? rec[lane] : 0u;
  const uint32_t rp = agent ? rec[4 * A + lane] : 1u, rn = (agent && have_next) ? rec[5 * A + lane] : 1u;
  for (int32_t c = 0; c < cfg.n_seg; ++c) {
    const int32_t nq = min(IPP_FLAG_QUADS, n_quads - c * IPP_FLAG_QUADS);
    const int32_t nt = (nq + 31) >> 5;
    const uint32_t tiles = nt >= 32 ? 0xFFFFFFFFu : ((1u << nt) - 1u);
    const uint32_t flags = agent ? st.map_flags[((int64_t)b * cfg.n_seg + c) * 8 + lane] & tiles : 0u;
    // lane i: tiles agent i's communicated footprint reaches; tiles that must be clamped as a whole; tiles with work
    uint32_t touch = agent ? (kout_one ? range_bits(rp, c * ITEM_TILES, nt) : tiles) : 0u;
    const uint32_t dirty = comm_i != 0u ? (kout_one ? flags : tiles) : 0u;
    uint32_t need = dirty | range_bits(rn, c * ITEM_TILES, nt);
    uint32_t tch = 0, drt = 0;  // lane t: bit j / i of tile t
    for (int j = 0; j < A; ++j) {
      const uint32_t touch_j = __shfl_sync(0xFFFFFFFFu, touch, j);
      const uint32_t dirty_j = __shfl_sync(0xFFFFFFFFu, dirty, j);
      if ((comm_i >> j) & 1u) need |= touch_j;
      tch |= ((touch_j >> lane) & 1u) << j;
      drt |= ((dirty_j >> lane) & 1u) << j;
    }
    uint32_t* out = step_meta + ((int64_t)b * cfg.n_seg + c) * rw;
    for (int w = lane; w < 4 * A; w += 32) out[w] = rec[w];
    if (agent) {
      out[4 * A + lane] = need;
      out[5 * A + lane] = flags;
    }
    // tile bytes, two tiles per word: lanes 0, 2, 4 .. write (own | neighbour << 16)
    const uint32_t tb = tch | (drt << 8);
    const uint32_t nb = __shfl_down_sync(0xFFFFFFFFu, tb, 1);
    if (lane < ITEM_TILES && (lane & 1) == 0) out[6 * A + (lane >> 1)] = tb | (nb << 16);
    if (lane < rw - (6 * A + ITEM_TILES / 2)) out[6 * A + ITEM_TILES / 2 + lane] = 0u;  // the record's padding words
  }
}

constexpr int PLAN_WARPS = 16;     // warps per block: one env per warp and round in phase 2
constexpr int PLAN_ENVS = 16;      // envs per block (default)
constexpr int PLAN_MAX_ENVS = 32;  // most envs per block (the launcher picks: see launch_plan)

// Phase 1: the first warp(s) plan the moves of the block's 16 envs, one lane per (env, agent): the parts that do not
// depend on the other agents' moves in parallel, then the agents' turns in id order (plan_moves_group).
// Phase 2: one warp per env builds the env's new code row.  When it fits (stage != 0) the row is assembled in
// SHARED memory next to a staged copy of the ground truth and written out with coalesced 16-byte stores:
// scattering one byte per (quad, agent) straight to global memory costs one L2 write transaction per byte
// (7.4 M per launch at 8192 envs — that, not the arithmetic, bounded earlier versions of this kernel).
__global__ void __launch_bounds__(PLAN_WARPS * 32, 2)  // two 93 KB blocks per SM (more do not fit: measured no gain)
    plan_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const ipp_step_io io, const int32_t t,
                const int32_t do_comm, const int32_t do_move, const int32_t stage, uint32_t* __restrict__ step_meta,
                const int32_t epb, const int32_t* __restrict__ gt_params) {
  extern __shared__ __align__(16) unsigned char plan_smem[];  // [PLAN_WARPS][gt_stride + code_stride] when stage
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t e0 = blockIdx.x * epb;  // epb envs per block, 16 <= epb <= PLAN_MAX_ENVS
  const int32_t n_here = min(epb, cfg.n_envs - e0);
  const int32_t A = cfg.n_agents;
  __shared__ int32_t s_npos[PLAN_MAX_ENVS][IPP_MAX_AGENTS][3];
  __shared__ Meas s_meas[PLAN_WARPS][IPP_MAX_AGENTS];
  __shared__ PlanShared s_plan[PLAN_MAX_ENVS];
  __shared__ uint32_t s_rec[PLAN_MAX_ENVS][6 * IPP_MAX_AGENTS];
  // Phase 1: groups of W lanes (W = power of two >= A), one lane per agent, in the block's first warps
  int W = 1;
  while (W < A) W <<= 1;
  if (warp < (epb * W + 31) / 32) {
    const int e = (warp * 32 + lane) / W, a = lane & (W - 1);
    const bool env_ok = e < n_here;
    const int es = min(e, PLAN_MAX_ENVS - 1);
    if (stage & 2) {  // debug: no planning, agents stay
      if (env_ok && a < A)
        for (int d = 0; d < 3; ++d) {
          s_npos[es][a][d] = io.pos_in[((int64_t)(e0 + es) * A + a) * 3 + d];
          io.pos_out[((int64_t)(e0 + es) * A + a) * 3 + d] = s_npos[es][a][d];
        }
    } else {
      plan_moves_group(cfg, e0 + es, env_ok, a, io, st, t, do_comm != 0, do_move != 0, s_plan[es], s_npos[es],
                       s_rec[es], gt_params);
    }
  }
  if (do_comm && !do_move) {  // ipp_observe: the map kernel will run: one ItemRec per (env, map segment)
    __syncthreads();
    for (int32_t e = warp; e < n_here; e += PLAN_WARPS)
      build_item_records(cfg, st, e0 + e, lane, s_rec[e], false, step_meta);
  }
  if (!do_move) return;
  if (stage & 4) return;  // debug: no code generation
  const int ap = A <= 4 ? 4 : 8;
  const int32_t n16 = cfg.code_stride >> 4;
  unsigned char* my_smem = plan_smem + (size_t)warp * (cfg.gt_stride + cfg.code_stride);
  // work that does not depend on the moves: stage the ground truth, clear the row (env of the first round)
  if ((stage & 1) && warp < n_here) {
    const uint4* src = reinterpret_cast<const uint4*>(st.ground_truth + (int64_t)(e0 + warp) * cfg.gt_stride);
    uint4* dst = reinterpret_cast<uint4*>(my_smem);
    for (int32_t i = lane; i < (cfg.gt_stride >> 4); i += 32) dst[i] = src[i];
  }
  __syncthreads();  // new positions visible to the whole block
  for (int32_t e = warp; e < n_here; e += PLAN_WARPS) {
    const int32_t b = e0 + e;
    // code row of the measurements taken after the move: half (t+1)&1 of the ping-pong buffer
    uint4* grow = reinterpret_cast<uint4*>(st.meas_codes + code_row_offset(cfg, t + 1, b));
    const uint8_t* gt = st.ground_truth + (int64_t)b * cfg.gt_stride;
    uint8_t* row = reinterpret_cast<uint8_t*>(grow);
    if (stage & 1) {
      if (e != warp) {  // later rounds (more envs than warps): stage this env's ground truth now
        const uint4* src = reinterpret_cast<const uint4*>(gt);
        uint4* dst = reinterpret_cast<uint4*>(my_smem);
        for (int32_t i = lane; i < (cfg.gt_stride >> 4); i += 32) dst[i] = src[i];
      }
      gt = my_smem;
      row = my_smem + cfg.gt_stride;
    }
    uint4* rz = reinterpret_cast<uint4*>(row);
    for (int32_t i = lane; i < n16; i += 32) rz[i] = make_uint4(0u, 0u, 0u, 0u);
    if (lane < A) s_meas[warp][lane] = make_meas(cfg, s_npos[e][lane], st.episodes[b], (uint32_t)lane, (uint32_t)t + 1u);
    if (do_comm) build_item_records(cfg, st, b, lane, s_rec[e], true, step_meta);  // ItemRec per (env, map segment)
    __syncwarp();  // zeroed row, staged ground truth and s_meas visible to all lanes
    write_all_codes<IPP_MAX_AGENTS>(cfg, s_meas[warp], A, ap, gt, row, lane);
    __syncwarp();
    if (stage & 1)
      for (int32_t i = lane; i < n16; i += 32) grow[i] = rz[i];
    __syncwarp();  // shared buffers are reused by the next env of this warp
  }
}

// =================================================================================================
// map kernel, direct-load variant (per-quad arithmetic: ipp_cell.cuh).
// One 640-thread block per (env, segment of IPP_FLAG_QUADS quads), one quad per thread, everything in registers:
// the loads that depend on nothing (code words, global quad) are issued first, then — once the env's comm bits
// and range flags are known — the local quads that some footprint reaches; (quad, local map) pairs without
// work are neither loaded nor stored, so this variant moves fewer HBM bytes than the dense contract figure.
// =================================================================================================
constexpr int DIRECT_THREADS = IPP_FLAG_QUADS;

template <int A, bool DO_OWN>
__global__ void __launch_bounds__(DIRECT_THREADS, A <= 4 ? 2 : 1)
    step_direct_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut,
                       const uint32_t* __restrict__ step_meta, const int32_t t, float* __restrict__ reward_rel,
                       float* __restrict__ reward_abs, double* __restrict__ partials, const int32_t n_chunks) {
  const int32_t b = blockIdx.x / n_chunks;
  const int32_t chunk = blockIdx.x - b * n_chunks;
  const int32_t tid = threadIdx.x;
  __shared__ EnvMeta<A> s_meta;
  __shared__ uint32_t s_dirty[A], s_bad[A];
  __shared__ double s_red[2][DIRECT_THREADS / 32];

  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  const bool have = chunk * IPP_FLAG_QUADS + tid < n_quads;
  // threads beyond the map's last quad run on a copy of that quad and do not store (no divergent regions)
  const int32_t q = min(chunk * IPP_FLAG_QUADS + tid, n_quads - 1);
  const int64_t c0 = (int64_t)q << 2;
  float* glob = st.global_map + (int64_t)b * stride + c0;
  float* loc = st.local_maps + (int64_t)b * A * stride + c0;

  // ---- loads that depend on nothing ----
  const CodeWord<A> cw = load_code<A>(st.meas_codes + code_row_offset(cfg, t, b), q);
  CodeWord<A> nw;
#pragma unroll
  for (int w = 0; w < CodeWord<A>::WORDS; ++w) nw.w[w] = 0u;
  if (DO_OWN) nw = load_code<A>(st.meas_codes + code_row_offset(cfg, t + 1, b), q);
  const float4 g4 = __ldcs(reinterpret_cast<const float4*>(glob));
  if (tid < 4 * A) {  // the env's facts from the item's record (comm bits, LUT rows)
    reinterpret_cast<uint32_t*>(&s_meta)[tid] = step_meta[((int64_t)b * cfg.n_seg + chunk) * rec_words(A) + tid];
  } else if (tid < 5 * A) {
    s_dirty[tid - 4 * A] = st.map_flags[((int64_t)b * cfg.n_seg + chunk) * 8 + (tid - 4 * A)];
    s_bad[tid - 4 * A] = 0u;
  }
  __syncthreads();

  // ---- local quads with work: cells of an enabled fuse pass or of the own new footprint; every quad if the
  //      map may hold out-of-range odds and a fuse pass (= whole-map clamp) runs, or if k_out != 1 ----
  const bool kout_one = (cfg.k_out == 1.0f);
  uint32_t in_prev = 0;
#pragma unroll
  for (int j = 0; j < A; ++j) in_prev |= (cw.byte(j) & 0xFu) << (4 * j);
  float4 l4[A];
  uint32_t mine = 0, pre = 0;
#pragma unroll
  for (int i = 0; i < A; ++i) {
    const uint32_t en = s_meta.comm[i];
    const bool dirty = en != 0u && ((s_dirty[i] >> (tid >> 5)) & 1u) != 0u;  // warp = tile
    const bool all = dirty || (en != 0u && !kout_one);
    pre |= (dirty ? 1u : 0u) << i;
    if (have && (all || ((in_prev & s_meta.comm4[i]) | (nw.byte(i) & 0xFu)) != 0u)) {
      mine |= 1u << i;
      l4[i] = __ldcs(reinterpret_cast<const float4*>(loc + (int64_t)i * stride));
    }
  }

  // ---- global map + reward terms ----
  double s1, s2;
  {
    float f1 = 0.0f, f2 = 0.0f;
    const float4 gn = global_quad<A>(cfg, s_meta, cw, lut, g4, have ? valid_mask4((int32_t)c0, n_cells) : 0u,
                                     (1u << A) - 1u, f1, f2);
    if (have && quad_changed(gn, g4)) __stcs(reinterpret_cast<float4*>(glob), gn);
    s1 = (double)warp_sum_f(f1);  // the same float32 sum over a warp's 32 quads as in the TMA kernel
    s2 = (double)warp_sum_f(f2);
  }
  // ---- local maps ----
#pragma unroll
  for (int i = 0; i < A; ++i) {
    if (!((mine >> i) & 1u)) continue;
    if (local_quad<A, DO_OWN>(cfg, s_meta, s_meta.comm[i], s_meta.comm[i], ((pre >> i) & 1u) != 0u, i, cw, nw.byte(i),
                              lut, l4[i]))
      atomicOr(&s_bad[i], 1u << (tid >> 5));
    __stcs(reinterpret_cast<float4*>(loc + (int64_t)i * stride), l4[i]);
  }

  // ---- per-env reward: warp shuffle + shared-memory reduction of the two float64 sums (fixed order) ----
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = s1;
    s_red[1][tid >> 5] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int w = 0; w < DIRECT_THREADS / 32; ++w) {
      t1 += s_red[0][w];
      t2 += s_red[1][w];
    }
    if (n_chunks == 1) {
      write_rewards(reward_rel, reward_abs, b, t1, t2, n_cells);
    } else {
      partials[((int64_t)b * n_chunks + chunk) * 2 + 0] = t1;
      partials[((int64_t)b * n_chunks + chunk) * 2 + 1] = t2;
    }
  } else if (tid >= 32 && tid < 32 + A) {
    // new range flag: some result left the range, or nothing clamped an already flagged map
    const int i = tid - 32;
    // a fuse pass clamped every tile it had to (flagged tiles are processed densely): only this step's results can
    // be out of range; without a fuse pass the old bits stay
    st.map_flags[((int64_t)b * cfg.n_seg + chunk) * 8 + i] = s_bad[i] | (s_meta.comm[i] == 0u ? s_dirty[i] : 0u);
  }
}

__global__ void reward_finalize_kernel(const double* __restrict__ partials, const int32_t n_envs,
                                       const int32_t n_chunks, const int32_t n_cells, float* __restrict__ reward_rel,
                                       float* __restrict__ reward_abs) {
  const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_envs) return;
  double t1 = 0.0, t2 = 0.0;
  for (int32_t c = 0; c < n_chunks; ++c) {
    t1 += partials[((int64_t)b * n_chunks + c) * 2 + 0];
    t2 += partials[((int64_t)b * n_chunks + c) * 2 + 1];
  }
  write_rewards(reward_rel, reward_abs, b, t1, t2, n_cells);
}

// =================================================================================================
// own measurement update of the local maps (ipp_act): mapping/mappings.py:32-61, driven by the codes
// plan_kernel just wrote; only quads inside the footprint are read and written.
// =================================================================================================
template <int A>
__global__ void __launch_bounds__(256)
    own_update_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut,
                      const int32_t* __restrict__ pos_out, const int32_t t, const int32_t blocks_per_env) {
  const int32_t b = blockIdx.x / blocks_per_env;
  const int32_t n_quads = (cfg.gx * cfg.gy + 3) >> 2;
  const uint8_t* code_next = st.meas_codes + code_row_offset(cfg, t + 1, b);
  const int64_t stride = cfg.map_stride;
  __shared__ uint32_t s_row[A];
  if (threadIdx.x < A) s_row[threadIdx.x] = lut_row(cfg, pos_out + ((int64_t)b * A + threadIdx.x) * 3);
  __syncthreads();
  for (int32_t q = (blockIdx.x - b * blocks_per_env) * 256 + threadIdx.x; q < n_quads; q += blocks_per_env * 256) {
    const CodeWord<A> cw = load_code<A>(code_next, q);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const uint32_t byte = cw.byte(i);
      const uint32_t own = byte & 0xFu;
      if (own == 0u) continue;
      float* lp = st.local_maps + ((int64_t)b * A + i) * stride + ((int64_t)q << 2);
      const F4 o = f4_from(*reinterpret_cast<const float4*>(lp));
      const F4 upd = f4_select(own, f4_mul(f4_clamp(o, cfg.o_min, cfg.o_max), f4_from(lut[s_row[i] + byte])), o);
      *reinterpret_cast<float4*>(lp) = f4_to(upd);
      if (f4_out_of_range(upd, cfg.o_min, cfg.o_max))
        atomicOr(&st.map_flags[((int64_t)b * cfg.n_seg + q / IPP_FLAG_QUADS) * 8 + i], 1u << ((q % IPP_FLAG_QUADS) >> 5));
    }
  }
}

// =================================================================================================
// reset: mapping/ground_truths.py:42-56, agent/state_space.py:28-51, mapping/mappings.py:126-132,
// agent/agent.py:44-49.  numpy's legacy RandomState is MT19937 seeded by init_genrand; its
// randint() draws 32-bit words, masks them and rejects values above the range.
// =================================================================================================
constexpr int MT_DRAWS = 40;
constexpr int RESET_PREP_THREADS = 128;

// The first MT_DRAWS outputs of MT19937 need mt[0 .. MT_DRAWS] and mt[397 .. 397 + MT_DRAWS - 1] of the seeded state.
// They live in SHARED memory, one column per thread ([word][thread]: conflict-free) — as per-thread local arrays
// with a run-time index they were the whole cost of this kernel (32 us for 41 k streams).
struct MtStream {
  uint32_t* lo;  // lo[i * RESET_PREP_THREADS] = mt[i],        i = 0 .. MT_DRAWS
  uint32_t* hi;  // hi[i * RESET_PREP_THREADS] = mt[397 + i],  i = 0 .. MT_DRAWS - 1
  int32_t next;
  __device__ void seed(uint32_t s) {
    uint32_t v = s;
    lo[0] = v;
#pragma unroll 4
    for (int32_t i = 1; i <= MT_DRAWS; ++i) {
      v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
      lo[i * RESET_PREP_THREADS] = v;
    }
#pragma unroll 4
    for (int32_t i = MT_DRAWS + 1; i < 397; ++i) v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
#pragma unroll 4
    for (int32_t i = 397; i < 397 + MT_DRAWS; ++i) {
      v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
      hi[(i - 397) * RESET_PREP_THREADS] = v;
    }
    next = 0;
  }
  __device__ uint32_t draw() {
    const int32_t k = min(next, MT_DRAWS - 1);
    ++next;
    const uint32_t a = lo[k * RESET_PREP_THREADS], b = lo[(k + 1) * RESET_PREP_THREADS];
    const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
    uint32_t v = hi[k * RESET_PREP_THREADS] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    v ^= v >> 11;
    v ^= (v << 7) & 0x9D2C5680u;
    v ^= (v << 15) & 0xEFC60000u;
    v ^= v >> 18;
    return v;
  }
  // numpy legacy randint: value in [0, rng] by masked rejection
  __device__ uint32_t bounded(uint32_t rng) {
    if (rng == 0u) return 0u;
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    do {
      v = draw() & mask;
    } while (v > rng && next < MT_DRAWS);
    return min(v, rng);
  }
};

// thread (b, a): a < A -> start position of agent a; a == A -> ground-truth parameters
__global__ void __launch_bounds__(RESET_PREP_THREADS) reset_prep_kernel(const __grid_constant__ ipp_config cfg,
                                                         const uint32_t* __restrict__ episodes,
                                                         int32_t* __restrict__ pos_out,
                                                         int32_t* __restrict__ gt_params,
                                                         uint32_t* __restrict__ flags) {
  __shared__ uint32_t s_mt[(2 * MT_DRAWS + 1) * RESET_PREP_THREADS];
  const int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t A = cfg.n_agents;
  if (idx >= cfg.n_envs * (A + 1)) return;
  const int32_t b = idx / (A + 1), a = idx % (A + 1);
  const uint32_t ep = episodes[b];
  MtStream mt;
  mt.lo = s_mt + threadIdx.x;
  mt.hi = s_mt + (MT_DRAWS + 1) * RESET_PREP_THREADS + threadIdx.x;
  if (a < A) {
    // RandomState(seed = seed * episode * agent_id): state_space.py:29 (numpy rejects seeds >= 2^32;
    // we wrap, documented in DESIGN.md)
    mt.seed((uint32_t)((uint64_t)cfg.seed * (uint64_t)ep * (uint64_t)a));
    const int32_t x = cfg.spacing * (int32_t)mt.bounded((uint32_t)cfg.px - 1u);
    const int32_t y = cfg.spacing * (int32_t)mt.bounded((uint32_t)cfg.py - 1u);
    int32_t* p = pos_out + ((int64_t)b * A + a) * 3;
    p[0] = x;
    p[1] = y;
    p[2] = 15;  // state_space.py:32 hard-codes the start altitude
    // range flags of the local map after the t = 0 measurement (prior odds times k_hi / k_lo of that altitude)
    {
      const int32_t iz = clampi(15 / cfg.spacing - cfg.min_altitude / cfg.spacing, 0, cfg.n_alt - 1);
      const float o = fminf(fmaxf(to_odds(cfg.prior), cfg.o_min), cfg.o_max);
      const bool in_range = o * cfg.k_hi[iz] <= cfg.o_max && o * cfg.k_hi[iz] >= cfg.o_min &&
                            o * cfg.k_lo[iz] <= cfg.o_max && o * cfg.k_lo[iz] >= cfg.o_min &&
                            to_odds(cfg.prior) == o;
      for (int32_t sgm = 0; sgm < cfg.n_seg; ++sgm) flags[((int64_t)b * cfg.n_seg + sgm) * 8 + a] = in_range ? 0u : 0xFFFFFFFFu;
    }
  } else {
    mt.seed(ep);  // np.random.seed(episode): ground_truths.py:43
    const int32_t split = (int32_t)mt.bounded(3u);
    const int32_t pct = 30 + (int32_t)mt.bounded(30u);
    // boundaries of the half plane, in rows (split 0/1) or columns (split 2/3): ground_truths.py:47-56
    const int32_t n = (split < 2) ? cfg.gx : cfg.gy;
    int32_t lo, hi;
    if ((split & 1) == 0) {
      lo = 0;
      hi = (n * pct) / 100;
    } else {
      lo = n - (n * (pct - 1)) / 100;  // int(n*(1-pct)/100) is a NEGATIVE start index in the reference
      hi = n;
    }
    gt_params[(int64_t)b * 4 + 0] = split;
    gt_params[(int64_t)b * 4 + 1] = lo;
    gt_params[(int64_t)b * 4 + 2] = hi;
    gt_params[(int64_t)b * 4 + 3] = pct;
  }
}

// dense write of the ground truth, the prior maps with the t = 0 measurement applied, and the codes of
// that measurement (half 0 of the ping-pong buffer; half 1 is cleared by the first plan_kernel)
template <int A>
__global__ void __launch_bounds__(STEP_THREADS)
    reset_fill_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut,
                      const int32_t* __restrict__ pos, const int32_t* __restrict__ gt_params, const int32_t n_chunks,
                      const int32_t quads_per_chunk) {
  const int32_t b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks, tid = threadIdx.x;
  __shared__ Meas s_m[A];
  __shared__ uint32_t s_row[A];
  const uint32_t ep = st.episodes[b];
  if (tid < A) {
    s_m[tid] = make_meas(cfg, pos + ((int64_t)b * A + tid) * 3, ep, tid, 0u);
    s_row[tid] = lut_row(cfg, pos + ((int64_t)b * A + tid) * 3);
  }
  __syncthreads();
  const int32_t split = gt_params[(int64_t)b * 4 + 0], lo = gt_params[(int64_t)b * 4 + 1],
                hi = gt_params[(int64_t)b * 4 + 2];
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (int32_t)(cfg.map_stride >> 2);
  const int64_t stride = cfg.map_stride;
  const float prior = to_odds(cfg.prior);  // the maps hold odds: prior/(1-prior) in float32
  const F4 o_prior = f4_clamp(f4_splat(prior), cfg.o_min, cfg.o_max);
  constexpr int AP = A <= 4 ? 4 : 8;
  uint8_t* codes = st.meas_codes + code_row_offset(cfg, 0, b);  // half 0
  const int32_t q_end = min((chunk + 1) * quads_per_chunk, n_quads);
  const float4 prior4 = make_float4(prior, prior, prior, prior);
  for (int32_t q = chunk * quads_per_chunk + tid; q < q_end; q += STEP_THREADS) {
    const int32_t c0 = q << 2;
    const int32_t x0 = c0 / cfg.gy, y0 = c0 - x0 * cfg.gy;  // first cell of the quad; it wraps into row x0 + 1 after n0
    const int32_t n0 = min(4, cfg.gy - y0);
    const int32_t left = n_cells - c0;
    const uint32_t valid = left >= 4 ? 0xFu : ((1u << max(left, 0)) - 1u);
    uint32_t g4 = 0;
    {
      int32_t x = x0, y = y0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int32_t v = (split < 2) ? x : y;
        if (((valid >> c) & 1u) && v >= lo && v < hi) g4 |= 1u << (8 * c);
        if (++y == cfg.gy) { y = 0; ++x; }
      }
    }
    *reinterpret_cast<uint32_t*>(st.ground_truth + (int64_t)b * cfg.gt_stride + c0) = g4;
    __stcs(reinterpret_cast<float4*>(st.global_map + (int64_t)b * stride + c0), prior4);
    uint32_t cw[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const Meas& m = s_m[i];
      // most quads lie in no footprint row of agent i: one range test, prior odds, nothing else
      const uint32_t rows = (uint32_t)(m.xr - m.xl);
      const bool near = (uint32_t)(x0 - m.xl) < rows || (n0 < 4 && (uint32_t)(x0 + 1 - m.xl) < rows);
      float4 out = prior4;
      if (near) {
        const uint32_t in = rect_mask4(m, x0, y0, n0) & valid;
        if (in != 0u) {
          const uint32_t byte = in | ((seen_mask4(m.key, m.thresh, c0, g4) & in) << 4);
          cw[i >> 2] |= byte << (8 * (i & 3));
          out = f4_to(f4_select(in, f4_mul(o_prior, f4_from(lut[s_row[i] + byte])), f4_splat(prior)));
        }
      }
      __stcs(reinterpret_cast<float4*>(st.local_maps + ((int64_t)b * A + i) * stride + c0), out);
    }
    if ((int64_t)q * AP < cfg.code_stride) {
      if (AP == 4) reinterpret_cast<uint32_t*>(codes)[q] = cw[0];
      else reinterpret_cast<uint2*>(codes)[q] = make_uint2(cw[0], cw[1]);
    }
  }
}

// =================================================================================================
// export: the stored odds as probabilities, p = o/(1+o) (o < 1) or 1 - 1/(1+o) — IEEE float32 divisions,
// bit-identical to oracle/kernel_model.py::to_p.  This is what Agent.local_map / the accumulated global map
// of the reference hold (agent/agent.py:35, missions/episode_generator.py:47).
// =================================================================================================
__global__ void __launch_bounds__(256) export_beliefs_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                             const int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 o = src[i];
    dst[i] = make_float4(from_odds(o.x), from_odds(o.y), from_odds(o.z), from_odds(o.w));
  }
}

cudaError_t launch_export_beliefs(const float* src, float* dst, int64_t n_floats, cudaStream_t s) {
  const int64_t n4 = n_floats >> 2;  // map_stride is a multiple of 4
  if (n4 == 0) return cudaSuccess;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  export_beliefs_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                                         reinterpret_cast<float4*>(dst), n4);
  return cudaGetLastError();
}

// =================================================================================================
// launchers
// =================================================================================================
#define IPP_DISPATCH_A(A_, CALL)              \
  switch (A_) {                               \
    case 1: { constexpr int kA = 1; CALL; } break; \
    case 2: { constexpr int kA = 2; CALL; } break; \
    case 3: { constexpr int kA = 3; CALL; } break; \
    case 4: { constexpr int kA = 4; CALL; } break; \
    case 5: { constexpr int kA = 5; CALL; } break; \
    case 6: { constexpr int kA = 6; CALL; } break; \
    case 7: { constexpr int kA = 7; CALL; } break; \
    case 8: { constexpr int kA = 8; CALL; } break; \
    default: return cudaErrorInvalidValue;    \
  }

// function attributes of the plan kernel on the current device (idempotent; called from ipp_create so that no
// attribute has to be set while a stream is being captured into a CUDA graph)
cudaError_t configure_plan() {
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  return cudaSuccess;
}

cudaError_t launch_plan(const ipp_config& cfg, const ipp_state& st, const ipp_step_io& io, int32_t t, int do_comm,
                        int do_move, uint32_t* step_meta, const int32_t* gt_params, cudaStream_t s) {
  // ground truth + new code row of one env per warp in shared memory (falls back to global for big grids)
  const int stage_gt = (do_move && (size_t)PLAN_WARPS * (cfg.gt_stride + cfg.code_stride) <= 96 * 1024) ? 1 : 0;
  const size_t smem = stage_gt ? (size_t)PLAN_WARPS * (cfg.gt_stride + cfg.code_stride) : 0;
  int dbg = 0;
#ifdef IPP_PLAN_TIMING_KNOBS  // scripts/plan_bench.py: timing experiments only; not in the product build
  static const int dbg_env = getenv("IPP_PLAN_DEBUG") ? atoi(getenv("IPP_PLAN_DEBUG")) & 6 : 0;
  dbg = dbg_env;
#endif
  {
    cudaError_t e = configure_plan();
    if (e != cudaSuccess) return e;
  }
  // Envs per block: two 80 KB blocks fit an SM, so the GPU runs 2 * n_sm blocks at a time.  When the batch needs
  // more than one such wave of 16-env blocks but fits ONE wave of <= 32-env blocks, use the larger blocks: the
  // sequential phase 1 is then paid once instead of once per wave and no SM idles in a partial second wave.
  static PerDevice slots;
  int& n_slots = slots.cur();
  if (n_slots == 0) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    n_slots = 2 * (n_sm > 0 ? n_sm : 1);
  }
  int epb = PLAN_ENVS;
  if ((cfg.n_envs + PLAN_ENVS - 1) / PLAN_ENVS > n_slots && (cfg.n_envs + n_slots - 1) / n_slots <= PLAN_MAX_ENVS)
    epb = (cfg.n_envs + n_slots - 1) / n_slots;
  // a small batch does not fill the GPU with 16-env blocks: fewer envs per block so that every SM gets work
  if ((cfg.n_envs + PLAN_ENVS - 1) / PLAN_ENVS < n_slots / 2) {
    epb = (cfg.n_envs + n_slots - 1) / n_slots;
    if (epb < 2) epb = 2;
  }
#ifdef IPP_PLAN_TIMING_KNOBS
  static const int epb_env = getenv("IPP_PLAN_EPB") ? atoi(getenv("IPP_PLAN_EPB")) : 0;  // A/B aid
  if (epb_env >= 1 && epb_env <= PLAN_MAX_ENVS) epb = epb_env;
#endif
  plan_kernel<<<(cfg.n_envs + epb - 1) / epb, PLAN_WARPS * 32, smem, s>>>(cfg, st, io, t, do_comm, do_move,
                                                                         stage_gt | dbg, step_meta, epb, gt_params);
  return cudaGetLastError();
}

cudaError_t launch_step_dense(const ipp_config& cfg, const ipp_state& st, const float4* lut,
                              const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                              double* partials, bool do_own, cudaStream_t s) {
  const int32_t n_chunks = cfg.n_seg;  // one block per (env, flag segment)
  const dim3 grid((unsigned)n_chunks * (unsigned)cfg.n_envs);
  if (do_own) {
    IPP_DISPATCH_A(cfg.n_agents, (step_direct_kernel<kA, true><<<grid, DIRECT_THREADS, 0, s>>>(
                                     cfg, st, lut, step_meta, t, reward_rel, reward_abs, partials, n_chunks)));
  } else {
    IPP_DISPATCH_A(cfg.n_agents, (step_direct_kernel<kA, false><<<grid, DIRECT_THREADS, 0, s>>>(
                                     cfg, st, lut, step_meta, t, reward_rel, reward_abs, partials, n_chunks)));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_chunks > 1) e = launch_reward_finalize(cfg, partials, n_chunks, reward_rel, reward_abs, s);
  return e;
}

cudaError_t launch_reward_finalize(const ipp_config& cfg, const double* partials, int32_t n_chunks, float* reward_rel,
                                   float* reward_abs, cudaStream_t s) {
  const int threads = 128;
  reward_finalize_kernel<<<(cfg.n_envs + threads - 1) / threads, threads, 0, s>>>(
      partials, cfg.n_envs, n_chunks, cfg.gx * cfg.gy, reward_rel, reward_abs);
  return cudaGetLastError();
}

cudaError_t launch_own_update(const ipp_config& cfg, const ipp_state& st, const float4* lut, const int32_t* pos_out,
                              int32_t t, cudaStream_t s) {
  const int32_t n_quads = (cfg.gx * cfg.gy + 3) >> 2;
  int32_t bpe = (n_quads + 1023) / 1024;
  if (bpe < 1) bpe = 1;
  IPP_DISPATCH_A(cfg.n_agents,
                 (own_update_kernel<kA><<<(unsigned)cfg.n_envs * (unsigned)bpe, 256, 0, s>>>(cfg, st, lut, pos_out, t, bpe)));
  return cudaGetLastError();
}

cudaError_t launch_reset(const ipp_config& cfg, const ipp_state& st, const float4* lut, const LaunchPlan& plan,
                         int32_t* pos_out, int32_t* gt_params, cudaStream_t s) {
  const int threads = RESET_PREP_THREADS;
  const int n = cfg.n_envs * (cfg.n_agents + 1);
  reset_prep_kernel<<<(n + threads - 1) / threads, threads, 0, s>>>(cfg, st.episodes, pos_out, gt_params,
                                                                    st.map_flags);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const dim3 grid((unsigned)plan.n_chunks * (unsigned)cfg.n_envs);
  IPP_DISPATCH_A(cfg.n_agents, (reset_fill_kernel<kA><<<grid, STEP_THREADS, 0, s>>>(
                                   cfg, st, lut, pos_out, gt_params, plan.n_chunks, plan.quads_per_chunk)));
  return cudaGetLastError();
}

}  // namespace ipp

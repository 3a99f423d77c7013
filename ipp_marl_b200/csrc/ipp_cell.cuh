// Per-quad (4 consecutive cells) map arithmetic shared by the direct-load and the TMA-staged step
// kernels.  Reference: mapping/mappings.py:80-124 (fuse), :32-78 (own update),
// utils/reward.py:68-82 + utils/state.py:53-76,118-121 (reward terms).
// Specification: oracle/kernel_model.py::_apply / _reward.
#pragma once
#include "ipp_device.cuh"

namespace ipp {

// Everything about one env that the per-cell code needs; lives in shared memory.
template <int A>
struct EnvMeta {
  Meas prev[A];       // communicated measurements (taken at pos_in, index t)
  Meas next[A];       // measurements after the move (taken at pos_out, index t+1)
  uint32_t comm[A];   // bit j: agent i fuses agent j's measurement (own bit cleared)
};

template <int A>
__device__ __forceinline__ void load_env_meta(const ipp_config& cfg, EnvMeta<A>* m, int lane_or_tid, int32_t b,
                                              uint32_t ep, const int32_t* pos_in, const int32_t* pos_out,
                                              const uint8_t* comm, int32_t t, bool do_own) {
  if (lane_or_tid < A) {
    const int a = lane_or_tid;
    m->prev[a] = make_meas(cfg, pos_in + ((int64_t)b * A + a) * 3, ep, a, (uint32_t)t);
    m->comm[a] = (uint32_t)comm[(int64_t)b * A + a] & ~(1u << a);  // own measurement already used
  } else if (lane_or_tid < 2 * A && do_own) {
    const int a = lane_or_tid - A;
    m->next[a] = make_meas(cfg, pos_out + ((int64_t)b * A + a) * 3, ep, a, (uint32_t)t + 1u);
  }
}

// Coordinates + multipliers of the communicated measurements at the 4 cells of one quad.
template <int A>
struct QuadCtx {
  float kprev[A][4];
  uint32_t in_prev;  // bit (j*4+c): cell c lies inside agent j's communicated footprint
  uint32_t valid;    // bit c: cell c0+c < n_cells
  int32_t xs[4], ys[4];
  uint32_t g4;       // 4 ground-truth bytes
  int32_t c0;
};

template <int A>
__device__ __forceinline__ void make_quad_ctx(const ipp_config& cfg, const EnvMeta<A>& meta, int32_t c0, uint32_t g4,
                                              int32_t n_cells, QuadCtx<A>& q) {
  q.c0 = c0;
  q.g4 = g4;
  q.valid = 0;
  {
    int32_t x = c0 / cfg.gy, y = c0 - x * cfg.gy;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      q.xs[c] = x;
      q.ys[c] = y;
      if (c0 + c < n_cells) q.valid |= 1u << c;
      if (++y == cfg.gy) { y = 0; ++x; }
    }
  }
  q.in_prev = 0;
#pragma unroll
  for (int j = 0; j < A; ++j) {
    const Meas m = meta.prev[j];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float k = cfg.k_out;
      if (((q.valid >> c) & 1u) && in_rect(m, q.xs[c], q.ys[c])) {
        k = meas_k(m, (uint32_t)(c0 + c), (g4 >> (8 * c)) & 0xFFu);
        q.in_prev |= 1u << (j * 4 + c);
      }
      q.kprev[j][c] = k;
    }
  }
}

// Global map: fuse every agent's communicated measurement (coma_wrapper.py:93-95) and accumulate
// s1 = sum w*(H(last)-H(next)), s2 = sum w*H(last).
template <int A>
__device__ __forceinline__ void update_global_quad(const ipp_config& cfg, const QuadCtx<A>& q, float (&pv)[4],
                                                   double& s1, double& s2) {
  const bool kout_one = (cfg.k_out == 1.0f);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (!((q.valid >> c) & 1u)) continue;
    const float p = pv[c];
    const float pc = clamp_p(cfg, p);
    const bool touched = ((q.in_prev >> c) & 0x11111111u) != 0u || !kout_one;
    float pn = pc;
    if (touched) {
      float o = to_odds(pc);
#pragma unroll
      for (int j = 0; j < A; ++j) o = odds_pass(o, q.kprev[j][c], cfg.o_min, cfg.o_max);
      pn = from_odds(o);
    }
    const float hl = shannon(cfg, p);
    const float hn = touched ? shannon(cfg, pn) : hl;
    const float w = weight_of(pn);
    s1 += (double)(w * (hl - hn));
    s2 += (double)(w * hl);
    pv[c] = pn;
  }
}

// Local map of agent i: fuse the received peers' measurements (agent/agent.py:62-71), then the own
// measurement at the new position (agent/agent.py:91-94) when DO_OWN.
template <int A, bool DO_OWN>
__device__ __forceinline__ void update_local_quad(const ipp_config& cfg, const EnvMeta<A>& meta,
                                                  const QuadCtx<A>& q, int i, float (&pv)[4]) {
  const bool kout_one = (cfg.k_out == 1.0f);
  const uint32_t en = meta.comm[i];
  const bool any_fuse = en != 0u;
  uint32_t en4 = 0;  // enabled-peer bits replicated over the 4 cells
#pragma unroll
  for (int j = 0; j < A; ++j)
    if ((en >> j) & 1u) en4 |= 0xFu << (j * 4);
  Meas mn;
  if (DO_OWN) mn = meta.next[i];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (!((q.valid >> c) & 1u)) continue;
    const float p = pv[c];
    bool own_in = false;
    if (DO_OWN) own_in = in_rect(mn, q.xs[c], q.ys[c]);
    const bool touched = (((q.in_prev & en4) >> c) & 0x11111111u) != 0u || (any_fuse && !kout_one) || own_in;
    const bool clamped = any_fuse || own_in;
    const float pc = clamp_p(cfg, p);
    float out = clamped ? pc : p;
    if (touched) {
      float o = to_odds(pc);
#pragma unroll
      for (int j = 0; j < A; ++j)
        if ((en >> j) & 1u) o = odds_pass(o, q.kprev[j][c], cfg.o_min, cfg.o_max);
      if (DO_OWN && own_in)
        o = odds_pass(o, meas_k(mn, (uint32_t)(q.c0 + c), (q.g4 >> (8 * c)) & 0xFFu), cfg.o_min, cfg.o_max);
      out = from_odds(o);
    }
    pv[c] = out;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, off);
  return v;
}

__device__ __forceinline__ void write_rewards(float* reward_rel, float* reward_abs, int32_t b, double t1, double t2,
                                              int32_t n_cells) {
  if (reward_rel != nullptr) reward_rel[b] = (float)(22.0 * (t1 / t2) - 0.5);             // utils/reward.py:39-41
  if (reward_abs != nullptr) reward_abs[b] = (float)(10.0 * (t1 / (double)n_cells) - 0.17);  // utils/reward.py:38
}

}  // namespace ipp

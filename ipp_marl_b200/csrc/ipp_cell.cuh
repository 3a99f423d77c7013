// Per-quad (4 consecutive cells) belief arithmetic shared by the direct-load and the TMA-staged map
// kernels.  Reference: mapping/mappings.py:80-124 (fuse), :32-78 (own update),
// utils/reward.py:68-82 + utils/state.py:53-76,118-121 (reward terms).
// Specification: oracle/kernel_model.py::_apply / _reward  (belief maps bit-exact).
//
// The belief maps live in HBM as float32 ODDS o = p/(1-p) (include/ipp_b200.h), so one Bayes pass is
// o = min(max(o, o_min), o_max) * k: no division, no transcendental on the belief path.  What keeps the
// map kernels off the issue limit (profiles/):
//   * everything geometric / random about a measurement (footprint test, noise hash, ground truth) is
//     done once per step by plan_kernel, rect-sparse, and handed over as one CODE BYTE per (quad, agent):
//     low nibble = cell inside the footprint, high nibble = cell seen as 1.  The map kernels turn a
//     code byte into the 4 odds multipliers with ONE 16-byte table load (lut[altitude][byte]);
//   * the 4 cells of a quad are processed branch-free, the multiplies two cells per instruction with the
//     sm_100 packed-float32 FMUL2 form (IEEE per lane => still bit-exact);
//   * a fuse pass whose footprint touches no cell of the whole warp degenerates to a clamp, clamps
//     are idempotent, so such passes are skipped warp-uniformly and one clamp is applied instead;
//   * only the reward needs more: H(p) of the global map's cells from their odds (1 MUFU.RCP + 2 MUFU.LG2).
#pragma once
#include "ipp_device.cuh"

namespace ipp {

// Per-env facts the per-quad code needs (shared memory).
template <int A>
struct EnvMeta {
  uint32_t comm[A];      // bit j: agent i fuses agent j's measurement (own bit cleared)
  uint32_t comm4[A];     // the same, one nibble (0xF) per enabled peer: mask for QuadCtx::in_prev
  uint32_t lut_prev[A];  // float4 index of the LUT row (altitude) of agent j's communicated measurement
  uint32_t lut_next[A];  // same for the measurement after the move
};

__device__ __forceinline__ uint32_t lut_row(const ipp_config& c, const int32_t* pos) {
  const int32_t iz = clampi(pos[2] / c.spacing - c.min_altitude / c.spacing, 0, c.n_alt - 1);
  return (uint32_t)iz * 256u;
}

template <int A>
__device__ __forceinline__ void load_env_meta(const ipp_config& cfg, EnvMeta<A>* m, int lane_or_tid, int32_t b,
                                              const int32_t* pos_in, const int32_t* pos_out, const uint8_t* comm,
                                              bool do_own) {
  if (lane_or_tid < A) {
    const int a = lane_or_tid;
    m->lut_prev[a] = lut_row(cfg, pos_in + ((int64_t)b * A + a) * 3);
    const uint32_t en = (uint32_t)comm[(int64_t)b * A + a] & ~(1u << a);  // own measurement already used
    uint32_t en4 = 0;
    for (int j = 0; j < A; ++j)
      if ((en >> j) & 1u) en4 |= 0xFu << (4 * j);
    m->comm[a] = en;
    m->comm4[a] = en4;
  } else if (lane_or_tid < 2 * A) {
    const int a = lane_or_tid - A;
    m->lut_next[a] = do_own ? lut_row(cfg, pos_out + ((int64_t)b * A + a) * 3) : 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// packed float32 helpers (cells (0,1) in .lo, (2,3) in .hi)
// ------------------------------------------------------------------------------------------------
struct F4 {
  float2 lo, hi;
};

__device__ __forceinline__ F4 f4_from(const float4 v) { return F4{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ float4 f4_to(const F4 v) { return make_float4(v.lo.x, v.lo.y, v.hi.x, v.hi.y); }
__device__ __forceinline__ float f4_get(const F4& v, int c) {
  return c == 0 ? v.lo.x : (c == 1 ? v.lo.y : (c == 2 ? v.hi.x : v.hi.y));
}
__device__ __forceinline__ F4 f4_splat(float s) { return F4{make_float2(s, s), make_float2(s, s)}; }
__device__ __forceinline__ F4 f4_fma(const F4 a, const F4 b, const F4 c) {
  return F4{__ffma2_rn(a.lo, b.lo, c.lo), __ffma2_rn(a.hi, b.hi, c.hi)};
}
__device__ __forceinline__ F4 f4_mul(const F4 a, const F4 b) {
  return F4{__fmul2_rn(a.lo, b.lo), __fmul2_rn(a.hi, b.hi)};
}
__device__ __forceinline__ F4 f4_clamp(const F4 v, float lo, float hi) {
  return F4{make_float2(fminf(fmaxf(v.lo.x, lo), hi), fminf(fmaxf(v.lo.y, lo), hi)),
            make_float2(fminf(fmaxf(v.hi.x, lo), hi), fminf(fmaxf(v.hi.y, lo), hi))};
}
__device__ __forceinline__ F4 f4_select(uint32_t mask4, const F4 a, const F4 b) {  // bit c set -> a else b
  return F4{make_float2((mask4 & 1u) ? a.lo.x : b.lo.x, (mask4 & 2u) ? a.lo.y : b.lo.y),
            make_float2((mask4 & 4u) ? a.hi.x : b.hi.x, (mask4 & 8u) ? a.hi.y : b.hi.y)};
}

// some cell lies outside [lo, hi]
__device__ __forceinline__ bool f4_out_of_range(const F4 v, float lo, float hi) {
  const float mx = fmaxf(fmaxf(v.lo.x, v.lo.y), fmaxf(v.hi.x, v.hi.y));
  const float mn = fminf(fminf(v.lo.x, v.lo.y), fminf(v.hi.x, v.hi.y));
  return mx > hi || mn < lo;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ------------------------------------------------------------------------------------------------
// measurement codes: one byte per (quad, agent), AP = 4 (A <= 4) or 8 bytes per quad
// ------------------------------------------------------------------------------------------------
template <int A>
struct CodeWord {
  static constexpr int WORDS = (A + 3) / 4;
  uint32_t w[WORDS];
  __device__ __forceinline__ uint32_t byte(int j) const { return (w[j >> 2] >> (8 * (j & 3))) & 0xFFu; }
};

template <int A>
__device__ __forceinline__ CodeWord<A> load_code(const void* base, int32_t quad) {
  CodeWord<A> c;
  if (CodeWord<A>::WORDS == 1) {
    c.w[0] = reinterpret_cast<const uint32_t*>(base)[quad];
  } else {
    const uint2 v = reinterpret_cast<const uint2*>(base)[quad];
    c.w[0] = v.x;
    c.w[CodeWord<A>::WORDS - 1] = v.y;
  }
  return c;
}

template <int A>
struct QuadCtx {
  F4 kprev[A];       // multipliers of the communicated measurements (k_out outside a footprint)
  uint32_t in_prev;  // bits 4j..4j+3: cells inside agent j's communicated footprint
  uint32_t wcov;     // bit j: some active lane of this warp has a cell inside footprint j
};

// Must be called with the warp's loop-active lanes converged (it votes over __activemask()).
// `lut` is the [n_alt][256] float4 table (shared or global memory).
template <int A>
__device__ __forceinline__ void make_quad_ctx(const ipp_config& cfg, const EnvMeta<A>& meta, const CodeWord<A>& prev,
                                              const float4* lut, QuadCtx<A>& q) {
  q.in_prev = 0;
  q.wcov = 0;
  const uint32_t active = __activemask();
  const bool kout_one = (cfg.k_out == 1.0f);
#pragma unroll
  for (int j = 0; j < A; ++j) {
    const uint32_t byte = prev.byte(j);
    const uint32_t in = byte & 0xFu;
    q.in_prev |= in << (4 * j);
    q.kprev[j] = f4_splat(cfg.k_out);
    if (__any_sync(active, in != 0u) || !kout_one) {  // warp-uniform
      q.wcov |= 1u << j;
      q.kprev[j] = f4_from(lut[meta.lut_prev[j] + byte]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One belief map (float32 odds) of the quad through its chain of passes.
//   en      : bit j = fuse pass j enabled for this map (warp-uniform)
//   own     : 4-bit mask of cells inside the own new footprint, k_own their multipliers
// Semantics per cell (oracle/kernel_model.py::_apply): every enabled fuse pass clamps the odds and
// multiplies by k_j (k_out outside footprint j); then, inside the own footprint only, clamp and
// multiply by k_own.  Untouched cells keep o (or clamp(o) if some fuse pass ran) bit for bit.
// Padding cells beyond gx*gy are never inside a footprint, hold the prior and stay unchanged.
// oc_out = the clamped input odds (the reward's "last" map), touched = cells some pass multiplied.
// ------------------------------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ F4 update_map_quad(const ipp_config& cfg, const QuadCtx<A>& q, const F4 o_in,
                                              const uint32_t en, const uint32_t en4, const uint32_t own,
                                              const F4 k_own, F4& oc_out, uint32_t& touched) {
  const bool kout_one = (cfg.k_out == 1.0f);
  const bool any_fuse = en != 0u;
  uint32_t x = q.in_prev & en4;  // footprint nibbles of the enabled passes, OR-folded into one nibble
  if (A > 4) x |= x >> 16;
  if (A > 2) x |= x >> 8;
  if (A > 1) x |= x >> 4;
  uint32_t t = (x & 0xFu) | own;
  if (any_fuse && !kout_one) t = 0xFu;
  touched = t;
  F4 o = o_in;
  bool clean = false;  // o is known to lie inside [o_min, o_max] (just clamped)
  if (any_fuse) {      // the first enabled fuse pass clamps every cell of the map
    o = f4_clamp(o, cfg.o_min, cfg.o_max);
    clean = true;
  }
  oc_out = o;
#pragma unroll
  for (int j = 0; j < A; ++j) {
    if (!((en >> j) & 1u)) continue;  // warp-uniform
    if ((q.wcov >> j) & 1u) {         // warp-uniform: somebody's cell is inside footprint j
      if (!clean) o = f4_clamp(o, cfg.o_min, cfg.o_max);
      o = f4_mul(o, q.kprev[j]);
      clean = false;
    }
    // else: the pass is a pure clamp for the whole warp; clamps are idempotent, so it is absorbed by
    // the clamp of the next executed pass or by the one below
  }
  // A fuse pass skipped AFTER the last executed one still owes its clamp to every cell (warp-uniform).
  {
    const uint32_t exec = en & q.wcov;
    const bool pending = exec != 0u && (en >> (32 - __clz(exec))) != 0u;  // enabled pass above the last executed
    if (pending) {
      o = f4_clamp(o, cfg.o_min, cfg.o_max);
      clean = true;
    }
  }
  if (own != 0u) {  // own update: only the cells inside the own footprint are clamped and multiplied
    const F4 oc = clean ? o : f4_clamp(o, cfg.o_min, cfg.o_max);
    o = f4_select(own, f4_mul(oc, k_own), o);
  }
  return o;
}

// float32 reward terms; H in bits (utils/state.py:118-121) of a cell given its CLAMPED odds:
// q = 1/(1+o), p = o*q, H = -(p lg p + q lg q)
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float entropy_bits_odds(float oc) {
  const float qc = rcp_approx(1.0f + oc);
  const float pc = oc * qc;
  return -(pc * lg2_approx(pc) + qc * lg2_approx(qc));
}

// utils/state.py:67-73 thresholds p > 0.501 / p < 0.499 in odds space (oracle/kernel_model.py W_HI / W_LO)
#define IPP_W_HI 1.0040080547332764f
#define IPP_W_LO 0.9960079789161682f

// Global map: fuse every agent's communicated measurement (coma_wrapper.py:93-95) and accumulate
// s1 = sum w*(H(last)-H(next)), s2 = sum w*H(last)  (utils/reward.py:68-82).  `valid`: bit c = cell exists.
template <int A>
__device__ __forceinline__ float4 update_global_quad(const ipp_config& cfg, const QuadCtx<A>& q, const float4 o4,
                                                     const uint32_t valid, double& s1, double& s2) {
  F4 oc;
  uint32_t touched;
  const F4 on = update_map_quad<A>(cfg, q, f4_from(o4), (1u << A) - 1u, 0xFFFFFFFFu, 0u, f4_splat(1.0f), oc,
                                   touched);
  float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (!((valid >> c) & 1u)) continue;
    const float next = f4_get(on, c);
    const float hl = entropy_bits_odds(f4_get(oc, c));
    float hn = hl;
    if (touched != 0u)  // quad-level branch; untouched cells of a touched quad reuse hl
      hn = ((touched >> c) & 1u) ? entropy_bits_odds(fminf(fmaxf(next, cfg.o_min), cfg.o_max)) : hl;
    const float w = next > IPP_W_HI ? 1.0f : (next < IPP_W_LO ? 0.0f : 0.5f);
    a1 += w * (hl - hn);
    a2 += w * hl;
  }
  s1 += (double)a1;
  s2 += (double)a2;
  return f4_to(on);
}

// Local map of agent i: fuse the received peers' measurements (agent/agent.py:62-71), then the own
// measurement at the new position (agent/agent.py:91-94) when DO_OWN (code byte `own_byte`).
template <int A, bool DO_OWN>
__device__ __forceinline__ float4 update_local_quad(const ipp_config& cfg, const EnvMeta<A>& meta,
                                                    const QuadCtx<A>& q, int i, const uint32_t own_byte,
                                                    const float4* lut, const float4 o4) {
  uint32_t own = 0;
  F4 k_own = f4_splat(1.0f);
  if (DO_OWN) {
    own = own_byte & 0xFu;
    k_own = f4_from(lut[meta.lut_next[i] + own_byte]);
  }
  F4 oc;
  uint32_t touched;
  return f4_to(update_map_quad<A>(cfg, q, f4_from(o4), meta.comm[i], meta.comm4[i], own, k_own, oc, touched));
}

__device__ __forceinline__ uint32_t valid_mask4(int32_t c0, int32_t n_cells) {
  const int32_t left = n_cells - c0;
  return left >= 4 ? 0xFu : ((1u << max(left, 0)) - 1u);
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

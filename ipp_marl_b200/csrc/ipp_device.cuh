// Device-side building blocks shared by all kernels: counter-based hash, footprint geometry,
// odds-space Bayes update.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ipp_b200.h"

namespace ipp {

// ------------------------------------------------------------------------------------------------
// Counter-based hash (specification: oracle/noise.py).  Replaces the reference's global-RNG draws
// (mapping/simulations.py:56-58, actor/network.py:94, agent/communication_log.py:46).
// ------------------------------------------------------------------------------------------------
enum : uint32_t { PURPOSE_NOISE = 0, PURPOSE_ACTION = 1, PURPOSE_COMM = 2 };

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x21F0AAADu;
  x ^= x >> 15;
  x *= 0x735A2D97u;
  x ^= x >> 15;
  return x;
}

__host__ __device__ __forceinline__ uint32_t stream_key(uint32_t seed, uint32_t episode, uint32_t agent,
                                                        uint32_t index, uint32_t purpose) {
  uint32_t k = mix32(seed + 0x9E3779B9u);
  k = mix32(k ^ episode);
  k = mix32(k + ((purpose << 24) | (agent << 16) | index));
  return k;
}

__host__ __device__ __forceinline__ uint32_t cell_hash(uint32_t key, uint32_t cell) {
  return mix32(key ^ (cell * 0x9E3779B1u));
}

// ------------------------------------------------------------------------------------------------
// Footprint of a measurement (sensors/cameras.py:46-79 through the host tables).
// ------------------------------------------------------------------------------------------------
struct Meas {
  int32_t xl, xr, yu, yd;  // clipped, half-open like the reference's slices (mappings.py:46-49)
  uint32_t key;            // noise stream of this measurement
  uint32_t thresh;         // flip threshold of its altitude
  float k_hi, k_lo;        // odds multipliers
};

__device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return min(max(v, lo), hi); }

// Measurement codes: [n_envs][2][code_stride] — the two ping-pong halves of one env are adjacent (the map kernel
// fetches both with one bulk copy); half (t & 1) holds the measurements communicated at step t.
__host__ __device__ __forceinline__ int64_t code_row_offset(const ipp_config& c, int32_t half, int64_t b) {
  return (b * 2 + (half & 1)) * (int64_t)c.code_stride;
}

__device__ __forceinline__ Meas make_meas(const ipp_config& c, const int32_t* pos, uint32_t episode, uint32_t agent,
                                          uint32_t index) {
  Meas m;
  // clamped so that a corrupt position can never index outside the tables
  const int32_t ix = clampi(pos[0] / c.spacing, 0, c.px - 1), iy = clampi(pos[1] / c.spacing, 0, c.py - 1);
  const int32_t iz = clampi(pos[2] / c.spacing - c.min_altitude / c.spacing, 0, c.n_alt - 1);
  const int32_t cx = c.cell_x[ix], cy = c.cell_y[iy];
  const int32_t rx = c.radius_x[iz], ry = c.radius_y[iz];
  m.xl = clampi(cx - rx, 0, c.gx - 1);
  m.xr = clampi(cx + rx, 0, c.gx - 1);
  m.yu = clampi(cy - ry, 0, c.gy - 1);
  m.yd = clampi(cy + ry, 0, c.gy - 1);
  m.key = stream_key(c.seed, episode, agent, index, PURPOSE_NOISE);
  m.thresh = c.flip_thresh[iz];
  m.k_hi = c.k_hi[iz];
  m.k_lo = c.k_lo[iz];
  return m;
}

// 4-bit footprint mask of a quad.  Cells c0..c0+3 lie in row x0 from column y0 (n0 = cells before
// the row wraps, 1..4) and, when n0 < 4, continue in row x0+1 from column 0 (requires gy >= 4).
__device__ __forceinline__ uint32_t rect_mask4(const Meas& m, int32_t x0, int32_t y0, int32_t n0) {
  uint32_t mask = 0;
  {
    const int32_t lo = max(m.yu - y0, 0), hi = min(m.yd - y0, n0);
    if ((uint32_t)(x0 - m.xl) < (uint32_t)(m.xr - m.xl) && hi > lo) mask = (1u << hi) - (1u << lo);
  }
  if (n0 < 4) {
    const int32_t lo = m.yu, hi = min(m.yd, 4 - n0);
    if ((uint32_t)(x0 + 1 - m.xl) < (uint32_t)(m.xr - m.xl) && hi > lo) mask |= ((1u << hi) - (1u << lo)) << n0;
  }
  return mask;
}

// 4-bit "seen as 1" mask of the quad starting at cell c0 (c0 % 4 == 0) for measurement stream `key`:
// one hash per quad + one xor/multiply per cell (oracle/noise.py::noise_word), then the flip test
// against the altitude's threshold and the ground truth (mapping/simulations.py:53-65).
__device__ __forceinline__ uint32_t seen_mask4(uint32_t key, uint32_t thresh, int32_t c0, uint32_t g4) {
  const uint32_t h = cell_hash(key, (uint32_t)c0 >> 2);
  const uint32_t w0 = h * 0x9E3779B1u, w1 = (h ^ 0x85EBCA6Bu) * 0x85EBCA77u;
  const uint32_t w2 = (h ^ 0xC2B2AE35u) * 0xC2B2AE3Du, w3 = (h ^ 0x27D4EB2Fu) * 0x27D4EB2Fu;
  uint32_t seen = 0;
  seen |= ((((g4)&0xFFu) != 0u) != (w0 < thresh)) ? 1u : 0u;
  seen |= ((((g4 >> 8) & 0xFFu) != 0u) != (w1 < thresh)) ? 2u : 0u;
  seen |= ((((g4 >> 16) & 0xFFu) != 0u) != (w2 < thresh)) ? 4u : 0u;
  seen |= ((((g4 >> 24) & 0xFFu) != 0u) != (w3 < thresh)) ? 8u : 0u;
  return seen;
}

// noise word of one cell (used where cells are visited one by one: facade measure kernel)
__host__ __device__ __forceinline__ uint32_t noise_word(uint32_t key, uint32_t cell) {
  const uint32_t h = cell_hash(key, cell >> 2);
  switch (cell & 3u) {
    case 0: return h * 0x9E3779B1u;
    case 1: return (h ^ 0x85EBCA6Bu) * 0x85EBCA77u;
    case 2: return (h ^ 0xC2B2AE35u) * 0xC2B2AE3Du;
    default: return (h ^ 0x27D4EB2Fu) * 0x27D4EB2Fu;
  }
}

// Measurement code byte of quad (cells c0..c0+3) for measurement m: low nibble = cell inside the
// footprint, high nibble = cell seen as 1.
__device__ __forceinline__ uint32_t meas_code_byte(const ipp_config& c, const Meas& m, int32_t c0, uint32_t g4) {
  const int32_t x0 = c0 / c.gy, y0 = c0 - x0 * c.gy;
  const int32_t left = c.gx * c.gy - c0;
  const uint32_t valid = left >= 4 ? 0xFu : ((1u << max(left, 0)) - 1u);
  const uint32_t in = rect_mask4(m, x0, y0, min(4, c.gy - y0)) & valid;
  if (in == 0u) return 0u;
  return in | ((seen_mask4(m.key, m.thresh, c0, g4) & in) << 4);
}

// ------------------------------------------------------------------------------------------------
// Bayes update in odds space (specification: oracle/kernel_model.py::_apply).
// Only IEEE float32 +,-,*,/,min,max: the belief path is bit-reproducible on the CPU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_p(const ipp_config& c, float p) { return fminf(fmaxf(p, c.p_min), c.p_max); }

__device__ __forceinline__ float to_odds(float pc) { return __fdiv_rn(pc, __fsub_rn(1.0f, pc)); }

__device__ __forceinline__ float odds_pass(float o, float k, float o_min, float o_max) {
  return __fmul_rn(fminf(fmaxf(o, o_min), o_max), k);
}

__device__ __forceinline__ float from_odds(float o) {
  const float d = __fadd_rn(1.0f, o);
  return (o < 1.0f) ? __fdiv_rn(o, d) : __fsub_rn(1.0f, __fdiv_rn(1.0f, d));
}

// The same within ~2 ulp from one MUFU.RCP instead of an IEEE division (26 instructions with its two-sided branch):
// for consumers whose own tolerance is far above float32 rounding (the area-pooled network inputs; NOT the exported
// beliefs, which are bit-exact against the kernel model).
__device__ __forceinline__ float from_odds_fast(float o) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + o));
  return (o < 1.0f) ? o * r : 1.0f - r;
}

// Weighted entropy terms of the reward (utils/state.py:53-76,118-121; utils/reward.py:68-82).
__device__ __forceinline__ float shannon(const ipp_config& c, float p) {
  const float pc = clamp_p(c, p);
  const float q = 1.0f - pc;
  return -pc * log2f(pc) - q * log2f(q);
}

__device__ __forceinline__ float weight_of(float p_next) {
  const double d = (double)p_next;
  return d > 0.501 ? 1.0f : (d < 0.499 ? 0.0f : 0.5f);
}

}  // namespace ipp

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
//   * a local-map tile whose range flag (ipp_state.map_flags) is clear lies inside [o_min, o_max], so the whole-map
//     clamp of a fuse pass changes nothing outside the footprints: (tile, map) pairs no footprint reaches are
//     neither loaded nor stored (the plan kernel's ItemRec says which have work), and inside a chain a clamp is
//     executed only where the value may actually be out of range (fuse_chain);
//   * only the reward needs more: H(p) of the global map's cells from their odds (1 MUFU.RCP + 2 MUFU.LG2).
#pragma once
#include "ipp_device.cuh"

namespace ipp {

// Per-env facts the per-quad code needs (shared memory; part of the record written by the plan kernel).
template <int A>
struct EnvMeta {
  uint32_t comm[A];      // bit j: agent i fuses agent j's measurement (own bit cleared)
  uint32_t comm4[A];     // the same, one nibble (0xF) per enabled peer: mask for the in-footprint nibbles of a quad (in_prev)
  uint32_t lut_prev[A];  // float4 index of the LUT row (altitude) of agent j's communicated measurement
  uint32_t lut_next[A];  // same for the measurement after the move
};

constexpr int ITEM_TILES = IPP_FLAG_QUADS / 32;  // tiles (32 quads = 128 cells) per work item / flag segment

// Record of one WORK ITEM = (env, segment of IPP_FLAG_QUADS quads), written by the plan kernel once per step and
// read by the map kernels as a plain copy: the env-level facts plus, for this segment, which tiles of which local map
// have work in this step — everything the map kernel's producer needs before it can fetch the item, worked out
// where there are thousands of warps to do it.
template <int A>
struct ItemRec {
  EnvMeta<A> env;
  uint32_t need[A];   // bit t: tile t of local map i has work: flagged out of range while a fuse pass runs, reached by
                      // an enabled peer's communicated footprint, or by the own new footprint (a superset is allowed)
  uint32_t flags[A];  // the segment's range flags (ipp_state.map_flags) as they were before this step
  uint16_t tile[ITEM_TILES];  // per tile: low byte bit j = agent j's communicated footprint reaches the tile; high
                              // byte bit i = a fuse pass runs on local map i and the tile may hold out-of-range odds
                              // (or k_out != 1): every quad of the tile changes
};
// words of one record in global memory (padded to 16 bytes, the bulk-copy granularity)
__host__ __device__ constexpr int rec_words(int n_agents) { return (6 * n_agents + ITEM_TILES / 2 + 3) & ~3; }

__device__ __forceinline__ uint32_t lut_row(const ipp_config& c, const int32_t* pos) {
  const int32_t iz = clampi(pos[2] / c.spacing - c.min_altitude / c.spacing, 0, c.n_alt - 1);
  return (uint32_t)iz * 256u;
}

// Tiles of 128 cells reached by the footprint of a measurement taken at `pos` (same geometry as make_meas): the
// footprint's rows xl..xr-1 hold cells [x*gy+yu, x*gy+yd), so everything it touches lies between its first and its
// last cell.  (Exact when a tile is longer than the gap between two footprint rows, always a superset.)
__device__ __forceinline__ uint32_t tile_range(const ipp_config& c, const int32_t* pos) {
  const int32_t ix = clampi(pos[0] / c.spacing, 0, c.px - 1), iy = clampi(pos[1] / c.spacing, 0, c.py - 1);
  const int32_t iz = clampi(pos[2] / c.spacing - c.min_altitude / c.spacing, 0, c.n_alt - 1);
  const int32_t cx = c.cell_x[ix], cy = c.cell_y[iy], rx = c.radius_x[iz], ry = c.radius_y[iz];
  const int32_t xl = clampi(cx - rx, 0, c.gx - 1), xr = clampi(cx + rx, 0, c.gx - 1);
  const int32_t yu = clampi(cy - ry, 0, c.gy - 1), yd = clampi(cy + ry, 0, c.gy - 1);
  if (xr <= xl || yd <= yu) return 1u;  // first 1 > last 0: empty footprint
  const uint32_t first = (uint32_t)(xl * c.gy + yu), last = (uint32_t)((xr - 1) * c.gy + yd - 1);
  return (first >> 7) | ((last >> 7) << 16);
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

// float32 reward terms; H in bits (utils/state.py:118-121) of a cell given its CLAMPED odds (1 MUFU.RCP + 2 MUFU.LG2)
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float entropy_bits_odds(float oc) {
  // with t = 1 + o, q = 1/t, p = o q:  H = -(p lg p + q lg q) = lg t - p lg o   (lg p = lg o - lg t, p + q = 1)
  const float t = 1.0f + oc;
  const float pc = oc * rcp_approx(t);
  return fmaf(-pc, lg2_approx(oc), lg2_approx(t));
}

// utils/state.py:67-73 thresholds p > 0.501 / p < 0.499 in odds space (oracle/kernel_model.py W_HI / W_LO)
#define IPP_W_HI 1.0040080547332764f
#define IPP_W_LO 0.9960079789161682f

// ------------------------------------------------------------------------------------------------
// One fuse chain on a quad: the passes of the ENABLED agents in id order; each pass clamps its input and multiplies
// (mapping/mappings.py:106-124; by exactly k_out outside footprint j).  `touch` (warp-uniform) marks the passes whose
// footprint reaches one of the warp's quads; with k_out == 1 the others multiply every cell of the warp by exactly 1,
// so only their clamp can matter — and a clamp is the identity on a value that is known to lie inside
// [o_min, o_max].  `oor` (warp-uniform, in/out) = "the value may lie outside the range": a clamp is executed exactly
// when it is set.  Multipliers come from the LUT row of agent j's altitude, indexed by the quad's code byte.
// ------------------------------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ F4 fuse_chain(const ipp_config& cfg, const EnvMeta<A>& meta, const CodeWord<A>& cw,
                                         const float4* lut, F4 o, const uint32_t enabled, const uint32_t touch,
                                         bool& oor) {
  const float lo = cfg.o_min, hi = cfg.o_max;
#pragma unroll
  for (int j = 0; j < A; ++j) {
    if (!((enabled >> j) & 1u)) continue;  // warp-uniform
    if (oor) o = f4_clamp(o, lo, hi);
    oor = false;
    if ((touch >> j) & 1u) {
      o = f4_mul(o, f4_from(lut[meta.lut_prev[j] + cw.byte(j)]));
      oor = true;
    }
  }
  return o;
}

// ------------------------------------------------------------------------------------------------
// GLOBAL map, one quad: all A fuse passes (coma_wrapper.py:93-95) + the reward terms
// s1 = sum w*(H(last)-H(next)), s2 = sum w*H(last) (utils/reward.py:68-82).
// Must be called by all 32 lanes of the warp, converged (lanes without a quad pass valid = 0).
// ------------------------------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ float4 global_quad(const ipp_config& cfg, const EnvMeta<A>& meta, const CodeWord<A>& cw,
                                              const float4* lut, float4 o4, const uint32_t valid,
                                              const uint32_t touch, float& s1, float& s2) {
  const float lo = cfg.o_min, hi = cfg.o_max;
  // Cells that do not count for the reward — beyond gx*gy in the map's last quad, or all four of a lane without a
  // quad — are given odds 0: clamped to o_min, never inside a footprint, their weight is 0 (o_min k_out^A < W_LO)
  // and so are both of their reward terms.  (Rare, so a branch; what is stored for padding cells is never read.)
  if (valid != 0xFu) {
    if (!(valid & 1u)) o4.x = 0.0f;
    if (!(valid & 2u)) o4.y = 0.0f;
    if (!(valid & 4u)) o4.z = 0.0f;
    if (!(valid & 8u)) o4.w = 0.0f;
  }
  const F4 oc = f4_clamp(f4_from(o4), lo, hi);
  bool oor = false;  // oc is clamped
  const F4 o = fuse_chain<A>(cfg, meta, cw, lut, oc, (1u << A) - 1u, touch, oor);
  // H(next) only where some footprint reaches the warp's tile (warp-uniform): an untouched cell has
  // next == clamp(last) bit for bit, so its H(next) == H(last) exactly
  const bool any_touched = touch != 0u || cfg.k_out != 1.0f;
  float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float next = f4_get(o, c);
    const float hl = entropy_bits_odds(f4_get(oc, c));
    float hn = hl;
    if (any_touched) hn = entropy_bits_odds(fminf(fmaxf(next, lo), hi));
    const float w = next > IPP_W_HI ? 1.0f : (next < IPP_W_LO ? 0.0f : 0.5f);
    a1 += w * (hl - hn);
    a2 += w * hl;
  }
  s1 = a1;  // float32 per-quad terms; the callers sum 32 quads in float32, then everything in float64
  s2 = a2;
  return f4_to(o);
}

// ------------------------------------------------------------------------------------------------
// LOCAL map of agent i, one quad (agent/agent.py:62-71,91-94; mapping/mappings.py:80-124,32-61): the enabled
// fuse passes in id order, then, inside the new own footprint only, clamp and multiply.  `tile_dirty`
// (warp-uniform): the stored odds of this tile may lie outside [o_min, o_max] (range flag).  Returns true when a
// result left the range (the reference clamps lazily, at the next update that reads the cell).
// ------------------------------------------------------------------------------------------------
template <int A, bool DO_OWN>
__device__ __forceinline__ bool local_quad(const ipp_config& cfg, const EnvMeta<A>& meta, const uint32_t enabled,
                                           const uint32_t touch, const bool tile_dirty, const int i,
                                           const CodeWord<A>& cw, const uint32_t own_byte, const float4* lut,
                                           float4& v) {
  const float lo = cfg.o_min, hi = cfg.o_max;
  bool oor = tile_dirty;
  F4 o = fuse_chain<A>(cfg, meta, cw, lut, f4_from(v), enabled, touch, oor);
  if (DO_OWN) {
    const uint32_t own = own_byte & 0xFu;
    if (own != 0u) o = f4_select(own, f4_mul(f4_clamp(o, lo, hi), f4_from(lut[meta.lut_next[i] + own_byte])), o);
  }
  v = f4_to(o);
  return f4_out_of_range(o, lo, hi);
}

// some component of a differs from b bit for bit
__device__ __forceinline__ bool quad_changed(const float4 a, const float4 b) {
  return __float_as_uint(a.x) != __float_as_uint(b.x) || __float_as_uint(a.y) != __float_as_uint(b.y) ||
         __float_as_uint(a.z) != __float_as_uint(b.z) || __float_as_uint(a.w) != __float_as_uint(b.w);
}

__device__ __forceinline__ uint32_t valid_mask4(int32_t c0, int32_t n_cells) {
  const int32_t left = n_cells - c0;
  return left >= 4 ? 0xFu : ((1u << max(left, 0)) - 1u);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, off);
  return v;
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

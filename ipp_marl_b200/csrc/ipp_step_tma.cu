// TMA-staged, warp-specialised, persistent variant of the map kernel (sm_100a).
//
// One CTA per SM walks the (env, chunk) work items.  A chunk is 640 quads (2560 cells) = 20 TILES of 32 quads, so
// the 50x50 grid is a single chunk.  Shared memory holds two rings:
//   * 16 MAP slots of 10 KB: one belief map of one item per slot (an item takes A + 1 consecutive slots);
//   * 4 ENV slots: the item's two measurement-code rows, its StageMeta record and its reward partial sums.
// Roles (no block-wide barrier after start-up; three mbarriers per env slot do all the synchronisation):
//   producer warp : (one lane) per item: arms env_full with the item's total byte count, then bulk-loads
//                   (cp.async.bulk, SASS UBLKCP) the env's record from the plan kernel, its range flags, the code
//                   rows and the A + 1 maps, each map as soon as the item that used its slot before has been
//                   consumed.  It issues no ordinary load, so it never waits on memory itself — with the record
//                   read by plain loads this warp's latency was the whole kernel's bottleneck (profiles/);
//   consumer warps: pull (item, tile) tasks from a shared counter (dynamic load balance), wait ONCE for the item
//                   (env_full), decode the tile's code bytes and take its 32 quads through the item's maps in
//                   registers — all loads, clamp / multiply chains and stores of the tile are independent work:
//                   * GLOBAL map: every cell gets all A fuse passes and its reward terms — straight-line, the four
//                     multipliers of a quad from one LUT load per agent;
//                   * LOCAL maps: the enabled fuse passes and the own update, but only where it matters: a map
//                     whose range flag is clear lies inside [o_min, o_max], so the whole-map clamp of the
//                     reference is a no-op outside every footprint and a (tile, map) pair that no footprint
//                     reaches is skipped by one warp vote, without touching shared memory;
//                   results go from registers straight to global memory with coalesced streaming 16-byte stores
//                   (512 contiguous bytes per warp); the slots are only ever READ by the SM, so they are free as
//                   soon as the item's last tile has arrived on env_tiles — shared memory holds data that is
//                   loading or waiting for a warp, never data that is draining to HBM;
//   finisher warp : (one lane) waits env_tiles, finishes the per-env reward from the tiles' partial sums in a
//                   fixed order, writes the local maps' new range flags and hands the env slot back (env_done).
// HBM traffic is one read of every belief map and of the code rows plus the write of the quads that changed.
#include <cstdlib>

#include "ipp_cell.cuh"
#include "ipp_launch.h"
#include "ipp_ptx.cuh"

namespace ipp {

template <int A>
struct alignas(16) StageMeta {
  alignas(16) EnvMeta<A> env;     // bulk-copied: the env's record written by the plan kernel (16 A bytes)
  alignas(16) uint32_t dirty[8];  // bulk-copied: per local map, the tiles flagged "may be out of range" (map_flags)
  int32_t b, chunk, nq, pad;
  uint32_t bad[8];                // per local map, the tiles whose results left [o_min, o_max] in this step
};

static_assert(TMA_QPC == IPP_FLAG_QUADS, "one range flag per (local map, work item)");

constexpr int TMA_D_MAP = 16;     // map slots (power of two)
constexpr int TMA_D_ENV = 4;      // env slots (power of two)
constexpr int TMA_NT = TMA_QPC / 32;  // tiles per item

// Shared-memory layout:
//   [D_MAP][slot_bytes] map slots | [D_ENV][env_bytes] code rows | lut[n_alt*256] float4 | StageMeta[D_ENV] |
//   mbarriers: env_full[D_ENV] env_tiles[D_ENV] env_done[D_ENV] | reward partials [D_ENV][2][NT] double | tile counter
template <int A, bool DO_OWN>
__global__ void __launch_bounds__(tma_threads(A), 1)
    step_tma_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut_g,
                    const uint32_t* __restrict__ step_meta, const int32_t t, float* __restrict__ reward_rel,
                    float* __restrict__ reward_abs, double* __restrict__ partials, const int32_t n_chunks,
                    const int32_t n_items, const int32_t slot_bytes, const int32_t env_bytes, const int32_t dbg_arg) {
  extern __shared__ __align__(128) unsigned char smem[];
#ifdef IPP_TMA_TIMING_KNOBS  // scripts/dbg_bench.py: IPP_TMA_DEBUG bit 0 = load pipeline only, 1 = no global map, 2 = no
  const int32_t dbg = dbg_arg;  // local maps (results are then wrong); compiled out of the product build
#else
  constexpr int32_t dbg = 0;
  (void)dbg_arg;
#endif
  constexpr int NT = TMA_NT;
  constexpr int AP = A <= 4 ? 4 : 8;
  constexpr int QPC = TMA_QPC;
  constexpr int CONSUMER_THREADS = tma_consumer_warps(A) * 32;
  constexpr int TMA_THREADS = tma_threads(A);
  unsigned char* map_slots = smem;
  unsigned char* env_slots = map_slots + (size_t)TMA_D_MAP * slot_bytes;
  float4* lut = reinterpret_cast<float4*>(env_slots + (size_t)TMA_D_ENV * env_bytes);
  StageMeta<A>* meta = reinterpret_cast<StageMeta<A>*>(lut + cfg.n_alt * 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + TMA_D_ENV);
  double* red = reinterpret_cast<double*>(bars + 3 * TMA_D_ENV);  // [D_ENV][2][NT]
  uint32_t* tile_counter = reinterpret_cast<uint32_t*>(red + TMA_D_ENV * 2 * NT);
  // 32-bit shared addresses of the barrier arrays (8 bytes per barrier)
  const uint32_t env_full = ptx::smem_u32(bars);            // producer's arrive.expect_tx + the bulk copies' bytes
  const uint32_t env_tiles = env_full + 8u * TMA_D_ENV;     // one arrival per finished tile task
  const uint32_t env_done = env_tiles + 8u * TMA_D_ENV;     // finisher

  const int32_t tid = threadIdx.x;
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  const uint32_t code_row = (uint32_t)QPC * AP;

  if (tid == 0) {
    for (int s = 0; s < TMA_D_ENV; ++s) {
      ptx::mbar_init(env_full + 8u * s, 1);
      ptx::mbar_init(env_tiles + 8u * s, NT);
      ptx::mbar_init(env_done + 8u * s, 1);
    }
    *tile_counter = 0u;
    ptx::fence_mbar_init();
  }
  for (int32_t i = tid; i < cfg.n_alt * 256; i += TMA_THREADS) lut[i] = lut_g[i];
  __syncthreads();

  if (tid >= CONSUMER_THREADS + 32) {
    // ================================================================== finisher warp (one lane)
    // An env slot is recycled only after this lane's own arrival on env_done, so the phase parity of env_tiles
    // can never run two phases ahead of the wait.
    if (tid != CONSUMER_THREADS + 32) return;
    uint32_t k = 0;
    for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const uint32_t es = k & (TMA_D_ENV - 1), pe = (k / TMA_D_ENV) & 1u;
      ptx::mbar_wait(env_tiles + 8u * es, pe);  // every tile task of the item has finished all A + 1 maps
      const StageMeta<A>& sm = meta[es];
      const int32_t b = sm.b, chunk = sm.chunk;
      {
        const double* r = red + (size_t)es * 2 * NT;
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int w = 0; w < NT; ++w) {
          t1 += r[w];
          t2 += r[NT + w];
        }
        if (n_chunks == 1) {
          write_rewards(reward_rel, reward_abs, b, t1, t2, n_cells);
        } else {
          partials[((int64_t)b * n_chunks + chunk) * 2 + 0] = t1;
          partials[((int64_t)b * n_chunks + chunk) * 2 + 1] = t2;
        }
      }
#pragma unroll 1
      for (int i = 0; i < A; ++i) {
        // new tile flags: a fuse pass clamped every tile it had to (flagged tiles are processed densely), so only
        // this step's results can be out of range; without a fuse pass the old bits stay
        st.map_flags[((int64_t)b * cfg.n_seg + chunk) * 8 + i] = sm.bad[i] | (sm.env.comm[i] == 0u ? sm.dirty[i] : 0u);
      }
      ptx::mbar_arrive(env_done + 8u * es);
    }
    return;
  }

  if (tid >= CONSUMER_THREADS) {
    // ================================================================== producer warp (one lane)
    if (tid != CONSUMER_THREADS) return;
    uint32_t k = 0;
    uint32_t released = 0;  // items [0, released) are known to be fully consumed (their map slots may be reused)
    for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      const uint32_t es = k & (TMA_D_ENV - 1), pe = (k / TMA_D_ENV) & 1u;
      ptx::mbar_wait(env_done + 8u * es, pe ^ 1u);  // the env slot's previous item (k - D_ENV) is finished
      if (k >= TMA_D_ENV && released < k - TMA_D_ENV + 1) released = k - TMA_D_ENV + 1;
      const int32_t b = item / n_chunks;
      const int32_t chunk = item - b * n_chunks;
      const int32_t nq = min(QPC, n_quads - chunk * QPC);
      meta[es].b = b;
      meta[es].chunk = chunk;
      meta[es].nq = nq;
#pragma unroll
      for (int i = 0; i < A; ++i) meta[es].bad[i] = 0u;
      const uint32_t efull = env_full + 8u * es;
      const uint32_t code_bytes = ((uint32_t)nq * AP + 15u) & ~15u;
      const uint32_t map_bytes = (uint32_t)nq * 16u;
      const uint32_t edst = ptx::smem_u32(env_slots + (size_t)es * env_bytes);
      const int64_t code0 = (int64_t)chunk * QPC * AP;
      // ONE barrier phase per item: armed with the byte count of everything the item needs before the first copy
      // is issued, so it completes exactly when the last byte has landed
      ptx::mbar_arrive_expect_tx(efull, code_bytes * (DO_OWN ? 2u : 1u) + 16u * A + 32u + (A + 1) * map_bytes);
      ptx::bulk_load(ptx::smem_u32(&meta[es].env), step_meta + (int64_t)b * 4 * A, 16u * A, efull);
      ptx::bulk_load(ptx::smem_u32(&meta[es].dirty[0]), st.map_flags + ((int64_t)b * cfg.n_seg + chunk) * 8, 32u, efull);
      ptx::bulk_load(edst, st.meas_codes + ((int64_t)(t & 1) * cfg.n_envs + b) * cfg.code_stride + code0, code_bytes,
                     efull);
      if (DO_OWN)
        ptx::bulk_load(edst + code_row,
                       st.meas_codes + ((int64_t)((t + 1) & 1) * cfg.n_envs + b) * cfg.code_stride + code0,
                       code_bytes, efull);
      const int64_t cell0 = (int64_t)chunk * QPC * 4;
#pragma unroll 1
      for (int m = 0; m <= A; ++m) {
        const uint32_t g = k * (A + 1) + (uint32_t)m;  // running map index; slot = g mod D_MAP
        if (g >= (uint32_t)TMA_D_MAP) {
          // the slot's previous map belongs to item (g - D_MAP) / (A + 1): wait until that item has been consumed.
          // Items <= k - D_ENV are covered by the env_done wait above; the others are younger than k - D_ENV, so
          // their env slot cannot have been recycled and the phase parity of env_tiles is unambiguous.
          const uint32_t need = (g - TMA_D_MAP) / (A + 1);
          while (released <= need) {
            ptx::mbar_wait(env_tiles + 8u * (released & (TMA_D_ENV - 1)), (released / TMA_D_ENV) & 1u);
            ++released;
          }
        }
        const float* src = (m == 0) ? st.global_map + (int64_t)b * stride + cell0
                                    : st.local_maps + ((int64_t)b * A + (m - 1)) * stride + cell0;
        ptx::bulk_load(ptx::smem_u32(map_slots + (size_t)(g & (TMA_D_MAP - 1)) * slot_bytes), src, map_bytes, efull);
      }
    }
    return;
  }

  // ==================================================================== consumer warps
  const int lane = tid & 31;
  const uint32_t my_items = (blockIdx.x < (uint32_t)n_items)
                                ? (uint32_t)(n_items - (int32_t)blockIdx.x + (int32_t)gridDim.x - 1) / gridDim.x
                                : 0u;
  const uint32_t total_tiles = my_items * NT;
  const bool kout_one = (cfg.k_out == 1.0f);
  while (true) {
    uint32_t n = 0;
    if (lane == 0) n = atomicAdd(tile_counter, 1u);
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    if (n >= total_tiles) break;
    const uint32_t k = n / NT, tile = n - k * NT;
    const uint32_t es = k & (TMA_D_ENV - 1), pe = (k / TMA_D_ENV) & 1u;
    // Tasks are handed out in item order and an env slot is recycled only after all NT tiles of its item have
    // arrived, so a warp can never be two phases behind on this barrier.
    ptx::mbar_wait(env_full + 8u * es, pe);
    StageMeta<A>& sm = meta[es];
    const int32_t ql = (int32_t)tile * 32 + lane;
    const bool have = ql < sm.nq;
    if (dbg & 1) {  // timing experiment (IPP_TMA_DEBUG): the load pipeline alone, results are NOT computed
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(env_tiles + 8u * es);
      continue;
    }
    // Lanes beyond the item's last quad (only in its last tile) run on a copy of that quad and simply do not
    // store: no divergent region in the whole task.
    const int32_t qi = min(ql, sm.nq - 1);
    const int32_t cell_q = sm.chunk * QPC + qi;  // quad index inside the whole map
    const unsigned char* code_prev = env_slots + (size_t)es * env_bytes;
    const CodeWord<A> cw = load_code<A>(code_prev, qi);
    CodeWord<A> nw;
#pragma unroll
    for (int w = 0; w < CodeWord<A>::WORDS; ++w) nw.w[w] = 0u;
    if (DO_OWN) nw = load_code<A>(code_prev + code_row, qi);
    // which local maps have work in this tile?  cells of an enabled fuse pass or of the own footprint; every quad
    // if the map may hold out-of-range odds and a fuse pass (= whole-map clamp) runs, or if k_out != 1
    // Warp-uniform facts by ballot (the compiler then knows the branches on them are uniform): bit i of all_mask
    // = local map i needs every quad; en_bits bit (i * A + j) = local map i fuses agent j's measurement.
    uint32_t all_mask;
    uint64_t en_bits;
    {
      bool a_l = false;
      if (lane < A) a_l = sm.env.comm[lane] != 0u && (((sm.dirty[lane] >> tile) & 1u) != 0u || !kout_one);
      all_mask = __ballot_sync(0xFFFFFFFFu, a_l);
      const int p0 = lane, p1 = lane + 32;
      const bool b0 = p0 < A * A && ((sm.env.comm[p0 / A] >> (p0 % A)) & 1u) != 0u;
      en_bits = __ballot_sync(0xFFFFFFFFu, b0);
      if (A * A > 32) {
        const bool b1 = p1 < A * A && ((sm.env.comm[p1 / A] >> (p1 % A)) & 1u) != 0u;
        en_bits |= (uint64_t)__ballot_sync(0xFFFFFFFFu, b1) << 32;
      }
    }
    uint32_t in_prev = 0;  // bits 4j..4j+3: cells of this quad inside agent j's communicated footprint
#pragma unroll
    for (int j = 0; j < A; ++j) in_prev |= (cw.byte(j) & 0xFu) << (4 * j);
    uint32_t mine = 0, any = 0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const bool m_i = ((all_mask >> i) & 1u) != 0u ||
                       ((in_prev & sm.env.comm4[i]) | (DO_OWN ? (nw.byte(i) & 0xFu) : 0u)) != 0u;
      mine |= (m_i ? 1u : 0u) << i;
      any |= (__any_sync(0xFFFFFFFFu, m_i) ? 1u : 0u) << i;
    }
    if (dbg & 4) any = 0u, mine = 0u;  // timing experiment: global map only
    if (!have) mine = 0u;
    const uint32_t g0 = k * (A + 1);
    const float4 g4 = reinterpret_cast<const float4*>(map_slots + (size_t)(g0 & (TMA_D_MAP - 1)) * slot_bytes)[qi];
    float4 l4[A];
    constexpr bool kWide = (A <= 4);  // registers for all of the item's quads + multipliers at once
    if (kWide) {
#pragma unroll
      for (int i = 0; i < A; ++i)
        if ((any >> i) & 1u)  // warp-uniform
          l4[i] = reinterpret_cast<const float4*>(map_slots + (size_t)((g0 + 1 + i) & (TMA_D_MAP - 1)) * slot_bytes)[qi];
    }
    // ---- global map + reward terms ----
    F4 kj[A];
    float s1 = 0.0f, s2 = 0.0f;
    float4* const out_g = reinterpret_cast<float4*>(st.global_map + (int64_t)sm.b * stride) + cell_q;
    float4* const out_l = reinterpret_cast<float4*>(st.local_maps + (int64_t)sm.b * A * stride) + cell_q;
    const int64_t stride4 = stride >> 2;
    if (!(dbg & 2)) {  // (dbg & 2: timing experiment, local maps only)
      const float4 gn = global_quad<A>(cfg, sm.env, cw, lut, g4, have ? valid_mask4(cell_q << 2, n_cells) : 0u, kj,
                                       s1, s2);
      if (have) __stcs(out_g, gn);
    } else {
#pragma unroll
      for (int j = 0; j < A; ++j) kj[j] = f4_splat(1.0f);
    }
    // ---- local maps ----
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      if (!((any >> i) & 1u)) continue;  // warp-uniform: no footprint reaches this (tile, map)
      bool b_i;
      if (kWide) {
        b_i = local_quad<A, DO_OWN>(cfg, (uint32_t)(en_bits >> (i * A)), kj, DO_OWN ? nw.byte(i) : 0u,
                                    sm.env.lut_next[i], lut, l4[i]);
      } else {  // A > 4: one map at a time, multipliers re-read from the LUT
        l4[i] = reinterpret_cast<const float4*>(map_slots + (size_t)((g0 + 1 + i) & (TMA_D_MAP - 1)) * slot_bytes)[qi];
        b_i = local_quad_lut<A, DO_OWN>(cfg, sm.env, i, cw, DO_OWN ? nw.byte(i) : 0u, lut, l4[i]);
      }
      if ((mine >> i) & 1u) {  // lanes whose quad no footprint reaches hold an unchanged copy: nothing to store
        if (b_i) bad |= 1u << i;
        __stcs(out_l + i * stride4, l4[i]);
      }
    }
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    s1 = warp_sum_f(s1);
    s2 = warp_sum_f(s2);
    if (lane == 0) {
      double* r = red + (size_t)es * 2 * NT;
      r[tile] = (double)s1;
      r[NT + tile] = (double)s2;
#pragma unroll
      for (int i = 0; i < A; ++i)
        if ((bad >> i) & 1u) atomicOr(&sm.bad[i], 1u << tile);
    }
    __syncwarp();  // every lane has read its quads out of the slots
    if (lane == 0) ptx::mbar_arrive(env_tiles + 8u * es);
  }
}

// --------------------------------------------------------------------------------------------------
static size_t stage_meta_bytes(int A) {
  switch (A) {
    case 1: return sizeof(StageMeta<1>);
    case 2: return sizeof(StageMeta<2>);
    case 3: return sizeof(StageMeta<3>);
    case 4: return sizeof(StageMeta<4>);
    case 5: return sizeof(StageMeta<5>);
    case 6: return sizeof(StageMeta<6>);
    case 7: return sizeof(StageMeta<7>);
    default: return sizeof(StageMeta<8>);
  }
}

TmaPlan plan_tma(const ipp_config& cfg, int max_smem_optin) {
  TmaPlan p;
  const int A = cfg.n_agents;
  const int ap = A <= 4 ? 4 : 8;
  const int n_quads = (cfg.gx * cfg.gy + 3) >> 2;
  p.quads_per_chunk = TMA_QPC;
  p.n_chunks = (n_quads + TMA_QPC - 1) / TMA_QPC;
  p.slot_bytes = TMA_QPC * 16;  // 10 KB, 128-byte multiple
  p.env_bytes = (2 * TMA_QPC * ap + 127) & ~127;
  p.d_env = TMA_D_ENV;
  p.d_map = TMA_D_MAP;
  p.smem_bytes = TMA_D_MAP * p.slot_bytes + TMA_D_ENV * p.env_bytes + cfg.n_alt * 256 * 16 +
                 TMA_D_ENV * (int)stage_meta_bytes(A) + 3 * TMA_D_ENV * 8 +
                 TMA_D_ENV * 2 * TMA_NT * 8 + 16 + 128;
  // one whole item + at least one slot of prefetch must fit in the map ring
  p.ok = TMA_D_MAP >= (A + 1) + 1 && p.smem_bytes <= max_smem_optin;
  return p;
}

template <int A, bool DO_OWN>
static cudaError_t launch_tma_t(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                                int n_sm, const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                                double* partials, cudaStream_t s) {
  auto kern = step_tma_kernel<A, DO_OWN>;
  static PerDevice configured;  // per template instantiation and device; grows to the largest plan seen
  int& configured_bytes = configured.cur();
  if (plan.smem_bytes > configured_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (e != cudaSuccess) return e;
    configured_bytes = plan.smem_bytes;
  }
  const int n_items = cfg.n_envs * plan.n_chunks;
  const int grid = n_items < n_sm ? n_items : n_sm;
  int dbg = 0;
  if (const char* v = getenv("IPP_TMA_DEBUG")) dbg = atoi(v);  // timing experiments only (results are wrong)
  kern<<<grid, tma_threads(A), plan.smem_bytes, s>>>(cfg, st, lut, step_meta, t, reward_rel, reward_abs, partials,
                                                   plan.n_chunks, n_items, plan.slot_bytes, plan.env_bytes, dbg);
  return cudaGetLastError();
}

cudaError_t launch_step_tma(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                            int n_sm, const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                            double* partials, bool do_own, cudaStream_t s) {
#define IPP_TMA_CASE(A_)                                                                                          \
  case A_:                                                                                                        \
    return do_own ? launch_tma_t<A_, true>(cfg, st, lut, plan, n_sm, step_meta, t, reward_rel, reward_abs,        \
                                           partials, s)                                                           \
                  : launch_tma_t<A_, false>(cfg, st, lut, plan, n_sm, step_meta, t, reward_rel, reward_abs,       \
                                            partials, s);
  switch (cfg.n_agents) {
    IPP_TMA_CASE(1)
    IPP_TMA_CASE(2)
    IPP_TMA_CASE(3)
    IPP_TMA_CASE(4)
    IPP_TMA_CASE(5)
    IPP_TMA_CASE(6)
    IPP_TMA_CASE(7)
    IPP_TMA_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef IPP_TMA_CASE
}

}  // namespace ipp

// TMA-staged, warp-specialised, persistent variant of the map kernel (sm_100a).
//
// One CTA per SM walks the (env, chunk) work items.  A chunk is 640 quads (2560 cells) = 20 TILES of 32 quads, so
// the 50x50 grid is a single chunk.  The SM takes its data in over TWO paths at once (measured, scripts/trace_tma.py:
// the bulk-copy engine of an SM moves ~20 bytes per clock, which alone caps a kernel that stages everything through
// it at the 81 us of round 1):
//   * bulk copies (cp.async.bulk, SASS UBLKCP) into shared memory for what every tile task of an item needs: the
//     item's record from the plan kernel (ItemRec: which tiles of which local map have work), its two measurement-
//     code rows and its GLOBAL map — 8 item slots, three copies per item, issued by one producer lane;
//   * plain loads (ld.global.cg.v4: L2 only — with the evict-first hint of ld.global.cs the lines the producer had
//     prefetched into L2 were gone again too often: 118 -> 115 us) straight into registers for the LOCAL-map quads, issued by the tile
//     task itself as soon as the item's record has landed and BEFORE it waits for the item's bulk data, so that
//     their latency overlaps that wait; (tile, map) pairs without work are not loaded at all.
// Roles (no block-wide barrier after start-up; mbarriers do all the synchronisation):
//   producer warp : (one lane) per item: waits until the slot's previous item is finished, then arms rec_full and
//                   env_full with their byte counts and issues the three copies.  It never waits for memory;
//   consumer warps: pull (item, tile) tasks from a shared counter (dynamic load balance); wait for the record, issue
//                   the loads of the local quads with work, wait for the bulk data, then take the tile's 32 quads
//                   through the maps in registers:
//                   * GLOBAL map: all A fuse passes — a pass whose footprint misses the tile is only a clamp, and a
//                     clamp of a value known to be in range is nothing — and its reward terms;
//                   * LOCAL maps with work: the enabled fuse passes and the own update;
//                   results go from registers straight to global memory with coalesced streaming 16-byte stores
//                   (512 contiguous bytes per warp), only for quads that can have changed;
//   finisher warp : (one lane) waits env_tiles, finishes the per-env reward from the tiles' partial sums in a
//                   fixed order, writes the local maps' new range flags and hands the slot back (env_done).
// HBM traffic is one read of the global map, of the code rows and of the local-map tiles with work, plus the write of
// the quads that changed.
#include <cstdlib>

#include "ipp_cell.cuh"
#include "ipp_launch.h"
#include "ipp_ptx.cuh"

#ifdef IPP_TMA_TRACE  // development only: per-item timeline of CTA 0 (scripts/trace_tma.py)
__device__ unsigned long long g_tma_trace[64 * 8];
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define IPP_TRACE_SET(k, e) \
  do { if (blockIdx.x == 0 && (k) < 64) g_tma_trace[(k) * 8 + (e)] = trace_now(); } while (0)
#define IPP_TRACE_MIN(k, e) \
  do { if (blockIdx.x == 0 && (k) < 64) atomicMin(&g_tma_trace[(k) * 8 + (e)], trace_now()); } while (0)
#define IPP_TRACE_MAX(k, e) \
  do { if (blockIdx.x == 0 && (k) < 64) atomicMax(&g_tma_trace[(k) * 8 + (e)], trace_now()); } while (0)
#define IPP_TRACE_ADD(k, e, v) \
  do { if (blockIdx.x == 0 && (k) < 64) atomicAdd(&g_tma_trace[(k) * 8 + (e)], (unsigned long long)(v)); } while (0)
extern "C" int ipp_debug_tma_trace(unsigned long long* out, int reset) {
  static unsigned long long init[64 * 8];
  if (reset) {
    for (int i = 0; i < 64 * 8; ++i) init[i] = ((i & 7) == 3) ? ~0ull : 0ull;
    return (int)cudaMemcpyToSymbol(g_tma_trace, init, sizeof(init));
  }
  return (int)cudaMemcpyFromSymbol(out, g_tma_trace, sizeof(init));
}
#else
#define IPP_TRACE_SET(k, e)
#define IPP_TRACE_MIN(k, e)
#define IPP_TRACE_MAX(k, e)
#define IPP_TRACE_ADD(k, e, v)
#endif

namespace ipp {

constexpr int TMA_NT = TMA_QPC / 32;  // tiles per item
constexpr int TMA_D = 8;              // item slots (power of two)
constexpr int TMA_PF = 16;            // producer's ring of prefetched item records (power of two, > TMA_D)

template <int A>
struct alignas(16) StageMeta {
  union {
    ItemRec<A> rec;                  // bulk-copied: the item's record written by the plan kernel
    uint32_t rec_raw[rec_words(A)];  // (the copy brings the record's padding words along: they must land in here)
  };
  uint32_t bad[8];  // per local map, the tiles whose results left [o_min, o_max] in this step
  int32_t b, chunk, nq, pad;
};

static_assert(TMA_QPC == IPP_FLAG_QUADS && TMA_NT == ITEM_TILES, "one range flag per (local map, work item)");

// Shared-memory layout:
//   [D][slot_bytes] global-map slots | [D][env_bytes] code rows | lut[n_alt*256] float4 | StageMeta[D] |
//   mbarriers: rec_full[D] env_full[D] env_tiles[D] env_done[D] pf_full[PF] | reward partials [D][2][NT] double |
//   tile counter | producer's record ring [PF][rec_words]
template <int A, bool DO_OWN>
__global__ void __launch_bounds__(tma_threads(A), 1)
    step_tma_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut_g,
                    const uint32_t* __restrict__ step_meta, const int32_t t, float* __restrict__ reward_rel,
                    float* __restrict__ reward_abs, double* __restrict__ partials, const int32_t n_chunks,
                    const int32_t n_items, const int32_t slot_bytes, const int32_t env_bytes, const int32_t dbg_arg) {
  extern __shared__ __align__(128) unsigned char smem[];
#ifdef IPP_TMA_TIMING_KNOBS  // scripts/dbg_bench.py: IPP_TMA_DEBUG bit 0 = load pipeline only, 1 = no global map, 2 = no
  const int32_t dbg = dbg_arg;  // local maps (results are then wrong), 3 = dense loads; not in the product build
#else
  constexpr int32_t dbg = 0;
  (void)dbg_arg;
#endif
  constexpr int NT = TMA_NT;
  constexpr int AP = A <= 4 ? 4 : 8;
  constexpr int QPC = TMA_QPC;
  constexpr int CONSUMER_THREADS = tma_consumer_warps(A) * 32;
  constexpr int TMA_THREADS = tma_threads(A);
  constexpr int RW = rec_words(A);
  unsigned char* map_slots = smem;
  unsigned char* env_slots = map_slots + (size_t)TMA_D * slot_bytes;
  float4* lut = reinterpret_cast<float4*>(env_slots + (size_t)TMA_D * env_bytes);
  StageMeta<A>* meta = reinterpret_cast<StageMeta<A>*>(lut + cfg.n_alt * 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + TMA_D);
  double* red = reinterpret_cast<double*>(bars + 4 * TMA_D + TMA_PF);  // [D][2][NT]
  uint32_t* tile_counter = reinterpret_cast<uint32_t*>(red + TMA_D * 2 * NT);
  uint32_t* pf = tile_counter + 4;  // [PF][rec_words] (16-byte aligned)
  // 32-bit shared addresses of the barrier arrays (8 bytes per barrier)
  const uint32_t rec_full = ptx::smem_u32(bars);        // the item's record has landed
  const uint32_t env_full = rec_full + 8u * TMA_D;      // ... its code rows and its global map, too
  const uint32_t env_tiles = env_full + 8u * TMA_D;     // one arrival per finished tile task
  const uint32_t env_done = env_tiles + 8u * TMA_D;     // finisher
  const uint32_t pf_full = env_done + 8u * TMA_D;       // a record of the producer's ring has landed

  const int32_t tid = threadIdx.x;
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  // code rows in a slot: row h = ping-pong half h.  A single-segment map fetches both halves with ONE copy (they
  // are adjacent in global memory), so the rows sit code_stride apart; else one copy per row and segment
  const uint32_t code_pitch = (n_chunks == 1) ? (uint32_t)cfg.code_stride : (uint32_t)QPC * AP;
  const uint32_t row_prev = (uint32_t)(t & 1) * code_pitch, row_next = (uint32_t)((t + 1) & 1) * code_pitch;

  if (tid == 0) {
    for (int s = 0; s < TMA_D; ++s) {
      ptx::mbar_init(rec_full + 8u * s, 1);
      ptx::mbar_init(env_full + 8u * s, 1);
      ptx::mbar_init(env_tiles + 8u * s, NT);
      ptx::mbar_init(env_done + 8u * s, 1);
    }
    for (int s = 0; s < TMA_PF; ++s) ptx::mbar_init(pf_full + 8u * s, 1);
    *tile_counter = 0u;
    ptx::fence_mbar_init();
  }
  for (int32_t i = tid; i < cfg.n_alt * 256; i += TMA_THREADS) lut[i] = lut_g[i];
  __syncthreads();

  const uint32_t my_items = (blockIdx.x < (uint32_t)n_items)
                                ? (uint32_t)(n_items - (int32_t)blockIdx.x + (int32_t)gridDim.x - 1) / gridDim.x
                                : 0u;

  if (tid >= CONSUMER_THREADS + 32) {
    // ================================================================== finisher warp (one lane)
    // A slot is recycled only after this lane's own arrival on env_done, so the phase parity of env_tiles can never
    // run two phases ahead of the wait.
    if (tid != CONSUMER_THREADS + 32) return;
    for (uint32_t k = 0; k < my_items; ++k) {
      const uint32_t es = k & (TMA_D - 1), pe = (k / TMA_D) & 1u;
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
        st.map_flags[((int64_t)b * cfg.n_seg + chunk) * 8 + i] =
            sm.bad[i] | (sm.rec.env.comm[i] == 0u ? sm.rec.flags[i] : 0u);
      }
      IPP_TRACE_SET(k, 5);
      ptx::mbar_arrive(env_done + 8u * es);
    }
    return;
  }

  if (tid >= CONSUMER_THREADS) {
    // ================================================================== producer warp (one lane)
    if (tid != CONSUMER_THREADS) return;
    // The producer's own copy of the item records, fetched PF - 1 items ahead: it tells which tiles of which local
    // map the item's tile tasks are going to load, and the producer pulls exactly those into L2 (bulk prefetch, no
    // destination) while the tasks are still D items away — their plain loads then see L2, not DRAM, latency
    // (same-box A/B at 8192 x 4 x 50x50: 117.7 us with, 122.0 us without).  Measured alternatives that were slower:
    // a single barrier per item (122 us), tile tasks of 2 / 4 tiles (131 / 134 us), a static deal of the tasks
    // (135 us), 4 or 12 item slots (120 / 133 us), an evict-first L2 hint on the bulk copies of the global map and the
    // code rows (118.7 vs 114.6 us), all maps through the bulk-copy engine (130 .. 150 us: the engine
    // moves ~20 bytes per clock and SM, and a bulk copy costs the issuing lane ~0.13 us; scripts/trace_tma.py).
    auto fetch_rec = [&](uint32_t kk) {
      const int32_t item = (int32_t)blockIdx.x + (int32_t)kk * (int32_t)gridDim.x;
      const uint32_t bar = pf_full + 8u * (kk & (TMA_PF - 1));
      ptx::mbar_arrive_expect_tx(bar, (uint32_t)RW * 4u);
      ptx::bulk_load(ptx::smem_u32(pf + (kk & (TMA_PF - 1)) * RW), step_meta + (int64_t)item * RW, (uint32_t)RW * 4u, bar);
    };
    for (uint32_t kk = 0; kk < (uint32_t)(TMA_PF - 1) && kk < my_items; ++kk) fetch_rec(kk);
    for (uint32_t k = 0; k < my_items; ++k) {
      const int32_t item = (int32_t)blockIdx.x + (int32_t)k * (int32_t)gridDim.x;
      const uint32_t es = k & (TMA_D - 1), pe = (k / TMA_D) & 1u;
      const int32_t b = item / n_chunks;
      const int32_t chunk = item - b * n_chunks;
      const int32_t nq = min(QPC, n_quads - chunk * QPC);
      if (!(dbg & 16)) {
        ptx::mbar_spin(pf_full + 8u * (k & (TMA_PF - 1)), (k / TMA_PF) & 1u);
        const uint32_t* e = pf + (k & (TMA_PF - 1)) * RW;
        const int32_t nt = (nq + 31) >> 5;
#pragma unroll 1
        for (int i = 0; i < A; ++i) {
          const uint32_t need = e[4 * A + i];
          if (need == 0u) continue;
          const int first = __ffs(need) - 1, last = 31 - __clz(need);
          uint32_t bytes = (uint32_t)(last - first + 1) * 512u;
          if (last == nt - 1) bytes -= 512u - (uint32_t)(nq - 32 * (nt - 1)) * 16u;
          ptx::bulk_prefetch_l2(st.local_maps + ((int64_t)b * A + i) * stride + (int64_t)chunk * QPC * 4 + first * 128,
                                bytes);
        }
      }
      if (k + TMA_PF - 1 < my_items) fetch_rec(k + TMA_PF - 1);  // its ring entry was item k - 1's
      ptx::mbar_spin(env_done + 8u * es, pe ^ 1u);  // the slot's previous item (k - D) is finished
      IPP_TRACE_SET(k, 0);
      StageMeta<A>& sm = meta[es];
      sm.b = b;
      sm.chunk = chunk;
      sm.nq = nq;
#pragma unroll
      for (int i = 0; i < 8; ++i) sm.bad[i] = 0u;
      const uint32_t rfull = rec_full + 8u * es, efull = env_full + 8u * es;
      ptx::mbar_arrive_expect_tx(rfull, (uint32_t)RW * 4u);
      ptx::bulk_load(ptx::smem_u32(&sm.rec), step_meta + (int64_t)item * RW, (uint32_t)RW * 4u, rfull);
      const uint32_t code_bytes = ((uint32_t)nq * AP + 15u) & ~15u;
      const uint32_t map_bytes = (uint32_t)nq * 16u;
      const bool both_rows = DO_OWN && n_chunks == 1;  // the two code halves are adjacent: one copy
      const uint32_t codes_total = both_rows ? 2u * (uint32_t)cfg.code_stride : code_bytes * (DO_OWN ? 2u : 1u);
      // ONE barrier phase for the item's bulk data: armed with its byte count before the first copy is issued
      ptx::mbar_arrive_expect_tx(efull, codes_total + map_bytes);
      const uint32_t edst = ptx::smem_u32(env_slots + (size_t)es * env_bytes);
      if (both_rows) {
        ptx::bulk_load(edst, st.meas_codes + code_row_offset(cfg, 0, b), codes_total, efull);
      } else {
        const int64_t code0 = (int64_t)chunk * QPC * AP;
        ptx::bulk_load(edst + row_prev, st.meas_codes + code_row_offset(cfg, t, b) + code0, code_bytes, efull);
        if (DO_OWN)
          ptx::bulk_load(edst + row_next, st.meas_codes + code_row_offset(cfg, t + 1, b) + code0, code_bytes, efull);
      }
      ptx::bulk_load(ptx::smem_u32(map_slots + (size_t)es * slot_bytes),
                     st.global_map + (int64_t)b * stride + (int64_t)chunk * QPC * 4, map_bytes, efull);
      IPP_TRACE_SET(k, 2);
    }
    return;
  }

  // ==================================================================== consumer warps
  const int lane = tid & 31;
  const uint32_t total_tiles = my_items * NT;
  // (item, tile) tasks in item order from a shared counter: the tasks differ in cost (0 .. A local maps with work), and
  // a static round-robin deal was measured 8 % (A = 4) to 30 % (A = 8) slower than this dynamic one
  while (true) {
    uint32_t n = 0;
    if (lane == 0) n = atomicAdd(tile_counter, 1u);
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    if (n >= total_tiles) break;
    const uint32_t k = n / NT, tile = n - k * NT;
    const uint32_t es = k & (TMA_D - 1), pe = (k / TMA_D) & 1u;
    // Tasks are handed out in item order and a slot is recycled only after all NT tiles of its item have arrived, so
    // a warp can never be two phases behind on these barriers.
#ifdef IPP_TMA_TRACE
    const unsigned long long tw0 = trace_now();
#endif
    ptx::mbar_wait(rec_full + 8u * es, pe);
#ifdef IPP_TMA_TRACE
    if (lane == 0) IPP_TRACE_ADD(k, 6, trace_now() - tw0);
#endif
    StageMeta<A>& sm = meta[es];
    const int32_t nq = sm.nq;  // (written by the producer before it armed rec_full)
    const int32_t ql = (int32_t)tile * 32 + lane;
    const bool have = ql < nq;
    if ((int32_t)tile * 32 >= nq) {  // tile beyond the item's last quad (short last item of a map): nothing to do
      if (lane == 0) {
        double* r = red + (size_t)es * 2 * NT;
        r[tile] = 0.0;
        r[NT + tile] = 0.0;
        ptx::mbar_arrive(env_tiles + 8u * es);
      }
      continue;
    }
    // Lanes beyond the item's last quad (only in its last tile) run on a copy of that quad and simply do not
    // store: no divergent region in the whole task.
    const int32_t qi = min(ql, nq - 1);
    const int32_t cell_q = sm.chunk * QPC + qi;  // quad index inside the whole map
    float4* const out_g = reinterpret_cast<float4*>(st.global_map + (int64_t)sm.b * stride) + cell_q;
    float4* const out_l = reinterpret_cast<float4*>(st.local_maps + (int64_t)sm.b * A * stride) + cell_q;
    const int64_t stride4 = stride >> 2;
    // warp-uniform facts of the tile (ItemRec) — and the loads of the local quads with work, issued now so that they
    // fly while this warp waits for the item's bulk data
    const uint32_t tf = sm.rec.tile[tile];
    const uint32_t touch = tf & 0xFFu, dirty = tf >> 8;
    uint32_t work = 0;  // bit i: local map i has work in this tile
#pragma unroll
    for (int i = 0; i < A; ++i) work |= (((dbg & 8) ? 1u : (sm.rec.need[i] >> tile)) & 1u) << i;
    if (dbg & 4) work = 0u;
    float4 l4[A];
    constexpr bool kWide = (A <= 4);  // registers for all of the item's quads at once
    if (kWide) {
#pragma unroll
      for (int i = 0; i < A; ++i)
        if ((work >> i) & 1u) l4[i] = __ldcg(out_l + i * stride4);  // warp-uniform branch
    }
    ptx::mbar_wait(env_full + 8u * es, pe);
    if (lane == 0) IPP_TRACE_MIN(k, 3);
    if (dbg & 1) {  // timing experiment (IPP_TMA_DEBUG): the load pipeline alone, results are NOT computed
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(env_tiles + 8u * es);
      continue;
    }
    const unsigned char* code_slot = env_slots + (size_t)es * env_bytes;
    const CodeWord<A> cw = load_code<A>(code_slot + row_prev, qi);
    CodeWord<A> nw;
#pragma unroll
    for (int w = 0; w < CodeWord<A>::WORDS; ++w) nw.w[w] = 0u;
    if (DO_OWN) nw = load_code<A>(code_slot + row_next, qi);
    const float4 g4 = reinterpret_cast<const float4*>(map_slots + (size_t)es * slot_bytes)[qi];
    uint32_t in_prev = 0;  // bits 4j..4j+3: cells of this quad inside agent j's communicated footprint
#pragma unroll
    for (int j = 0; j < A; ++j) in_prev |= (cw.byte(j) & 0xFu) << (4 * j);
    // ---- global map + reward terms ----
    float s1 = 0.0f, s2 = 0.0f;
    if (!(dbg & 2)) {  // (dbg & 2: timing experiment, local maps only)
      const float4 gn = global_quad<A>(cfg, sm.rec.env, cw, lut, g4, have ? valid_mask4(cell_q << 2, n_cells) : 0u,
                                       touch, s1, s2);
      // an untouched, in-range quad comes out bit-identical: nothing to write
      if (have && quad_changed(gn, g4)) __stcs(out_g, gn);
    }
    // ---- local maps ----
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      if (!((work >> i) & 1u)) continue;  // warp-uniform: no work for this (tile, map); it was not loaded
      if (!kWide) l4[i] = __ldcg(out_l + i * stride4);  // A > 4: one map at a time
      const bool tile_dirty = ((dirty >> i) & 1u) != 0u;
      const bool b_i = local_quad<A, DO_OWN>(cfg, sm.rec.env, sm.rec.env.comm[i], touch, tile_dirty, i, cw,
                                             DO_OWN ? nw.byte(i) : 0u, lut, l4[i]);
      // only a quad inside an enabled or the own footprint (or any quad of a tile that had to be clamped) can have
      // changed; the others hold the bits they were loaded with: nothing to store
      const bool mine = tile_dirty || ((in_prev & sm.rec.env.comm4[i]) | (DO_OWN ? (nw.byte(i) & 0xFu) : 0u)) != 0u;
      if (have && mine) {
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
    }
    if (lane < A && ((bad >> lane) & 1u)) atomicOr(&sm.bad[lane], 1u << tile);
    __syncwarp();  // partial sums and bad bits of all lanes are in place, the slot has been read
    if (lane == 0) {
      IPP_TRACE_MAX(k, 4);
      ptx::mbar_arrive(env_tiles + 8u * es);
    }
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
  p.slot_bytes = TMA_QPC * 16;  // one segment of the global map: 10 KB, 128-byte multiple
  p.env_bytes = (2 * TMA_QPC * ap + 127) & ~127;
  p.d_env = TMA_D;
  p.d_map = TMA_D;
  p.smem_bytes = TMA_D * (p.slot_bytes + p.env_bytes) + cfg.n_alt * 256 * 16 + TMA_D * (int)stage_meta_bytes(A) +
                 (4 * TMA_D + TMA_PF) * 8 + TMA_D * 2 * TMA_NT * 8 + 16 + TMA_PF * rec_words(A) * 4 + 128;
  // tile numbers are 16-bit in the plan kernel's footprint ranges
  p.ok = p.smem_bytes <= max_smem_optin && (n_quads + 31) / 32 <= 0xFFFF;
  return p;
}

template <int A, bool DO_OWN>
static cudaError_t launch_tma_t(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                                int n_sm, const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                                double* partials, cudaStream_t s) {
  auto kern = step_tma_kernel<A, DO_OWN>;
  static PerDevice configured;  // per template instantiation and device (normally done by configure_step_tma)
  if (!configured.cur()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured.cur() = 1;
  }
  const int n_items = cfg.n_envs * plan.n_chunks;
  const int grid = n_items < n_sm ? n_items : n_sm;
  int dbg = 0;
#ifdef IPP_TMA_TIMING_KNOBS  // timing experiments only (results are wrong); not in the product build
  static const int dbg_env = getenv("IPP_TMA_DEBUG") ? atoi(getenv("IPP_TMA_DEBUG")) : 0;
  dbg = dbg_env;
#endif
  kern<<<grid, tma_threads(A), plan.smem_bytes, s>>>(cfg, st, lut, step_meta, t, reward_rel, reward_abs, partials,
                                                   plan.n_chunks, n_items, plan.slot_bytes, plan.env_bytes, dbg);
  return cudaGetLastError();
}

// dynamic shared-memory limit of both instantiations (with / without the own update) for this shape on the current
// device; called from ipp_create so that nothing has to be configured while a stream is being captured
template <int A>
static cudaError_t configure_tma_t(const TmaPlan& plan) {
  cudaError_t e = cudaFuncSetAttribute(step_tma_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  (void)plan;
  return cudaFuncSetAttribute(step_tma_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t configure_step_tma(const ipp_config& cfg, const TmaPlan& plan) {
  switch (cfg.n_agents) {
    case 1: return configure_tma_t<1>(plan);
    case 2: return configure_tma_t<2>(plan);
    case 3: return configure_tma_t<3>(plan);
    case 4: return configure_tma_t<4>(plan);
    case 5: return configure_tma_t<5>(plan);
    case 6: return configure_tma_t<6>(plan);
    case 7: return configure_tma_t<7>(plan);
    case 8: return configure_tma_t<8>(plan);
    default: return cudaErrorInvalidValue;
  }
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

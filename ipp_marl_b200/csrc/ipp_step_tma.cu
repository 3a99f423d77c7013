// TMA-staged, warp-specialised, persistent variant of the map kernel (sm_100a).
//
// One CTA per SM walks the (env, chunk) work items.  A chunk is 640 quads (2560 cells): one quad per
// consumer thread, so the 50x50 grid is a single chunk.  Shared memory holds two rings:
//   * MAP slots (10 KB each, ~16 of them): one belief map of one item per slot;
//   * ENV slots (4): the item's two measurement-code rows + its EnvMeta + the reward partial sums.
// Roles (no block-wide barrier anywhere after start-up; everything is mbarrier based):
//   producer warp : per item, writes EnvMeta, bulk-loads the code rows (-> env_full) and then the
//                   item's A+1 maps, each into the next free map slot (cp.async.bulk, SASS UBLKCP,
//                   completion on map_full via expect_tx / complete_tx);
//   20 consumer warps : thread q owns quad q of the item for ALL its maps: it decodes the codes once
//                   (QuadCtx in registers), then for each map waits map_full, updates its quad in
//                   place (ipp_cell.cuh) and the warp arrives on map_done — a warp never waits for
//                   the other warps;
//   storer warp   : waits map_done, writes the slot back with cp.async.bulk shared->global, frees the
//                   slot (map_empty) once the bulk engine has read it, and finishes the per-env
//                   reward from the warps' partial sums.
// HBM traffic is one read + one write of every belief map plus the code rows; all global addressing
// is done by the TMA unit, so the SM issue slots go to the map arithmetic.
#include "ipp_cell.cuh"
#include "ipp_launch.h"

namespace ipp {

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace ptx

template <int A>
struct StageMeta {
  EnvMeta<A> env;
  int32_t b, chunk, nq, pad;
};

// Ring position helper: slot index + phase parity of a monotonically increasing counter.
struct Ring {
  int32_t slot;
  uint32_t phase;
  int32_t depth;
  __device__ __forceinline__ Ring(int32_t d) : slot(0), phase(0), depth(d) {}
  __device__ __forceinline__ void advance() {
    if (++slot == depth) {
      slot = 0;
      phase ^= 1u;
    }
  }
};

constexpr int TMA_STORE_LAG = 2;  // bulk-store groups allowed in flight before a map slot is recycled

// Shared-memory layout:
//   [d_map][slot_bytes] map slots | [d_env][env_bytes] code rows | lut[n_alt*256] float4 | StageMeta[d_env] |
//   mbarriers: map_full[d_map] map_done[d_map] map_empty[d_map] env_full[d_env] env_done[d_env] |
//   reward partials [d_env][2][NW] double
template <int A, bool DO_OWN>
__global__ void __launch_bounds__(TMA_THREADS, 1)
    step_tma_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut_g,
                    const int32_t* __restrict__ pos_in, const int32_t* __restrict__ pos_out,
                    const uint8_t* __restrict__ comm, const int32_t t, float* __restrict__ reward_rel,
                    float* __restrict__ reward_abs, double* __restrict__ partials, const int32_t n_chunks,
                    const int32_t n_items, const int32_t slot_bytes, const int32_t env_bytes, const int32_t d_map,
                    const int32_t d_env) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = TMA_CONSUMERS / 32;
  constexpr int AP = A <= 4 ? 4 : 8;
  constexpr int QPC = TMA_CONSUMERS;  // quads per chunk = one per consumer thread
  unsigned char* map_slots = smem;
  unsigned char* env_slots = map_slots + (size_t)d_map * slot_bytes;
  float4* lut = reinterpret_cast<float4*>(env_slots + (size_t)d_env * env_bytes);
  StageMeta<A>* meta = reinterpret_cast<StageMeta<A>*>(lut + cfg.n_alt * 256);
  uint64_t* map_full = reinterpret_cast<uint64_t*>(meta + d_env);
  uint64_t* map_done = map_full + d_map;
  uint64_t* map_empty = map_done + d_map;
  uint64_t* env_full = map_empty + d_map;
  uint64_t* env_done = env_full + d_env;
  double* red = reinterpret_cast<double*>(env_done + d_env);  // [d_env][2][NW]

  const int32_t tid = threadIdx.x;
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  const uint32_t code_row = (uint32_t)QPC * AP;

  if (tid == 0) {
    for (int s = 0; s < d_map; ++s) {
      ptx::mbar_init(ptx::smem_u32(&map_full[s]), 1);    // producer's arrive.expect_tx
      ptx::mbar_init(ptx::smem_u32(&map_done[s]), NW);   // one arrival per consumer warp
      ptx::mbar_init(ptx::smem_u32(&map_empty[s]), 1);   // storer
    }
    for (int s = 0; s < d_env; ++s) {
      ptx::mbar_init(ptx::smem_u32(&env_full[s]), 1);        // producer
      ptx::mbar_init(ptx::smem_u32(&env_done[s]), NW + 1);   // consumer warps + storer
    }
    ptx::fence_mbar_init();
  }
  for (int32_t i = tid; i < cfg.n_alt * 256; i += TMA_THREADS) lut[i] = lut_g[i];
  __syncthreads();

  if (tid >= TMA_CONSUMERS + 32) {
    // ================================================================== storer warp (one lane)
    if (tid != TMA_CONSUMERS + 32) return;
    Ring er(d_env), mr(d_map), freed(d_map);
    int32_t n_committed = 0, n_freed = 0;  // bulk-store groups committed / map slots handed back
    for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      ptx::mbar_wait(ptx::smem_u32(&env_full[er.slot]), er.phase);
      const StageMeta<A>& sm = meta[er.slot];
      const int32_t b = sm.b, chunk = sm.chunk;
      const uint32_t map_bytes = (uint32_t)sm.nq * 16u;
      const int64_t cell0 = (int64_t)chunk * QPC * 4;
#pragma unroll 1
      for (int m = 0; m <= A; ++m) {
        ptx::mbar_wait(ptx::smem_u32(&map_done[mr.slot]), mr.phase);
        float* dst;
        if (m == 0) {
          const double* r = red + (size_t)er.slot * 2 * NW;
          double t1 = 0.0, t2 = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            t1 += r[w];
            t2 += r[NW + w];
          }
          if (n_chunks == 1) {
            write_rewards(reward_rel, reward_abs, b, t1, t2, n_cells);
          } else {
            partials[((int64_t)b * n_chunks + chunk) * 2 + 0] = t1;
            partials[((int64_t)b * n_chunks + chunk) * 2 + 1] = t2;
          }
          dst = st.global_map + (int64_t)b * stride + cell0;
        } else {
          dst = st.local_maps + ((int64_t)b * A + (m - 1)) * stride + cell0;
        }
        ptx::bulk_store(dst, ptx::smem_u32(map_slots + (size_t)mr.slot * slot_bytes), map_bytes);
        ptx::bulk_commit();
        ++n_committed;
        mr.advance();
        ptx::bulk_wait_read<TMA_STORE_LAG>();  // all but the newest LAG groups have been read out of smem
        while (n_freed < n_committed - TMA_STORE_LAG) {
          ptx::mbar_arrive(ptx::smem_u32(&map_empty[freed.slot]));
          freed.advance();
          ++n_freed;
        }
      }
      ptx::mbar_arrive(ptx::smem_u32(&env_done[er.slot]));
      er.advance();
    }
    ptx::bulk_wait_read<0>();
    ptx::bulk_wait<0>();  // all writes to global memory complete before the CTA retires
    return;
  }

  if (tid >= TMA_CONSUMERS) {
    // ================================================================== producer warp
    const int lane = tid - TMA_CONSUMERS;
    Ring er(d_env), mr(d_map);
    for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      ptx::mbar_wait(ptx::smem_u32(&env_done[er.slot]), er.phase ^ 1u);
      const int32_t b = item / n_chunks;
      const int32_t chunk = item - b * n_chunks;
      const int32_t nq = min(QPC, n_quads - chunk * QPC);
      load_env_meta<A>(cfg, &meta[er.slot].env, lane, b, pos_in, pos_out, comm, DO_OWN);
      if (lane == 0) {
        meta[er.slot].b = b;
        meta[er.slot].chunk = chunk;
        meta[er.slot].nq = nq;
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t efull = ptx::smem_u32(&env_full[er.slot]);
        const uint32_t code_bytes = ((uint32_t)nq * AP + 15u) & ~15u;
        const uint32_t edst = ptx::smem_u32(env_slots + (size_t)er.slot * env_bytes);
        const int64_t code0 = (int64_t)chunk * QPC * AP;
        ptx::mbar_arrive_expect_tx(efull, code_bytes * (DO_OWN ? 2u : 1u));
        ptx::bulk_load(edst, st.meas_codes + ((int64_t)(t & 1) * cfg.n_envs + b) * cfg.code_stride + code0,
                       code_bytes, efull);
        if (DO_OWN)
          ptx::bulk_load(edst + code_row,
                         st.meas_codes + ((int64_t)((t + 1) & 1) * cfg.n_envs + b) * cfg.code_stride + code0,
                         code_bytes, efull);
        const uint32_t map_bytes = (uint32_t)nq * 16u;
        const int64_t cell0 = (int64_t)chunk * QPC * 4;
#pragma unroll 1
        for (int m = 0; m <= A; ++m) {
          ptx::mbar_wait(ptx::smem_u32(&map_empty[mr.slot]), mr.phase ^ 1u);
          const uint32_t mfull = ptx::smem_u32(&map_full[mr.slot]);
          const float* src = (m == 0) ? st.global_map + (int64_t)b * stride + cell0
                                      : st.local_maps + ((int64_t)b * A + (m - 1)) * stride + cell0;
          ptx::mbar_arrive_expect_tx(mfull, map_bytes);
          ptx::bulk_load(ptx::smem_u32(map_slots + (size_t)mr.slot * slot_bytes), src, map_bytes, mfull);
          mr.advance();
        }
      }
      er.advance();
    }
    return;
  }

  // ==================================================================== consumer warps
  const int lane = tid & 31, warp = tid >> 5;
  Ring er(d_env), mr(d_map);
  for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    ptx::mbar_wait(ptx::smem_u32(&env_full[er.slot]), er.phase);
    const StageMeta<A>& sm = meta[er.slot];
    const bool have = tid < sm.nq;
    const unsigned char* code_prev = env_slots + (size_t)er.slot * env_bytes;
    const unsigned char* code_next = code_prev + code_row;
    QuadCtx<A> qc;
    CodeWord<A> next;
    uint32_t valid = 0;
    if (have) {
      make_quad_ctx<A>(cfg, sm.env, load_code<A>(code_prev, tid), lut, qc);
      if (DO_OWN) next = load_code<A>(code_next, tid);
      valid = valid_mask4((sm.chunk * QPC + tid) << 2, n_cells);
    }
    // ---- global map ----
    {
      ptx::mbar_wait(ptx::smem_u32(&map_full[mr.slot]), mr.phase);
      double s1 = 0.0, s2 = 0.0;
      if (have) {
        float4* mp = reinterpret_cast<float4*>(map_slots + (size_t)mr.slot * slot_bytes) + tid;
        *mp = update_global_quad<A>(cfg, qc, *mp, valid, s1, s2);
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        double* r = red + (size_t)er.slot * 2 * NW;
        r[warp] = s1;
        r[NW + warp] = s2;
      }
      ptx::fence_proxy_async();  // my shared-memory writes -> visible to the bulk-copy (async) proxy
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&map_done[mr.slot]));
      mr.advance();
    }
    // ---- local maps ----
#pragma unroll
    for (int i = 0; i < A; ++i) {
      ptx::mbar_wait(ptx::smem_u32(&map_full[mr.slot]), mr.phase);
      if (have) {
        float4* mp = reinterpret_cast<float4*>(map_slots + (size_t)mr.slot * slot_bytes) + tid;
        *mp = update_local_quad<A, DO_OWN>(cfg, sm.env, qc, i, DO_OWN ? next.byte(i) : 0u, lut, *mp);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&map_done[mr.slot]));
      mr.advance();
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&env_done[er.slot]));
    er.advance();
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
  p.quads_per_chunk = TMA_CONSUMERS;
  p.n_chunks = (n_quads + TMA_CONSUMERS - 1) / TMA_CONSUMERS;
  p.slot_bytes = TMA_CONSUMERS * 16;                              // 10 KB, 128-byte multiple
  p.env_bytes = (2 * TMA_CONSUMERS * ap + 127) & ~127;
  p.d_env = 4;
  const int nw = TMA_CONSUMERS / 32;
  const int fixed = cfg.n_alt * 256 * 16 + p.d_env * (p.env_bytes + (int)stage_meta_bytes(A) + 2 * nw * 8 + 16) + 256;
  int d_map = (max_smem_optin - fixed) / (p.slot_bytes + 3 * 8);
  if (d_map > 24) d_map = 24;
  p.d_map = d_map;
  p.smem_bytes = fixed + d_map * (p.slot_bytes + 3 * 8);
  // at least one whole item + the store lag + one slot of prefetch
  p.ok = d_map >= (A + 1) + TMA_STORE_LAG + 1 && p.smem_bytes <= max_smem_optin;
  return p;
}

template <int A, bool DO_OWN>
static cudaError_t launch_tma_t(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                                int n_sm, const int32_t* pos_in, const int32_t* pos_out, const uint8_t* comm,
                                int32_t t, float* reward_rel, float* reward_abs, double* partials, cudaStream_t s) {
  auto kern = step_tma_kernel<A, DO_OWN>;
  static int configured_bytes = 0;  // per template instantiation; grows to the largest plan seen
  if (plan.smem_bytes > configured_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (e != cudaSuccess) return e;
    configured_bytes = plan.smem_bytes;
  }
  const int n_items = cfg.n_envs * plan.n_chunks;
  const int grid = n_items < n_sm ? n_items : n_sm;
  kern<<<grid, TMA_THREADS, plan.smem_bytes, s>>>(cfg, st, lut, pos_in, pos_out, comm, t, reward_rel, reward_abs,
                                                   partials, plan.n_chunks, n_items, plan.slot_bytes, plan.env_bytes,
                                                   plan.d_map, plan.d_env);
  return cudaGetLastError();
}

cudaError_t launch_step_tma(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                            int n_sm, const int32_t* pos_in, const int32_t* pos_out, const uint8_t* comm, int32_t t,
                            float* reward_rel, float* reward_abs, double* partials, bool do_own, cudaStream_t s) {
#define IPP_TMA_CASE(A_)                                                                                       \
  case A_:                                                                                                     \
    return do_own ? launch_tma_t<A_, true>(cfg, st, lut, plan, n_sm, pos_in, pos_out, comm, t, reward_rel,     \
                                           reward_abs, partials, s)                                            \
                  : launch_tma_t<A_, false>(cfg, st, lut, plan, n_sm, pos_in, pos_out, comm, t, reward_rel,    \
                                            reward_abs, partials, s);
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

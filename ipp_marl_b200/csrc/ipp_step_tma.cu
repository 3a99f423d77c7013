// TMA-staged, warp-specialised, persistent variant of the map kernel (sm_100a).
//
// One CTA per SM walks the (env, chunk) work items with a 3-4 stage shared-memory ring:
//   producer warp : per item, issues cp.async.bulk (TMA 1-D bulk copies, SASS UBLKCP) of the env's global
//                   map, its A local maps and its two measurement-code rows into the stage; completion is
//                   signalled on an mbarrier (expect_tx / complete_tx).
//   20 consumer warps : update the maps in place in shared memory (ipp_cell.cuh), reduce the two
//                   reward sums (warp shuffles + one named barrier), then one thread writes the maps
//                   back with cp.async.bulk shared->global and releases the stage once the bulk
//                   engine has finished reading it.
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
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
}  // namespace ptx

template <int A>
struct StageMeta {
  EnvMeta<A> env;
  int32_t b, chunk, nq, pad;
};

// Shared-memory layout: [n_stages][stage_bytes] | lut[n_alt*256] float4 | StageMeta[MAX] | mbarriers | reduction
// One stage: (A+1) maps of qpc float4 | qpc*AP bytes codes (communicated) | qpc*AP bytes codes (after move)
template <int A, bool DO_OWN>
__global__ void __launch_bounds__(TMA_THREADS, 1)
    step_tma_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const float4* __restrict__ lut_g,
                    const int32_t* __restrict__ pos_in, const int32_t* __restrict__ pos_out,
                    const uint8_t* __restrict__ comm, const int32_t t, float* __restrict__ reward_rel,
                    float* __restrict__ reward_abs, double* __restrict__ partials, const int32_t n_chunks,
                    const int32_t qpc, const int32_t n_items, const int32_t stage_bytes, const int32_t n_stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = TMA_CONSUMERS / 32;
  constexpr int AP = A <= 4 ? 4 : 8;
  unsigned char* stages = smem;
  float4* lut = reinterpret_cast<float4*>(smem + (size_t)n_stages * stage_bytes);
  StageMeta<A>* meta = reinterpret_cast<StageMeta<A>*>(lut + cfg.n_alt * 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + TMA_MAX_STAGES);  // full[MAX], empty[MAX]
  double* red = reinterpret_cast<double*>(bars + 2 * TMA_MAX_STAGES);   // [2 parities][2 sums][NW]

  const int32_t tid = threadIdx.x;
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  const uint32_t code_off = (uint32_t)(A + 1) * (uint32_t)qpc * 16u;  // offset of the code rows in a stage
  const uint32_t code_row = (uint32_t)qpc * AP;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars[s]), 1);                   // full: producer's arrive.expect_tx
      ptx::mbar_init(ptx::smem_u32(&bars[TMA_MAX_STAGES + s]), 1);  // empty: consumer thread 0
    }
    ptx::fence_mbar_init();
  }
  for (int32_t i = tid; i < cfg.n_alt * 256; i += TMA_THREADS) lut[i] = lut_g[i];
  __syncthreads();

  if (tid >= TMA_CONSUMERS) {
    // ------------------------------------------------------------------ producer warp
    const int lane = tid - TMA_CONSUMERS;
    int32_t k = 0, s = 0;
    uint32_t ph = 0;
    for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
      ptx::mbar_wait(ptx::smem_u32(&bars[TMA_MAX_STAGES + s]), ph ^ 1u);
      const int32_t b = item / n_chunks;
      const int32_t chunk = item - b * n_chunks;
      const int32_t nq = min(qpc, n_quads - chunk * qpc);
      load_env_meta<A>(cfg, &meta[s].env, lane, b, pos_in, pos_out, comm, DO_OWN);
      if (lane == 0) {
        meta[s].b = b;
        meta[s].chunk = chunk;
        meta[s].nq = nq;
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t full = ptx::smem_u32(&bars[s]);
        const uint32_t map_bytes = (uint32_t)nq * 16u;
        const uint32_t code_bytes = ((uint32_t)nq * AP + 15u) & ~15u;
        ptx::mbar_arrive_expect_tx(full, map_bytes * (A + 1) + code_bytes * (DO_OWN ? 2u : 1u));
        const uint32_t dst = ptx::smem_u32(stages + (size_t)s * stage_bytes);
        const int64_t cell0 = (int64_t)chunk * qpc * 4;
        ptx::bulk_load(dst, st.global_map + (int64_t)b * stride + cell0, map_bytes, full);
#pragma unroll
        for (int i = 0; i < A; ++i)
          ptx::bulk_load(dst + (uint32_t)(1 + i) * (uint32_t)qpc * 16u,
                         st.local_maps + ((int64_t)b * A + i) * stride + cell0, map_bytes, full);
        const int64_t code0 = (int64_t)chunk * qpc * AP;
        ptx::bulk_load(dst + code_off,
                       st.meas_codes + ((int64_t)(t & 1) * cfg.n_envs + b) * cfg.code_stride + code0, code_bytes, full);
        if (DO_OWN)
          ptx::bulk_load(dst + code_off + code_row,
                         st.meas_codes + ((int64_t)((t + 1) & 1) * cfg.n_envs + b) * cfg.code_stride + code0,
                         code_bytes, full);
      }
      if (++s == n_stages) { s = 0; ph ^= 1u; }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  int32_t k = 0, s = 0, s_prev = 0;
  uint32_t ph = 0;
  for (int32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
    ptx::mbar_wait(ptx::smem_u32(&bars[s]), ph);
    const StageMeta<A>& sm = meta[s];
    const int32_t b = sm.b, chunk = sm.chunk, nq = sm.nq;
    unsigned char* base = stages + (size_t)s * stage_bytes;
    float4* maps = reinterpret_cast<float4*>(base);
    const unsigned char* code_prev = base + code_off;
    const unsigned char* code_next = base + code_off + code_row;

    double s1 = 0.0, s2 = 0.0;
    for (int32_t ql = tid; ql < nq; ql += TMA_CONSUMERS) {
      QuadCtx<A> qc;
      make_quad_ctx<A>(cfg, sm.env, load_code<A>(code_prev, ql), lut, qc);
      CodeWord<A> next;
      if (DO_OWN) next = load_code<A>(code_next, ql);
      maps[ql] = update_global_quad<A>(cfg, qc, maps[ql], valid_mask4((chunk * qpc + ql) << 2, n_cells), s1, s2);
#pragma unroll
      for (int i = 0; i < A; ++i) {
        float4* mp = maps + (size_t)(1 + i) * qpc + ql;
        *mp = update_local_quad<A, DO_OWN>(cfg, sm.env, qc, i, DO_OWN ? next.byte(i) : 0u, lut, *mp);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    double* r = red + (size_t)(k & 1) * 2 * NW;
    if ((tid & 31) == 0) {
      r[tid >> 5] = s1;
      r[NW + (tid >> 5)] = s2;
    }
    ptx::fence_proxy_async();  // my shared-memory writes -> visible to the bulk-copy (async) proxy
    ptx::named_bar_sync(1, TMA_CONSUMERS);
    if (tid == 0) {
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
      const uint32_t src = ptx::smem_u32(base);
      const uint32_t map_bytes = (uint32_t)nq * 16u;
      const int64_t cell0 = (int64_t)chunk * qpc * 4;
      ptx::bulk_store(st.global_map + (int64_t)b * stride + cell0, src, map_bytes);
#pragma unroll
      for (int i = 0; i < A; ++i)
        ptx::bulk_store(st.local_maps + ((int64_t)b * A + i) * stride + cell0,
                        src + (uint32_t)(1 + i) * (uint32_t)qpc * 16u, map_bytes);
      ptx::bulk_commit();
      ptx::bulk_wait_read<1>();  // the previous item's stores have finished reading their stage
      if (k > 0) ptx::mbar_arrive(ptx::smem_u32(&bars[TMA_MAX_STAGES + s_prev]));
    }
    s_prev = s;
    if (++s == n_stages) { s = 0; ph ^= 1u; }
  }
  if (tid == 0) {
    ptx::bulk_wait_read<0>();
    if (k > 0) ptx::mbar_arrive(ptx::smem_u32(&bars[TMA_MAX_STAGES + s_prev]));
    ptx::bulk_wait<0>();  // all writes to global memory complete before the CTA retires
  }
}

// --------------------------------------------------------------------------------------------------
static size_t tma_fixed_smem(int A, int n_alt) {
  size_t meta = 0;
  switch (A) {
    case 1: meta = sizeof(StageMeta<1>); break;
    case 2: meta = sizeof(StageMeta<2>); break;
    case 3: meta = sizeof(StageMeta<3>); break;
    case 4: meta = sizeof(StageMeta<4>); break;
    case 5: meta = sizeof(StageMeta<5>); break;
    case 6: meta = sizeof(StageMeta<6>); break;
    case 7: meta = sizeof(StageMeta<7>); break;
    default: meta = sizeof(StageMeta<8>); break;
  }
  return (size_t)n_alt * 256 * sizeof(float4) + TMA_MAX_STAGES * meta + 2 * TMA_MAX_STAGES * sizeof(uint64_t) +
         2 * 2 * (TMA_CONSUMERS / 32) * sizeof(double) + 128;
}

TmaPlan plan_tma(const ipp_config& cfg, int max_smem_optin) {
  TmaPlan p;
  const int A = cfg.n_agents;
  const int ap = A <= 4 ? 4 : 8;
  const int n_quads = (cfg.gx * cfg.gy + 3) >> 2;
  const int per_quad = 16 * (A + 1) + 2 * ap;
  const int fixed = (int)tma_fixed_smem(A, cfg.n_alt);
  const int avail = max_smem_optin - fixed;
  // whole env in one stage if 3 stages of it fit (per-env reward finishes in the CTA); else 4 stages of chunks
  int stages = TMA_MAX_STAGES;
  int qpc = ((avail / stages) / per_quad) & ~3;
  const int whole = (n_quads + 3) & ~3;
  if (qpc >= whole) {
    qpc = whole;
  } else if ((((avail / 3) / per_quad) & ~3) >= whole) {
    stages = 3;
    qpc = whole;
  }
  p.n_stages = stages;
  p.quads_per_chunk = qpc;
  p.ok = avail > 0 && qpc >= 4;
  p.n_chunks = p.ok ? (n_quads + qpc - 1) / qpc : 0;
  p.stage_bytes = ((qpc * per_quad) + 127) & ~127;
  p.smem_bytes = stages * p.stage_bytes + fixed;
  p.ok = p.ok && p.smem_bytes <= max_smem_optin;
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
                                                   partials, plan.n_chunks, plan.quads_per_chunk, n_items,
                                                   plan.stage_bytes, plan.n_stages);
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

// Hot-path kernels of the batched IPP environment (sm_100a).
//
//   move_kernel          comm matrix + sequential masks / action choice / moves   (1 thread / env)
//   step_dense_kernel    per-cell fuse (local + global) + own measurement update + reward sums
//   reward_finalize      per-env reward from per-chunk partial sums (only when an env spans >1 chunk)
//   own_update_kernel    footprint-sparse measurement update (split observe/act mode)
//   reset_prep / reset_fill   episode reset (MT19937-compatible start positions + ground truth)
//
// Arithmetic specification: oracle/kernel_model.py (bit-exact for belief maps).
#include "ipp_cell.cuh"
#include "ipp_launch.h"

namespace ipp {

// =================================================================================================
// comm matrix + moves: agent/communication_log.py:39-58, agent/action_space.py:56-70,211-223,328-344
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

// One already-moved lower-id agent q against p: every rule is guarded by "more than one action
// still allowed", evaluated before the zeroing (action_space.py:328-344) => order dependent.
__device__ __forceinline__ uint32_t collide(const ipp_config& c, uint32_t m, const int32_t* p, const int32_t* q) {
  const int32_t dx = q[0] / c.spacing - p[0] / c.spacing;
  const int32_t dy = q[1] / c.spacing - p[1] / c.spacing;
  if (dx == 0 && dy == 0 && __popc(m) > 1) m &= ~((1u << 0) | (1u << 5));
  if (dx == -1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 1);
  if (dx == 0 && dy == -1 && __popc(m) > 1) m &= ~(1u << 2);
  if (dx == 0 && dy == 1 && __popc(m) > 1) m &= ~(1u << 3);
  if (dx == 1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 4);
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

__global__ void __launch_bounds__(128) move_kernel(const __grid_constant__ ipp_config cfg,
                                                   const uint32_t* __restrict__ episodes, const ipp_step_io io,
                                                   const int32_t t, const int32_t do_comm, const int32_t do_move) {
  const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cfg.n_envs) return;
  const int32_t A = cfg.n_agents;
  const uint32_t ep = episodes[b];
  int32_t pos[IPP_MAX_AGENTS][3];
  for (int32_t a = 0; a < A; ++a)
    for (int32_t d = 0; d < 3; ++d) pos[a][d] = io.pos_in[((int64_t)b * A + a) * 3 + d];

  if (do_comm && io.comm_out != nullptr) {
    for (int32_t i = 0; i < A; ++i) {
      const uint32_t key = stream_key(cfg.seed, ep, i, (uint32_t)t, PURPOSE_COMM);
      uint32_t row = 0;
      for (int32_t j = 0; j < A; ++j) {
        const int32_t dx = pos[i][0] - pos[j][0], dy = pos[i][1] - pos[j][1], dz = pos[i][2] - pos[j][2];
        const int32_t d2 = dx * dx + dy * dy + dz * dz;
        const uint32_t n24 = cell_hash(key, (uint32_t)j) >> 8;  // drawn for every ordered pair (:46)
        const bool ok = (d2 == 0) || (d2 <= cfg.comm_d2_max && n24 >= cfg.fail_thresh24);
        row |= (ok ? 1u : 0u) << j;
      }
      io.comm_out[(int64_t)b * A + i] = (uint8_t)row;
    }
  }
  if (!do_move) return;

  int32_t npos[IPP_MAX_AGENTS][3];
  uint32_t stuck = 0;
  for (int32_t a = 0; a < A; ++a) {
    uint32_t m = bounds_mask(cfg, pos[a]);
    for (int32_t j = 0; j < a; ++j) m = collide(cfg, m, pos[a], npos[j]);
    const int32_t cnt = __popc(m);
    int32_t act = -1;
    if (io.actions_in != nullptr) {
      // injected action (the reference's policy never emits a masked one): an action that would
      // leave the lattice is turned into "stay" and flagged, so positions always index the tables
      act = io.actions_in[(int64_t)b * A + a];
      if (act < -1 || act >= IPP_N_ACTIONS) act = -1;
      if (act >= 0 && !((bounds_mask(cfg, pos[a]) >> act) & 1u)) {
        act = -1;
        stuck |= 2u;
      }
    } else if (cnt > 0) {
      const uint32_t key = stream_key(cfg.seed, ep, a, (uint32_t)t, PURPOSE_ACTION);
      const float u = (float)(cell_hash(key, 0u) >> 8) * (1.0f / 16777216.0f);
      if (io.probs_in != nullptr) {
        // actor/network.py:63-66,90-96: probs * mask, then multinomial (train) or argmax (eval)
        float w[IPP_N_ACTIONS];
        float total = 0.0f;
        for (int32_t k = 0; k < IPP_N_ACTIONS; ++k) {
          const float pk = io.probs_in[((int64_t)b * A + a) * IPP_N_ACTIONS + k];
          w[k] = ((m >> k) & 1u) ? fmaxf(pk, 0.0f) : 0.0f;
          total += w[k];
        }
        if (!(total > 0.0f)) {
          act = kth_set_bit(m, min((int32_t)(u * (float)cnt), cnt - 1));
        } else if (io.greedy) {
          float best = -1.0f;
          for (int32_t k = 0; k < IPP_N_ACTIONS; ++k)
            if (((m >> k) & 1u) && w[k] > best) { best = w[k]; act = k; }
        } else {
          const float target = u * total;
          float acc = 0.0f;
          for (int32_t k = 0; k < IPP_N_ACTIONS; ++k) {
            if (w[k] > 0.0f) {
              act = k;  // last positive weight wins if rounding leaves target >= acc at the end
              acc += w[k];
              if (target < acc) break;
            }
          }
        }
      } else {
        act = kth_set_bit(m, min((int32_t)(u * (float)cnt), cnt - 1));
      }
    }
    if (cnt == 0) stuck |= 1u;  // the reference raises here; we stay in place and flag
    int32_t off[3] = {0, 0, 0};
    if (act == 0) off[2] = cfg.spacing;
    if (act == 1) off[0] = -cfg.spacing;
    if (act == 2) off[1] = -cfg.spacing;
    if (act == 3) off[1] = cfg.spacing;
    if (act == 4) off[0] = cfg.spacing;
    if (act == 5) off[2] = -cfg.spacing;
    for (int32_t d = 0; d < 3; ++d) {
      npos[a][d] = pos[a][d] + off[d];
      io.pos_out[((int64_t)b * A + a) * 3 + d] = npos[a][d];
    }
    if (io.actions_out != nullptr) io.actions_out[(int64_t)b * A + a] = act;
    if (io.mask_out != nullptr) io.mask_out[(int64_t)b * A + a] = (uint8_t)m;
  }
  if (io.stuck_out != nullptr) io.stuck_out[b] = (uint8_t)stuck;
}

// =================================================================================================
// dense per-cell pass, direct-load variant (per-quad arithmetic: ipp_cell.cuh)
// one block per (env, chunk); each thread owns quads tid, tid+256, ... of the chunk in all A+1 maps
// =================================================================================================
template <int A, bool DO_OWN>
__global__ void __launch_bounds__(STEP_THREADS)
    step_dense_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const int32_t* __restrict__ pos_in,
                      const int32_t* __restrict__ pos_out, const uint8_t* __restrict__ comm, const int32_t t,
                      float* __restrict__ reward_rel, float* __restrict__ reward_abs, double* __restrict__ partials,
                      const int32_t n_chunks, const int32_t quads_per_chunk) {
  const int32_t b = blockIdx.x / n_chunks;
  const int32_t chunk = blockIdx.x - b * n_chunks;
  const int32_t tid = threadIdx.x;
  __shared__ EnvMeta<A> s_meta;
  __shared__ double s_red[2][STEP_THREADS / 32];

  load_env_meta<A>(cfg, &s_meta, tid, b, st.episodes[b], pos_in, pos_out, comm, t, DO_OWN);
  __syncthreads();

  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (n_cells + 3) >> 2;
  const int64_t stride = cfg.map_stride;
  const uint8_t* gt_b = st.ground_truth + (int64_t)b * cfg.gt_stride;
  float* glob_b = st.global_map + (int64_t)b * stride;
  float* loc_b = st.local_maps + (int64_t)b * A * stride;

  double s1 = 0.0, s2 = 0.0;
  const int32_t q_end = min((chunk + 1) * quads_per_chunk, n_quads);
  for (int32_t q = chunk * quads_per_chunk + tid; q < q_end; q += STEP_THREADS) {
    const int32_t c0 = q << 2;
    QuadCtx<A> qc;
    make_quad_ctx<A>(cfg, s_meta, c0, *reinterpret_cast<const uint32_t*>(gt_b + c0), n_cells, qc);
    *reinterpret_cast<float4*>(glob_b + c0) =
        update_global_quad<A>(cfg, qc, *reinterpret_cast<const float4*>(glob_b + c0), s1, s2);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      float* lp = loc_b + (int64_t)i * stride + c0;
      *reinterpret_cast<float4*>(lp) =
          update_local_quad<A, DO_OWN>(cfg, s_meta, qc, i, *reinterpret_cast<const float4*>(lp));
    }
  }

  // ---- per-env reward: warp shuffle + shared-memory reduction of the two float64 sums ----
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = s1;
    s_red[1][tid >> 5] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int w = 0; w < STEP_THREADS / 32; ++w) {
      t1 += s_red[0][w];
      t2 += s_red[1][w];
    }
    if (n_chunks == 1) {
      write_rewards(reward_rel, reward_abs, b, t1, t2, n_cells);
    } else {
      partials[((int64_t)b * n_chunks + chunk) * 2 + 0] = t1;
      partials[((int64_t)b * n_chunks + chunk) * 2 + 1] = t2;
    }
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
  if (reward_rel != nullptr) reward_rel[b] = (float)(22.0 * (t1 / t2) - 0.5);
  if (reward_abs != nullptr) reward_abs[b] = (float)(10.0 * (t1 / (double)n_cells) - 0.17);
}

// =================================================================================================
// footprint-sparse own measurement update (ipp_act): mapping/mappings.py:32-61
// one block per (env, agent); threads walk the rect row by row
// =================================================================================================
__global__ void __launch_bounds__(256) own_update_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st,
                                                         const int32_t* __restrict__ pos_out, const int32_t t) {
  const int32_t A = cfg.n_agents;
  const int32_t b = blockIdx.x / A, a = blockIdx.x % A;
  __shared__ Meas s_m;
  if (threadIdx.x == 0)
    s_m = make_meas(cfg, pos_out + ((int64_t)b * A + a) * 3, st.episodes[b], (uint32_t)a, (uint32_t)t + 1u);
  __syncthreads();
  const Meas m = s_m;
  const int32_t w = m.yd - m.yu, h = m.xr - m.xl;
  if (w <= 0 || h <= 0) return;
  const int64_t stride = cfg.map_stride;
  float* lp = st.local_maps + ((int64_t)b * A + a) * stride;
  const uint8_t* gt = st.ground_truth + (int64_t)b * cfg.gt_stride;
  for (int32_t idx = threadIdx.x; idx < w * h; idx += blockDim.x) {
    const int32_t x = m.xl + idx / w, y = m.yu + idx % w;
    const int32_t cell = x * cfg.gy + y;
    const float pc = clamp_p(cfg, lp[cell]);
    float o = to_odds(pc);
    o = odds_pass(o, meas_k(m, (uint32_t)cell, gt[cell]), cfg.o_min, cfg.o_max);
    lp[cell] = from_odds(o);
  }
}

// =================================================================================================
// reset: mapping/ground_truths.py:42-56, agent/state_space.py:28-51, mapping/mappings.py:126-132,
// agent/agent.py:44-49.  numpy's legacy RandomState is MT19937 seeded by init_genrand; its
// randint() draws 32-bit words, masks them and rejects values above the range.
// =================================================================================================
constexpr int MT_DRAWS = 40;

struct MtStream {
  uint32_t lo[MT_DRAWS + 1];  // mt[0 .. MT_DRAWS]
  uint32_t hi[MT_DRAWS];      // mt[397 .. 397 + MT_DRAWS - 1]
  int32_t next;
  __device__ void seed(uint32_t s) {
    uint32_t v = s;
    lo[0] = v;
    for (int32_t i = 1; i < 397 + MT_DRAWS; ++i) {
      v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
      if (i <= MT_DRAWS) lo[i] = v;
      if (i >= 397) hi[i - 397] = v;
    }
    next = 0;
  }
  __device__ uint32_t draw() {
    const int32_t k = min(next, MT_DRAWS - 1);
    ++next;
    const uint32_t y = (lo[k] & 0x80000000u) | (lo[k + 1] & 0x7FFFFFFFu);
    uint32_t v = hi[k] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
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
__global__ void __launch_bounds__(128) reset_prep_kernel(const __grid_constant__ ipp_config cfg,
                                                         const uint32_t* __restrict__ episodes,
                                                         int32_t* __restrict__ pos_out,
                                                         int32_t* __restrict__ gt_params) {
  const int32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t A = cfg.n_agents;
  if (idx >= cfg.n_envs * (A + 1)) return;
  const int32_t b = idx / (A + 1), a = idx % (A + 1);
  const uint32_t ep = episodes[b];
  MtStream mt;
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

template <int A>
__global__ void __launch_bounds__(STEP_THREADS)
    reset_fill_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const int32_t* __restrict__ pos,
                      const int32_t* __restrict__ gt_params, const int32_t n_chunks, const int32_t quads_per_chunk) {
  const int32_t b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks, tid = threadIdx.x;
  __shared__ Meas s_m[A];
  const uint32_t ep = st.episodes[b];
  if (tid < A) s_m[tid] = make_meas(cfg, pos + ((int64_t)b * A + tid) * 3, ep, tid, 0u);
  __syncthreads();
  const int32_t split = gt_params[(int64_t)b * 4 + 0], lo = gt_params[(int64_t)b * 4 + 1],
                hi = gt_params[(int64_t)b * 4 + 2];
  const int32_t n_cells = cfg.gx * cfg.gy;
  const int32_t n_quads = (int32_t)(cfg.map_stride >> 2);
  const int64_t stride = cfg.map_stride;
  const float prior = cfg.prior;
  const float o_prior = to_odds(clamp_p(cfg, prior));
  const int32_t q_end = min((chunk + 1) * quads_per_chunk, n_quads);
  for (int32_t q = chunk * quads_per_chunk + tid; q < q_end; q += STEP_THREADS) {
    const int32_t c0 = q << 2;
    int32_t xs[4], ys[4];
    uint32_t g4 = 0;
    {
      int32_t x = c0 / cfg.gy, y = c0 - x * cfg.gy;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        xs[c] = x;
        ys[c] = y;
        const int32_t v = (split < 2) ? x : y;
        if (c0 + c < n_cells && v >= lo && v < hi) g4 |= 1u << (8 * c);
        if (++y == cfg.gy) { y = 0; ++x; }
      }
    }
    *reinterpret_cast<uint32_t*>(st.ground_truth + (int64_t)b * cfg.gt_stride + c0) = g4;
    *reinterpret_cast<float4*>(st.global_map + (int64_t)b * stride + c0) = make_float4(prior, prior, prior, prior);
#pragma unroll
    for (int i = 0; i < A; ++i) {
      const Meas m = s_m[i];
      float pv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pv[c] = prior;
        if (c0 + c < n_cells && in_rect(m, xs[c], ys[c]))
          pv[c] = from_odds(odds_pass(o_prior, meas_k(m, (uint32_t)(c0 + c), (g4 >> (8 * c)) & 0xFFu), cfg.o_min,
                                      cfg.o_max));
      }
      *reinterpret_cast<float4*>(st.local_maps + ((int64_t)b * A + i) * stride + c0) =
          make_float4(pv[0], pv[1], pv[2], pv[3]);
    }
  }
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

cudaError_t launch_move(const ipp_config& cfg, const uint32_t* episodes, const ipp_step_io& io, int32_t t, int do_comm,
                        int do_move, cudaStream_t s) {
  const int threads = 128;
  const int blocks = (cfg.n_envs + threads - 1) / threads;
  move_kernel<<<blocks, threads, 0, s>>>(cfg, episodes, io, t, do_comm, do_move);
  return cudaGetLastError();
}

cudaError_t launch_step_dense(const ipp_config& cfg, const ipp_state& st, const LaunchPlan& plan, const int32_t* pos_in,
                              const int32_t* pos_out, const uint8_t* comm, int32_t t, float* reward_rel,
                              float* reward_abs, double* partials, bool do_own, cudaStream_t s) {
  const dim3 grid((unsigned)plan.n_chunks * (unsigned)cfg.n_envs);
  if (do_own) {
    IPP_DISPATCH_A(cfg.n_agents, (step_dense_kernel<kA, true><<<grid, STEP_THREADS, 0, s>>>(
                                     cfg, st, pos_in, pos_out, comm, t, reward_rel, reward_abs, partials, plan.n_chunks,
                                     plan.quads_per_chunk)));
  } else {
    IPP_DISPATCH_A(cfg.n_agents, (step_dense_kernel<kA, false><<<grid, STEP_THREADS, 0, s>>>(
                                     cfg, st, pos_in, pos_out, comm, t, reward_rel, reward_abs, partials, plan.n_chunks,
                                     plan.quads_per_chunk)));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (plan.n_chunks > 1) e = launch_reward_finalize(cfg, partials, plan.n_chunks, reward_rel, reward_abs, s);
  return e;
}

cudaError_t launch_reward_finalize(const ipp_config& cfg, const double* partials, int32_t n_chunks, float* reward_rel,
                                   float* reward_abs, cudaStream_t s) {
  const int threads = 128;
  reward_finalize_kernel<<<(cfg.n_envs + threads - 1) / threads, threads, 0, s>>>(
      partials, cfg.n_envs, n_chunks, cfg.gx * cfg.gy, reward_rel, reward_abs);
  return cudaGetLastError();
}

cudaError_t launch_own_update(const ipp_config& cfg, const ipp_state& st, const int32_t* pos_out, int32_t t,
                              cudaStream_t s) {
  own_update_kernel<<<cfg.n_envs * cfg.n_agents, 256, 0, s>>>(cfg, st, pos_out, t);
  return cudaGetLastError();
}

cudaError_t launch_reset(const ipp_config& cfg, const ipp_state& st, const LaunchPlan& plan, int32_t* pos_out,
                         int32_t* gt_params, cudaStream_t s) {
  const int threads = 128;
  const int n = cfg.n_envs * (cfg.n_agents + 1);
  reset_prep_kernel<<<(n + threads - 1) / threads, threads, 0, s>>>(cfg, st.episodes, pos_out, gt_params);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const dim3 grid((unsigned)plan.n_chunks * (unsigned)cfg.n_envs);
  IPP_DISPATCH_A(cfg.n_agents, (reset_fill_kernel<kA><<<grid, STEP_THREADS, 0, s>>>(
                                   cfg, st, pos_out, gt_params, plan.n_chunks, plan.quads_per_chunk)));
  return cudaGetLastError();
}

}  // namespace ipp

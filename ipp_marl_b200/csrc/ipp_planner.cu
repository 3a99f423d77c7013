// Batched information-gain greedy planner and evaluation metrics, sm_100a.
//
//   ig_plan_kernel       IG_baseline.py:222-325 for every env of the batch: per agent the action mask against the
//                        CURRENT positions of the lower-id agents (:127-135), per allowed action the expected
//                        entropy reduction of the agent's local map over the candidate footprint (:222-285),
//                        per-agent normalisation (:287-295), the sequential discount of candidate cells other
//                        agents can also reach (:297-322) and the argmax (:324-325)
//   eval_metrics_kernel  masked entropy of the ground-truth-occupied cells and F1 of class 1 of the global map
//                        (IG_baseline.py:81-100,191-210; utils/utils.py:43-76; utils/state.py "eval" branch)
//
// Specification / checker: oracle/numpy_ig.py (pinned bit for bit to the reference by tests/golden/ig_*.npz).
#include "ipp_cell.cuh"
#include "ipp_launch.h"

namespace ipp {

namespace {

__device__ __forceinline__ uint32_t ig_bounds_mask(const ipp_config& c, const int32_t* p) {
  uint32_t m = 0x3Fu;
  if (p[2] == c.max_altitude) m &= ~1u;
  if (p[2] == c.min_altitude) m &= ~(1u << 5);
  if (p[1] == 0) m &= ~(1u << 2);
  if (p[1] == c.y_dim_m) m &= ~(1u << 3);
  if (p[0] == 0) m &= ~(1u << 1);
  if (p[0] == c.x_dim_m) m &= ~(1u << 4);
  return m;
}

// w of IG_baseline.py:247-254: the would-be posterior, snapped to 1 / 0 outside (0.499, 0.501)
__device__ __forceinline__ float snap_weight(float p) {
  const double d = (double)p;
  return d > 0.501 ? 1.0f : (d < 0.499 ? 0.0f : p);
}

// expected entropy reduction of one cell with stored odds o_in for a sensor with odds multipliers k_hi / k_lo
__device__ __forceinline__ float cell_gain(const ipp_config& cfg, float o_in, float k_hi, float k_lo) {
  const float o = fminf(fmaxf(o_in, cfg.o_min), cfg.o_max);
  const float sc = clamp_p(cfg, from_odds(o));
  const float p1 = from_odds(odds_pass(o, k_hi, cfg.o_min, cfg.o_max));  // measured "occupied"
  const float p2 = from_odds(odds_pass(o, k_lo, cfg.o_min, cfg.o_max));  // measured "free"
  const float h0 = shannon(cfg, sc);
  const float g1 = sc * (h0 - shannon(cfg, p1)) * snap_weight(p1);
  const float g2 = (1.0f - sc) * (h0 - shannon(cfg, p2)) * snap_weight(p2);
  return g1 + g2;
}

}  // namespace

// One block per env, one warp per agent.  Gains are summed in float64 (the reference sums a float64 array).
__global__ void __launch_bounds__(32 * IPP_MAX_AGENTS)
    ig_plan_kernel(const __grid_constant__ ipp_config cfg, const float* __restrict__ local_maps,
                   const int32_t* __restrict__ pos_in, const int32_t communication, int32_t* __restrict__ actions_out,
                   uint8_t* __restrict__ mask_out, double* __restrict__ gains_out, double* __restrict__ util_out) {
  const int32_t b = blockIdx.x;
  const int32_t A = cfg.n_agents;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int32_t s_pos[IPP_MAX_AGENTS][3];
  __shared__ int32_t s_cand[IPP_MAX_AGENTS][IPP_N_ACTIONS][3];
  __shared__ uint32_t s_mask[IPP_MAX_AGENTS];
  __shared__ double s_gain[IPP_MAX_AGENTS][IPP_N_ACTIONS];
  if (threadIdx.x < A * 3) s_pos[threadIdx.x / 3][threadIdx.x % 3] = pos_in[(int64_t)b * A * 3 + threadIdx.x];
  __syncthreads();

  if (warp < A) {
    const int a = warp;
    // action mask: bounds, then the collision rules against the lower-id agents where they stand NOW
    uint32_t m = ig_bounds_mask(cfg, s_pos[a]);
    const int32_t ixa = s_pos[a][0] / cfg.spacing, iya = s_pos[a][1] / cfg.spacing;
    for (int j = 0; j < a; ++j) {
      const int32_t dx = s_pos[j][0] / cfg.spacing - ixa, dy = s_pos[j][1] / cfg.spacing - iya;
      if (dx == 0 && dy == 0 && __popc(m) > 1) m &= ~((1u << 0) | (1u << 5));
      if (dx == -1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 1);
      if (dx == 0 && dy == -1 && __popc(m) > 1) m &= ~(1u << 2);
      if (dx == 0 && dy == 1 && __popc(m) > 1) m &= ~(1u << 3);
      if (dx == 1 && dy == 0 && __popc(m) > 1) m &= ~(1u << 4);
    }
    const float* map = local_maps + ((int64_t)b * A + a) * cfg.map_stride;
    for (int k = 0; k < IPP_N_ACTIONS; ++k) {
      double sum = 0.0;
      int32_t np[3] = {-1, -1, -1};
      if ((m >> k) & 1u) {
        const int32_t ox = (k == 4) - (k == 1), oy = (k == 3) - (k == 2), oz = (k == 0) - (k == 5);
        np[0] = s_pos[a][0] + ox * cfg.spacing;
        np[1] = s_pos[a][1] + oy * cfg.spacing;
        np[2] = s_pos[a][2] + oz * cfg.spacing;
        const Meas fp = make_meas(cfg, np, 0u, 0u, 0u);  // only the clipped footprint and the multipliers are used
        const int32_t h = fp.xr - fp.xl, w = fp.yd - fp.yu;
        if (h > 0 && w > 0) {
          for (int32_t idx = lane; idx < h * w; idx += 32) {
            const int32_t r = idx / w, c = idx - r * w;
            sum += (double)cell_gain(cfg, map[(fp.xl + r) * cfg.gy + fp.yu + c], fp.k_hi, fp.k_lo);
          }
        }
        sum = warp_sum(sum);
      }
      if (lane == 0) {
        s_gain[a][k] = sum / 1000.0;
        for (int d = 0; d < 3; ++d) s_cand[a][k][d] = np[d];
      }
    }
    if (lane == 0) s_mask[a] = m;
  }
  __syncthreads();

  if (threadIdx.x != 0) return;
  // relative gains, sequential discount, argmax: tiny and order dependent -> one thread per env
  double rel[IPP_MAX_AGENTS][IPP_N_ACTIONS];
  for (int a = 0; a < A; ++a) {
    double total = 0.0;
    for (int k = 0; k < IPP_N_ACTIONS; ++k) total += s_gain[a][k];
    for (int k = 0; k < IPP_N_ACTIONS; ++k) rel[a][k] = s_gain[a][k] / total;
  }
  if (communication) {
    for (int a = 0; a < A; ++a)
      for (int k1 = 0; k1 < IPP_N_ACTIONS; ++k1) {
        if (!((s_mask[a] >> k1) & 1u)) continue;
        const double r1 = rel[a][k1];
        for (int bb = 0; bb < A; ++bb) {
          if (bb == a) continue;
          for (int k2 = 0; k2 < IPP_N_ACTIONS; ++k2) {
            if (!((s_mask[bb] >> k2) & 1u)) continue;
            if (s_cand[a][k1][0] == s_cand[bb][k2][0] && s_cand[a][k1][1] == s_cand[bb][k2][1] &&
                s_cand[a][k1][2] == s_cand[bb][k2][2])
              rel[a][k1] = r1 * (1.0 - rel[bb][k2]);  // the last match wins (IG_baseline.py:309-320)
          }
        }
      }
  }
  for (int a = 0; a < A; ++a) {
    int best = 0;  // np.argmax: first maximum, a NaN counts as the maximum
    if (!isnan(rel[a][0]))
      for (int k = 1; k < IPP_N_ACTIONS; ++k) {
        if (isnan(rel[a][k])) { best = k; break; }
        if (rel[a][k] > rel[a][best]) best = k;
      }
    actions_out[(int64_t)b * A + a] = best;
    if (mask_out != nullptr) mask_out[(int64_t)b * A + a] = (uint8_t)s_mask[a];
    for (int k = 0; k < IPP_N_ACTIONS; ++k) {
      if (gains_out != nullptr) gains_out[((int64_t)b * A + a) * IPP_N_ACTIONS + k] = s_gain[a][k];
      if (util_out != nullptr) util_out[((int64_t)b * A + a) * IPP_N_ACTIONS + k] = rel[a][k];
    }
  }
}

// One block per env.  entropy = sum over ground-truth-occupied cells of H(clamp(p)) / their count;
// F1 of class 1 with the prediction p > 0.5.
__global__ void __launch_bounds__(256)
    eval_metrics_kernel(const __grid_constant__ ipp_config cfg, const float* __restrict__ global_map,
                        const uint8_t* __restrict__ gt, double* __restrict__ entropy_out,
                        double* __restrict__ f1_out) {
  const int32_t b = blockIdx.x;
  const int32_t n_cells = cfg.gx * cfg.gy;
  const float* g = global_map + (int64_t)b * cfg.map_stride;
  const uint8_t* t = gt + (int64_t)b * cfg.gt_stride;
  double h = 0.0;
  int32_t ones = 0, tp = 0, fp = 0;
  for (int32_t c = threadIdx.x; c < n_cells; c += blockDim.x) {
    const float p = from_odds(g[c]);  // the map holds odds
    const bool one = t[c] != 0, pred = p > 0.5f;
    if (one) h += (double)shannon(cfg, p);
    ones += one;
    tp += pred && one;
    fp += pred && !one;
  }
  __shared__ double s_h[8];
  __shared__ int32_t s_i[3][8];
  h = warp_sum(h);
  for (int off = 16; off > 0; off >>= 1) {
    ones += __shfl_xor_sync(0xFFFFFFFFu, ones, off);
    tp += __shfl_xor_sync(0xFFFFFFFFu, tp, off);
    fp += __shfl_xor_sync(0xFFFFFFFFu, fp, off);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_h[warp] = h;
    s_i[0][warp] = ones;
    s_i[1][warp] = tp;
    s_i[2][warp] = fp;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double ht = 0.0;
  int32_t n1 = 0, ntp = 0, nfp = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    ht += s_h[w];
    n1 += s_i[0][w];
    ntp += s_i[1][w];
    nfp += s_i[2][w];
  }
  // np.unique(...)[-1]: the count of the largest value present (all cells when the field is empty)
  entropy_out[b] = ht / (double)(n1 > 0 ? n1 : n_cells);
  const int32_t nfn = n1 - ntp;
  const int32_t den = 2 * ntp + nfp + nfn;
  f1_out[b] = den > 0 ? 2.0 * (double)ntp / (double)den : 0.0;
}

cudaError_t launch_ig_plan(const ipp_config& cfg, const ipp_state& st, const int32_t* pos_in, int communication,
                           int32_t* actions_out, uint8_t* mask_out, double* gains_out, double* util_out,
                           cudaStream_t s) {
  ig_plan_kernel<<<cfg.n_envs, 32 * cfg.n_agents, 0, s>>>(cfg, st.local_maps, pos_in, communication, actions_out,
                                                          mask_out, gains_out, util_out);
  return cudaGetLastError();
}

cudaError_t launch_eval_metrics(const ipp_config& cfg, const ipp_state& st, double* entropy_out, double* f1_out,
                                cudaStream_t s) {
  eval_metrics_kernel<<<cfg.n_envs, 256, 0, s>>>(cfg, st.global_map, st.ground_truth, entropy_out, f1_out);
  return cudaGetLastError();
}

}  // namespace ipp

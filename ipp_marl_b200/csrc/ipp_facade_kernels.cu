// Single-map kernels behind the drop-in facade entry points (ipp_update_cells, ipp_fuse_map,
// ipp_shannon_entropy, ipp_utility_reward).  Same odds-space arithmetic as the batched kernels, with
// the multiplier derived from an arbitrary measurement value y: k = odds(y) / odds(prior)
// (= exp(logit y - logit prior), mapping/mappings.py:112-117).
#include "ipp_device.cuh"
#include "ipp_launch.h"

namespace ipp {

__device__ __forceinline__ float k_of_y(float y, float o_prior) { return __fdiv_rn(to_odds(y), o_prior); }

// mode 0: out = update(x, y) with x clamped in place (Mapping.apply_update)
__global__ void update_cells_kernel(const __grid_constant__ ipp_config cfg, float* __restrict__ x,
                                    const float* __restrict__ y, const int y_is_scalar, const float y_scalar,
                                    const int64_t n, float* __restrict__ out) {
  const float o_prior = to_odds(cfg.prior);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pc = clamp_p(cfg, x[i]);
    x[i] = pc;
    const float yy = y_is_scalar ? y_scalar : y[i];
    out[i] = from_odds(odds_pass(to_odds(pc), k_of_y(yy, o_prior), cfg.o_min, cfg.o_max));
  }
}

__global__ void entropy_kernel(const __grid_constant__ ipp_config cfg, float* __restrict__ p, const int64_t n,
                               float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pc = clamp_p(cfg, p[i]);
    p[i] = pc;  // utils/state.py:119-120 clamps its argument in place
    out[i] = shannon(cfg, pc);
  }
}

__global__ void __launch_bounds__(256) utility_partial_kernel(const __grid_constant__ ipp_config cfg,
                                                              const float* __restrict__ last,
                                                              const float* __restrict__ next, const int64_t n,
                                                              double* __restrict__ partial) {
  __shared__ double s_red[2][8];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float hl = shannon(cfg, last[i]);
    const float hn = shannon(cfg, next[i]);
    const float w = weight_of(next[i]);
    s1 += (double)(w * (hl - hn));
    s2 += (double)(w * hl);
  }
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_down_sync(0xFFFFFFFFu, s1, off);
    s2 += __shfl_down_sync(0xFFFFFFFFu, s2, off);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = s1;
    s_red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t1 += s_red[0][w];
      t2 += s_red[1][w];
    }
    partial[2 * blockIdx.x + 0] = t1;
    partial[2 * blockIdx.x + 1] = t2;
  }
}

__global__ void utility_final_kernel(const double* __restrict__ partial, const int n_blocks, const int64_t n,
                                     double* __restrict__ out2) {
  double t1 = 0.0, t2 = 0.0;
  for (int i = 0; i < n_blocks; ++i) {
    t1 += partial[2 * i];
    t2 += partial[2 * i + 1];
  }
  const double absolute = t1 / (double)n;  // utils/reward.py:81
  out2[0] = absolute;
  out2[1] = absolute / (t2 / (double)n);  // utils/reward.py:82
}

// Simulation.get_measurement (mapping/simulations.py:42-65): noisy view of the ground truth inside a
// clipped footprint [yu, yd, xl, xr]; out is the [xr-xl, yd-yu] float32 block the reference returns.
__global__ void measure_kernel(const __grid_constant__ ipp_config cfg, const uint8_t* __restrict__ gt,
                               const int32_t yu, const int32_t yd, const int32_t xl, const int32_t xr,
                               const uint32_t key, const uint32_t thresh, const float y_hi, const float y_lo,
                               float* __restrict__ out) {
  const int32_t w = yd - yu, n = (xr - xl) * w;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t x = xl + i / w, y = yu + i % w;
    const uint32_t cell = (uint32_t)(x * cfg.gy + y);
    const bool wrong = noise_word(key, cell) < thresh;
    const bool seen_one = (gt[cell] != 0) != wrong;
    out[i] = seen_one ? y_hi : y_lo;
  }
}

cudaError_t launch_measure(const ipp_config& cfg, const uint8_t* gt, const int32_t* rect, uint32_t key,
                           uint32_t thresh, float y_hi, float y_lo, float* out, cudaStream_t s) {
  const int64_t n = (int64_t)(rect[3] - rect[2]) * (rect[1] - rect[0]);
  if (n <= 0) return cudaSuccess;
  measure_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(cfg, gt, rect[0], rect[1], rect[2], rect[3], key, thresh, y_hi,
                                                        y_lo, out);
  return cudaGetLastError();
}

static int grid_for(int64_t n, int threads, int cap) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

cudaError_t launch_update_cells(const ipp_config& cfg, float* x, const float* y, int y_is_scalar, float y_scalar,
                                int64_t n, float* out, cudaStream_t s) {
  update_cells_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(cfg, x, y, y_is_scalar, y_scalar, n, out);
  return cudaGetLastError();
}

cudaError_t launch_entropy(const ipp_config& cfg, float* p, int64_t n, float* out, cudaStream_t s) {
  entropy_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(cfg, p, n, out);
  return cudaGetLastError();
}

// out2 must have room for 2 + 2*UTILITY_BLOCKS doubles (result first, partials after)
cudaError_t launch_utility_reward(const ipp_config& cfg, const float* last, const float* next, int64_t n,
                                  double* out2, cudaStream_t s) {
  const int blocks = grid_for(n, 256, 296);
  utility_partial_kernel<<<blocks, 256, 0, s>>>(cfg, last, next, n, out2 + 2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  utility_final_kernel<<<1, 1, 0, s>>>(out2 + 2, blocks, n, out2);
  return cudaGetLastError();
}

}  // namespace ipp

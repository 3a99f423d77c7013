// Single-map kernels behind the drop-in facade entry points (ipp_update_cells, ipp_fuse_map,
// ipp_shannon_entropy, ipp_utility_reward, ipp_measure).
#include "ipp_device.cuh"
#include "ipp_launch.h"

namespace ipp {

// ------------------------------------------------------------------------------------------------------------------
// The single-map entry points reproduce the reference's DTYPE FLOW under numpy >= 2 (SURVEY.md section 7 "dtype
// drift"), not only its formula: mapping/mappings.py:109-124 clamps x in x's own dtype, takes logit(x) in x's dtype
// and logit(y) in y's dtype (float32 for measurement arrays, float64 for the Python float IG_baseline.py:240-245
// passes), adds them in the promoted dtype, and — because l_p = np.log(prior / (1 - prior)) is an np.float64 scalar —
// evaluates `1 - 1/(1 + exp(l))` in float64 and RETURNS float64.  Mapping.fuse_map therefore returns float64 whenever
// it fused at least one peer, every pass after the first runs entirely in float64, and Agent.local_map stays float64
// until update_grid_map writes into a float32 map again.  The batched kernels keep one float32 odds state instead
// (DESIGN.md section 3); these kernels exist so that the UNCHANGED reference callers see the same numbers through the
// facade as through the reference's own modules.  float32 logs are computed in float64 and rounded (correctly rounded
// float32 log; numpy's SIMD logf may differ from it in the last bit).
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Flow;
template <>
struct Flow<float> {
  static __device__ __forceinline__ float clamp(float x) {  // x[0.9999 < x] = 0.9999: the Python float is cast to float32
    return fminf(fmaxf(x, 0.0001f), 0.9999f);
  }
  static __device__ __forceinline__ float logit(float x) {
    return (float)log((double)__fdiv_rn(x, __fsub_rn(1.0f, x)));
  }
  static __device__ __forceinline__ float entropy(float p) {  // utils/state.py:121 in float32
    const float q = __fsub_rn(1.0f, p);
    const float lp = (float)log2((double)p), lq = (float)log2((double)q);
    return __fsub_rn(__fmul_rn(-p, lp), __fmul_rn(q, lq));
  }
  static __device__ __forceinline__ bool gt(float v, double c) { return v > (float)c; }
  static __device__ __forceinline__ bool lt(float v, double c) { return v < (float)c; }
};
template <>
struct Flow<double> {
  static __device__ __forceinline__ double clamp(double x) { return fmin(fmax(x, 0.0001), 0.9999); }
  static __device__ __forceinline__ double logit(double x) { return log(x / (1.0 - x)); }
  static __device__ __forceinline__ double entropy(double p) { return -p * log2(p) - (1.0 - p) * log2(1.0 - p); }
  static __device__ __forceinline__ bool gt(double v, double c) { return v > c; }
  static __device__ __forceinline__ bool lt(double v, double c) { return v < c; }
};

// logit(y) of a measurement value: float32 measurements come from a handful of values (y_hi / y_lo per altitude, 0.5)
// whose float32 logits the host evaluated with numpy itself (ipp_config.meas_y / meas_ly); anything else is computed
struct MeasTable {
  int n;
  float y[16], ly[16];
};
__device__ __forceinline__ float logit_y(const MeasTable& mt, float y) {
  for (int i = 0; i < mt.n; ++i)
    if (mt.y[i] == y) return mt.ly[i];
  return Flow<float>::logit(y);
}
__device__ __forceinline__ double logit_y(const MeasTable&, double y) { return Flow<double>::logit(y); }

template <typename XT, typename YT>
__device__ __forceinline__ double flow_update(XT& x, const YT y, const double l_prior, const MeasTable& mt) {
  x = Flow<XT>::clamp(x);
  const XT lx = Flow<XT>::logit(x);
  const YT ly = logit_y(mt, y);
  double lxy;
  if (sizeof(XT) == 4 && sizeof(YT) == 4) lxy = (double)__fadd_rn((float)lx, (float)ly);  // float32 + float32
  else lxy = (double)lx + (double)ly;
  const double l = lxy - l_prior;
  return 1.0 - 1.0 / (1.0 + exp(l));  // mappings.py:121-124 literally, in float64
}

// Mapping.apply_update: x clamped in place (in its dtype), out float64.  y: array or one scalar of type YT.
template <typename XT, typename YT>
__global__ void update_cells_kernel(XT* __restrict__ x, const YT* __restrict__ y, const int y_is_scalar,
                                    const double l_prior, const __grid_constant__ MeasTable mt, const int64_t n,
                                    double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    XT xv = x[i];
    out[i] = flow_update<XT, YT>(xv, y_is_scalar ? y[0] : y[i], l_prior, mt);
    x[i] = xv;
  }
}

// Mapping.fuse_map: own (already cast to float32 by `np.float32(own.copy())`, mappings.py:83,93,100) fused with
// n_others float32 maps in order; the first pass reads float32 and returns float64, the others run in float64.
__global__ void fuse_map_kernel(const float* __restrict__ own, const float* __restrict__ others, const int n_others,
                                const double l_prior, const __grid_constant__ MeasTable mt, const int64_t n,
                                double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float x0 = own[i];
    double v = (double)x0;
    for (int k = 0; k < n_others; ++k) {
      const float y = others[(int64_t)k * n + i];
      if (k == 0) v = flow_update<float, float>(x0, y, l_prior, mt);
      else v = flow_update<double, float>(v, y, l_prior, mt);
    }
    out[i] = v;
  }
}

// get_shannon_entropy (utils/state.py:118-121): clamps IN PLACE, H in the dtype of p
template <typename T>
__global__ void entropy_kernel(T* __restrict__ p, const int64_t n, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T pc = Flow<T>::clamp(p[i]);
    p[i] = pc;
    out[i] = Flow<T>::entropy(pc);
  }
}

// get_utility_reward (utils/reward.py:68-82) with the "reward" branch of get_w_entropy_map (utils/state.py:14-76):
// both maps are COPIED, clamped and their entropies taken in their own dtypes; weights from the next map
// (> 0.501 -> 1, < 0.499 -> 0, else 0.5); products and means in the promoted dtype (float64 unless both are float32).
template <typename LT, typename NT>
__global__ void __launch_bounds__(256) utility_partial_kernel(const LT* __restrict__ last, const NT* __restrict__ next,
                                                              const int64_t n, double* __restrict__ partial) {
  __shared__ double s_red[2][8];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const NT nv = next[i];
    const double w = Flow<NT>::gt(nv, 0.501) ? 1.0 : (Flow<NT>::lt(nv, 0.499) ? 0.0 : 0.5);
    const LT hl = Flow<LT>::entropy(Flow<LT>::clamp(last[i]));
    const NT hn = Flow<NT>::entropy(Flow<NT>::clamp(nv));
    if (sizeof(LT) == 4 && sizeof(NT) == 4) {
      s1 += (double)__fmul_rn((float)w, __fsub_rn((float)hl, (float)hn));
      s2 += (double)__fmul_rn((float)w, (float)hl);
    } else {
      s1 += w * ((double)hl - (double)hn);
      s2 += w * (double)hl;
    }
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

static MeasTable meas_table(const ipp_config& cfg) {
  MeasTable mt;
  mt.n = cfg.n_meas < 0 ? 0 : (cfg.n_meas > 16 ? 16 : cfg.n_meas);
  for (int i = 0; i < 16; ++i) {
    mt.y[i] = cfg.meas_y[i];
    mt.ly[i] = cfg.meas_ly[i];
  }
  return mt;
}

cudaError_t launch_update_cells(const ipp_config& cfg, void* x, int x_f64, const void* y, int y_f64, int y_is_scalar,
                                int64_t n, double* out, cudaStream_t s) {
  const int g = grid_for(n, 256, 148 * 8);
  const double l_prior = cfg.l_prior;
  const MeasTable mt = meas_table(cfg);
  if (x_f64 && y_f64)
    update_cells_kernel<double, double><<<g, 256, 0, s>>>((double*)x, (const double*)y, y_is_scalar, l_prior, mt, n, out);
  else if (x_f64)
    update_cells_kernel<double, float><<<g, 256, 0, s>>>((double*)x, (const float*)y, y_is_scalar, l_prior, mt, n, out);
  else if (y_f64)
    update_cells_kernel<float, double><<<g, 256, 0, s>>>((float*)x, (const double*)y, y_is_scalar, l_prior, mt, n, out);
  else
    update_cells_kernel<float, float><<<g, 256, 0, s>>>((float*)x, (const float*)y, y_is_scalar, l_prior, mt, n, out);
  return cudaGetLastError();
}

cudaError_t launch_fuse_map(const ipp_config& cfg, const float* own, const float* others, int n_others, int64_t n,
                            double* out, cudaStream_t s) {
  fuse_map_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(own, others, n_others, cfg.l_prior, meas_table(cfg), n,
                                                            out);
  return cudaGetLastError();
}

cudaError_t launch_entropy(void* p, int is_f64, int64_t n, void* out, cudaStream_t s) {
  const int g = grid_for(n, 256, 148 * 8);
  if (is_f64) entropy_kernel<double><<<g, 256, 0, s>>>((double*)p, n, (double*)out);
  else entropy_kernel<float><<<g, 256, 0, s>>>((float*)p, n, (float*)out);
  return cudaGetLastError();
}

// out2 must have room for 2 + 2*296 doubles (result first, partials after)
cudaError_t launch_utility_reward(const void* last, int last_f64, const void* next, int next_f64, int64_t n,
                                  double* out2, cudaStream_t s) {
  const int blocks = grid_for(n, 256, 296);
  double* part = out2 + 2;
  if (last_f64 && next_f64)
    utility_partial_kernel<double, double><<<blocks, 256, 0, s>>>((const double*)last, (const double*)next, n, part);
  else if (last_f64)
    utility_partial_kernel<double, float><<<blocks, 256, 0, s>>>((const double*)last, (const float*)next, n, part);
  else if (next_f64)
    utility_partial_kernel<float, double><<<blocks, 256, 0, s>>>((const float*)last, (const double*)next, n, part);
  else
    utility_partial_kernel<float, float><<<blocks, 256, 0, s>>>((const float*)last, (const float*)next, n, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  utility_final_kernel<<<1, 1, 0, s>>>(part, blocks, n, out2);
  return cudaGetLastError();
}

}  // namespace ipp

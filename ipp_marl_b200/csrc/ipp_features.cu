// Network-input feature builders: the actor observation
// [P, P, 7] (actor/transformations.py:14-176) and the critic state [P, P, 12]
// (critic/transformations.py:17-132), batched over envs x agents.
//
// The lattice-sized maps are area-pooled exactly like cv2.resize(INTER_AREA) does it
// (utils/state.py:22-41): separable tap tables built on the host by the same rule as OpenCV's
// computeResizeAreaTab (fractional-overlap box filter when shrinking; the INTER_AREA flavour of
// bilinear when the source is smaller than the lattice, e.g. the 10x10 footprint image at 5 m).
// Everything about the latest measurements comes from the code bytes (ipp_cell.cuh) and the positions.
#include <algorithm>
#include <cmath>
#include <vector>

#include "ipp_cell.cuh"
#include "ipp_launch.h"
#include "ipp_ptx.cuh"

namespace ipp {

// ------------------------------------------------------------------------------------------------
// host: tap tables
// ------------------------------------------------------------------------------------------------
static void area_taps(int ssize, int dsize, std::vector<std::vector<std::pair<int, float>>>& out) {
  out.assign(dsize, {});
  const double scale = (double)ssize / dsize;
  if (scale >= 1.0) {  // OpenCV computeResizeAreaTab
    for (int dx = 0; dx < dsize; ++dx) {
      const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
      const double cell = std::min(scale, ssize - fsx1);
      int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
      sx2 = std::min(sx2, ssize - 1);
      sx1 = std::min(sx1, sx2);
      if (sx1 - fsx1 > 1e-3) out[dx].push_back({sx1 - 1, (float)((sx1 - fsx1) / cell)});
      for (int sx = sx1; sx < sx2; ++sx) out[dx].push_back({sx, (float)(1.0 / cell)});
      if (fsx2 - sx2 > 1e-3) out[dx].push_back({sx2, (float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)});
    }
  } else {  // source smaller than the lattice: OpenCV's linear path with the INTER_AREA coefficient rule
    const double inv = (double)dsize / ssize;
    for (int dx = 0; dx < dsize; ++dx) {
      int sx = (int)std::floor(dx * scale);
      float fx = (float)((dx + 1) - (sx + 1) * inv);
      fx = fx <= 0 ? 0.f : fx - std::floor(fx);
      if (sx < 0) { fx = 0; sx = 0; }
      if (sx >= ssize - 1) { fx = 0; sx = ssize - 1; }
      out[dx].push_back({sx, 1.f - fx});
      if (sx + 1 < ssize) out[dx].push_back({sx + 1, fx});
    }
  }
}

// table ids: 0 map-x (gx -> px), 1 map-y (gy -> py), 2+2a axis-0 of the footprint image at altitude a
// (size 2*ry -> px), 3+2a axis-1 (size 2*rx -> py)   [mapping/mappings.py:41-43: the image is (yd-yu) x (xr-xl)]
cudaError_t build_pool_tables(const ipp_config& cfg, PoolTables* pt) {
  const int n_tabs = 2 + 2 * cfg.n_alt;
  std::vector<std::vector<std::vector<std::pair<int, float>>>> taps(n_tabs);
  area_taps(cfg.gx, cfg.px, taps[0]);
  area_taps(cfg.gy, cfg.py, taps[1]);
  for (int a = 0; a < cfg.n_alt; ++a) {
    area_taps(std::max(2 * cfg.radius_y[a], 1), cfg.px, taps[2 + 2 * a]);
    area_taps(std::max(2 * cfg.radius_x[a], 1), cfg.py, taps[3 + 2 * a]);
  }
  int maxt = 1;
  for (auto& tab : taps)
    for (auto& row : tab) maxt = std::max<int>(maxt, (int)row.size());
  const int pmax = IPP_MAX_LATTICE;
  std::vector<int32_t> ints((size_t)n_tabs * pmax * 2, 0);
  std::vector<float> w((size_t)n_tabs * pmax * maxt, 0.f);
  for (int k = 0; k < n_tabs; ++k)
    for (size_t d = 0; d < taps[k].size(); ++d) {
      const auto& row = taps[k][d];
      ints[((size_t)k * pmax + d) * 2 + 0] = row.empty() ? 0 : row[0].first;
      ints[((size_t)k * pmax + d) * 2 + 1] = (int32_t)row.size();
      for (size_t i = 0; i < row.size(); ++i) {
        // taps are consecutive source indices (both rules produce runs)
        w[((size_t)k * pmax + d) * maxt + i] = row[i].second;
      }
    }
  pt->maxt = maxt;
  pt->n_tabs = n_tabs;
  cudaError_t e = cudaMalloc(&pt->ints, ints.size() * sizeof(int32_t));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&pt->w, w.size() * sizeof(float));
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(pt->ints, ints.data(), ints.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(pt->w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
}

void free_pool_tables(PoolTables* pt) {
  if (pt->ints) cudaFree(pt->ints);
  if (pt->w) cudaFree(pt->w);
  pt->ints = nullptr;
  pt->w = nullptr;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
struct Tab {
  const int32_t* ints;  // [pmax][2] start, count
  const float* w;       // [pmax][maxt]
  int maxt;
  __device__ __forceinline__ int start(int d) const { return ints[2 * d]; }
  __device__ __forceinline__ int count(int d) const { return ints[2 * d + 1]; }
  __device__ __forceinline__ float weight(int d, int i) const { return w[d * maxt + i]; }
};

__device__ __forceinline__ Tab get_tab(const PoolTables& pt, int k) {
  return Tab{pt.ints + (size_t)k * IPP_MAX_LATTICE * 2, pt.w + (size_t)k * IPP_MAX_LATTICE * pt.maxt, pt.maxt};
}

// utils/state.py:67-76 + 118-121 on a pooled value: (w * H(clamp(v)), clamp(v))
__device__ __forceinline__ float2 w_entropy(const ipp_config& cfg, float v) {
  const float w = v > 0.501f ? 1.0f : (v < 0.499f ? 0.0f : 0.5f);
  const float c = fminf(fmaxf(v, cfg.p_min), cfg.p_max);
  const float h = -c * __log2f(c) - (1.0f - c) * __log2f(1.0f - c);
  return make_float2(w * h, c);
}

struct AgentGeo {
  int32_t ix, iy, iz;      // lattice indices (iz = altitude level, 0-based: index of the host tables)
  int32_t zi;              // agent/state_space.py:56: position[2] // spacing - 1 (what the features use)
  int32_t xl, xr, yu, yd;  // clipped footprint (half-open)
  int32_t rxl, ryu;        // raw footprint origin
  int32_t rx2, ry2;        // raw footprint extents 2*rx, 2*ry
};

__device__ __forceinline__ AgentGeo agent_geo(const ipp_config& c, const int32_t* pos) {
  AgentGeo g;
  g.ix = clampi(pos[0] / c.spacing, 0, c.px - 1);
  g.iy = clampi(pos[1] / c.spacing, 0, c.py - 1);
  g.iz = clampi(pos[2] / c.spacing - c.min_altitude / c.spacing, 0, c.n_alt - 1);
  g.zi = pos[2] / c.spacing - 1;
  const int32_t cx = c.cell_x[g.ix], cy = c.cell_y[g.iy], rx = c.radius_x[g.iz], ry = c.radius_y[g.iz];
  g.rxl = cx - rx;
  g.ryu = cy - ry;
  g.rx2 = 2 * rx;
  g.ry2 = 2 * ry;
  g.xl = clampi(cx - rx, 0, c.gx - 1);
  g.xr = clampi(cx + rx, 0, c.gx - 1);
  g.yu = clampi(cy - ry, 0, c.gy - 1);
  g.yd = clampi(cy + ry, 0, c.gy - 1);
  return g;
}

// ------------------------------------------------------------------------------------------------
// actor observation: one block per (env, agent)
// ------------------------------------------------------------------------------------------------
template <int A>
__global__ void __launch_bounds__(128)
    features_actor_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const PoolTables pt,
                          const int32_t* __restrict__ pos_in, const uint8_t* __restrict__ comm, const int32_t t,
                          float* __restrict__ obs_out) {
  const int32_t b = blockIdx.x / A, i = blockIdx.x - b * A;
  __shared__ AgentGeo s_geo[A];
  if (threadIdx.x < A) s_geo[threadIdx.x] = agent_geo(cfg, pos_in + ((int64_t)b * A + threadIdx.x) * 3);
  __syncthreads();
  const uint32_t received = comm[(int64_t)b * A + i];  // includes the agent itself (communication_log.py:47)
  const AgentGeo me = s_geo[i];
  const float* local = st.local_maps + ((int64_t)b * A + i) * cfg.map_stride;
  const uint8_t* codes = st.meas_codes + code_row_offset(cfg, t, b);
  const Tab tx = get_tab(pt, 0), ty = get_tab(pt, 1);
  const Tab fx = get_tab(pt, 2 + 2 * me.iz), fy = get_tab(pt, 3 + 2 * me.iz);
  // where the clipped measurement block sits inside the raw footprint image (utils/utils.py:79-98);
  // the image is (yd-yu) x (xr-xl) but is indexed [x-block, y-block] (mapping/mappings.py:41-43,72-76)
  const int32_t h = me.ry2, w = me.rx2;
  int32_t b_yu = 0, b_yd = h, b_xl = 0, b_xr = w;
  if (me.yu > me.ryu) b_yu = h - (me.yd - me.yu);
  if (me.yd < me.ryu + me.ry2) b_yd = me.yd - me.yu;
  if (me.xr < me.rxl + me.rx2) b_xr = me.xr - me.xl;
  if (me.xl > me.rxl) b_xl = w - (me.xr - me.xl);
  const float y_hi = cfg.y_hi[me.iz], y_lo = cfg.y_lo[me.iz];
  const float budget = (float)(cfg.budget - t) / (float)cfg.budget;   // transformations.py:91-93
  const float agent_id = (float)(i + 1) / (float)A;                   // :86-88
  const float own_alt = (float)(me.zi + 1) / (float)(cfg.n_alt + 1);  // :125-131

  for (int32_t c = threadIdx.x; c < cfg.px * cfg.py; c += blockDim.x) {
    const int32_t li = c / cfg.py, lj = c - li * cfg.py;
    // ---- pooled fused local map and pooled footprint-ownership map (taps over the G x G grid) ----
    float pl = 0.0f, pf_own = 0.0f;
    {
      const int32_t x0 = tx.start(li), nx = tx.count(li), y0 = ty.start(lj), ny = ty.count(lj);
      for (int32_t a = 0; a < nx; ++a) {
        const float wx = tx.weight(li, a);
        const int32_t x = x0 + a;
        float row_l = 0.0f, row_o = 0.0f;
        for (int32_t k = 0; k < ny; ++k) {
          const int32_t y = y0 + k;
          const int32_t cell = x * cfg.gy + y;
          const CodeWord<A> cw = load_code<A>(codes, cell >> 2);
          const uint32_t bit = 1u << (cell & 3);
          // actor/transformations.py:62-83: own measured cells 1, received peers' cells 0, else 0.5
          float o = 0.5f;
#pragma unroll
          for (int j = 0; j < A; ++j)
            if (j != i && ((received >> j) & 1u) && (cw.byte(j) & bit)) o = 0.0f;
          if (cw.byte(i) & bit) o = 1.0f;
          const float wy = ty.weight(lj, k);
          row_l += wy * from_odds_fast(local[cell]);  // the maps hold odds
          row_o += wy * o;
        }
        pl += wx * row_l;
        pf_own += wx * row_o;
      }
    }
    // ---- pooled footprint image (taps over the raw footprint window) ----
    float pimg = 0.0f;
    {
      const int32_t u0 = fx.start(li), nu = fx.count(li), v0 = fy.start(lj), nv = fy.count(lj);
      for (int32_t a = 0; a < nu; ++a) {
        const int32_t u = u0 + a;
        float row = 0.0f;
        for (int32_t k = 0; k < nv; ++k) {
          const int32_t v = v0 + k;
          float val = 0.5f;
          if (u >= b_xl && u < b_xr && v >= b_yu && v < b_yd) {
            const int32_t cell = (me.xl + (u - b_xl)) * cfg.gy + (me.yu + (v - b_yu));
            const uint32_t byte = load_code<A>(codes, cell >> 2).byte(i);
            val = ((byte >> (4 + (cell & 3))) & 1u) ? y_hi : y_lo;
          }
          row += fy.weight(lj, k) * val;
        }
        pimg += fx.weight(li, a) * row;
      }
    }
    // ---- ego-centred position map: actor/transformations.py:110-176 (window hard-wired around index 5) ----
    float pm = 1.0f;
    if (me.ix < 5 && li < 5 - me.ix) pm = 0.0f;
    if (me.iy < 5 && lj < 5 - me.iy) pm = 0.0f;
    if (me.ix > 5 && li >= cfg.px - 1 - (me.ix - 6)) pm = 0.0f;
    if (me.iy > 5 && lj >= cfg.py - 1 - (me.iy - 6)) pm = 0.0f;
    if (li == 5 && lj == 5) pm = own_alt;  // only written when inside [0, px) x [0, px): (5,5) always is
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == i || !((received >> j) & 1u)) continue;
      const int32_t ri = s_geo[j].ix - me.ix + 5, rj = s_geo[j].iy - me.iy + 5;
      if (ri >= 0 && ri < cfg.px && rj >= 0 && rj < cfg.px && ri == li && rj == lj)
        pm = (float)(s_geo[j].zi + 1) / (float)(cfg.n_alt + 1);
    }
    const float2 wl = w_entropy(cfg, pl);
    const float2 wf = w_entropy(cfg, pimg);
    float* o = obs_out + (((int64_t)b * A + i) * cfg.px * cfg.py + c) * 7;
    o[0] = budget;
    o[1] = agent_id;
    o[2] = pm;
    o[3] = wl.x;
    o[4] = wf.x;
    o[5] = wl.y;
    o[6] = pf_own;
  }
}

// ------------------------------------------------------------------------------------------------
// critic state: one block per env
// ------------------------------------------------------------------------------------------------
template <int A>
__global__ void __launch_bounds__(128)
    features_critic_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const PoolTables pt,
                           const int32_t* __restrict__ pos_in, const int32_t* __restrict__ actions, const int32_t t,
                           const float* __restrict__ obs_in, float* __restrict__ state_out) {
  const int32_t b = blockIdx.x;
  __shared__ AgentGeo s_geo[A];
  __shared__ int32_t s_act[A];
  if (threadIdx.x < A) {
    s_geo[threadIdx.x] = agent_geo(cfg, pos_in + ((int64_t)b * A + threadIdx.x) * 3);
    s_act[threadIdx.x] = actions[(int64_t)b * A + threadIdx.x];
  }
  __syncthreads();
  const float* glob = st.global_map + (int64_t)b * cfg.map_stride;
  const uint8_t* codes = st.meas_codes + code_row_offset(cfg, t, b);
  const Tab tx = get_tab(pt, 0), ty = get_tab(pt, 1);
  for (int32_t c = threadIdx.x; c < cfg.px * cfg.py; c += blockDim.x) {
    const int32_t li = c / cfg.py, lj = c - li * cfg.py;
    float pg = 0.0f, pu = 0.0f;
    const int32_t x0 = tx.start(li), nx = tx.count(li), y0 = ty.start(lj), ny = ty.count(lj);
    for (int32_t a = 0; a < nx; ++a) {
      const int32_t x = x0 + a;
      float row_g = 0.0f, row_u = 0.0f;
      for (int32_t k = 0; k < ny; ++k) {
        const int32_t cell = x * cfg.gy + (y0 + k);
        const CodeWord<A> cw = load_code<A>(codes, cell >> 2);
        const uint32_t bit = 1u << (cell & 3);
        float u = 0.5f;  // critic/transformations.py:91-108: any agent's measured cell -> 1, else 0.5
#pragma unroll
        for (int j = 0; j < A; ++j)
          if (cw.byte(j) & bit) u = 1.0f;
        const float wy = ty.weight(lj, k);
        row_g += wy * from_odds_fast(glob[cell]);  // the maps hold odds
        row_u += wy * u;
      }
      pg += tx.weight(li, a) * row_g;
      pu += tx.weight(li, a) * row_u;
    }
    const float2 wg = w_entropy(cfg, pg);
    float posv = 0.0f;  // critic/transformations.py:70-88 (later agents overwrite earlier ones)
#pragma unroll
    for (int j = 0; j < A; ++j)
      if (s_geo[j].ix == li && s_geo[j].iy == lj) posv = (float)(s_geo[j].zi + 1) / (float)cfg.n_alt;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      float actv = 0.0f;  // :111-132: the other agents' actions at their PRE-move cells
#pragma unroll
      for (int j = 0; j < A; ++j)
        if (j != i && s_geo[j].ix == li && s_geo[j].iy == lj) actv = (float)(s_act[j] + 1) / (float)IPP_N_ACTIONS;
      const float* o = obs_in + (((int64_t)b * A + i) * cfg.px * cfg.py + c) * 7;
      float* s = state_out + (((int64_t)b * A + i) * cfg.px * cfg.py + c) * 12;
#pragma unroll
      for (int k = 0; k < 7; ++k) s[k] = o[k];
      s[7] = posv;
      s[8] = wg.x;
      s[9] = wg.y;
      s[10] = pu;
      s[11] = actv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// staged variants (grids whose map fits in shared memory): the map and the per-cell footprint values are
// staged once with coalesced loads, then pooled separably (along y, then along x) out of shared memory.
// ------------------------------------------------------------------------------------------------
// Fixed-tap pooling: the tap tables are zero padded to maxt <= MT, so a fully unrolled MT-tap loop gives the same
// sums without per-tap loop control.  A zero-weight tap may index up to MT-1 elements past the end of its source
// array; the shared-memory layout puts finite data (zero pads or arrays already written) behind every pooled array,
// so those taps contribute exactly +0.  Thread (rr, lj) keeps the taps of lattice column lj in registers and
// walks rows rr, rr + rpr, ...
template <int MT>
struct Taps {
  float w[MT];
  int32_t first;
};

template <int MT>
__device__ __forceinline__ Taps<MT> load_taps(const Tab& t, int32_t d) {
  Taps<MT> r;
  r.first = t.start(d);
#pragma unroll
  for (int k = 0; k < MT; ++k) r.w[k] = k < t.maxt ? t.weight(d, k) : 0.0f;
  return r;
}

__device__ __forceinline__ float tap_value(const float* s, int k) { return s[k]; }
__device__ __forceinline__ float tap_value(const uint8_t* s, int k) { return (float)s[k]; }

// tmp[r][lj] = scale * sum_k w[k] * src[r][first + k]
template <int MT, typename T>
__device__ __forceinline__ void pool_rows(const Taps<MT>& tp, const T* __restrict__ src, float scale,
                                          float* __restrict__ tmp, int32_t n_rows, int32_t src_w, int32_t py,
                                          int32_t rr, int32_t rpr, int32_t lj) {
  if (rr >= rpr) return;
  for (int32_t r = rr; r < n_rows; r += rpr) {
    const T* s = src + r * src_w + tp.first;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < MT; ++k) acc = fmaf(tp.w[k], tap_value(s, k), acc);
    tmp[r * py + lj] = acc * scale;
  }
}

template <int MT>
__device__ __forceinline__ float pool_col(const Taps<MT>& tp, const float* __restrict__ tmp, int32_t lj, int32_t py) {
  const float* s = tmp + tp.first * py + lj;
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < MT; ++k) acc = fmaf(tp.w[k], s[k * py], acc);
  return acc;
}

// low nibbles (cell-in-footprint bits) of the bytes selected by `mask`, OR-folded into one nibble
template <int A>
__device__ __forceinline__ uint32_t fold_nibbles(const CodeWord<A>& cw, const uint32_t (&mask)[CodeWord<A>::WORDS]) {
  uint32_t x = 0;
#pragma unroll
  for (int k = 0; k < CodeWord<A>::WORDS; ++k) x |= cw.w[k] & mask[k];
  x |= x >> 16;
  x |= x >> 8;
  return x & 0xFu;
}

// bit c of a nibble -> byte c of a word (0 / 1)
__device__ __forceinline__ uint32_t spread_nibble(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ unsigned char* align16(void* p) {
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15u) & ~(uintptr_t)15u);
}

constexpr int FEAT_PAD = 8;  // floats / bytes behind a row-pooled array (>= MT - 1)

template <int A, int MT>
__global__ void __launch_bounds__(128, 9)
    features_actor_staged_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const PoolTables pt,
                                 const int32_t* __restrict__ pos_in, const uint8_t* __restrict__ comm,
                                 const int32_t t, float* __restrict__ obs_out) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int32_t n_cells = cfg.gx * cfg.gy, n_quads = (n_cells + 3) >> 2;
  const int32_t n_lat = cfg.px * cfg.py;
  const int32_t b = blockIdx.x / A, i = blockIdx.x - b * A;
  __shared__ AgentGeo s_geo[A];
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x < A) s_geo[threadIdx.x] = agent_geo(cfg, pos_in + ((int64_t)b * A + threadIdx.x) * 3);
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&s_bar), 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const AgentGeo me = s_geo[i];
  const int32_t h = me.ry2, w = me.rx2;
  // layout (every pooled array is followed by finite data, see Taps):
  //   s_map | pad | s_tl | s_to | s_ti | column pad | s_img | pad | s_own (bytes) | pad | codes
  float* s_map = reinterpret_cast<float*>(fsm);      // [map_stride] fused local map, filled by the bulk copy
  float* s_tl = s_map + cfg.map_stride + FEAT_PAD;   // [gx][py] row-pooled local map
  float* s_to = s_tl + cfg.gx * cfg.py;              // [gx][py] row-pooled ownership
  float* s_ti = s_to + cfg.gx * cfg.py;              // [h][py]  row-pooled footprint image
  float* s_cpad = s_ti + h * cfg.py;                 // [MT][py] zeros
  float* s_img = s_cpad + MT * cfg.py;               // [h][w]   footprint image
  uint8_t* s_own = reinterpret_cast<uint8_t*>(s_img + h * w + FEAT_PAD);  // [4 n_quads] 2 own / 1 unseen / 0 others
  unsigned char* codes = align16(s_own + 4 * n_quads + FEAT_PAD);         // [code_stride] this env's codes
  if (threadIdx.x == 0) {  // map + code row by the bulk-copy engine while the block sets up its tap registers
    const uint32_t bar = ptx::smem_u32(&s_bar), map_bytes = (uint32_t)n_quads * 16u;
    ptx::mbar_arrive_expect_tx(bar, map_bytes + (uint32_t)cfg.code_stride);
    ptx::bulk_load(ptx::smem_u32(s_map), st.local_maps + ((int64_t)b * A + i) * cfg.map_stride, map_bytes, bar);
    ptx::bulk_load(ptx::smem_u32(codes),
                   st.meas_codes + code_row_offset(cfg, t, b), (uint32_t)cfg.code_stride, bar);
  }
  // zero pads (disjoint from the bulk-copy destinations)
  for (int32_t k = 4 * n_quads + (int32_t)threadIdx.x; k < cfg.map_stride + FEAT_PAD; k += blockDim.x) s_map[k] = 0.0f;
  for (int32_t k = threadIdx.x; k < MT * cfg.py; k += blockDim.x) s_cpad[k] = 0.0f;
  if (threadIdx.x < FEAT_PAD) {
    s_img[h * w + threadIdx.x] = 0.0f;
    s_own[4 * n_quads + threadIdx.x] = 0;
  }
  const uint32_t received = comm[(int64_t)b * A + i];
  const Tab tx = get_tab(pt, 0), ty = get_tab(pt, 1);
  const Tab fx = get_tab(pt, 2 + 2 * me.iz), fy = get_tab(pt, 3 + 2 * me.iz);
  int32_t b_yu = 0, b_yd = h, b_xl = 0, b_xr = w;
  if (me.yu > me.ryu) b_yu = h - (me.yd - me.yu);
  if (me.yd < me.ryu + me.ry2) b_yd = me.yd - me.yu;
  if (me.xr < me.rxl + me.rx2) b_xr = me.xr - me.xl;
  if (me.xl > me.rxl) b_xl = w - (me.xr - me.xl);
  const float y_hi = cfg.y_hi[me.iz], y_lo = cfg.y_lo[me.iz];

  uint32_t others_mask[CodeWord<A>::WORDS] = {};
#pragma unroll
  for (int j = 0; j < A; ++j)
    if (j != i && ((received >> j) & 1u)) others_mask[j >> 2] |= 0x0Fu << (8 * (j & 3));
  // thread (rr, lj_r): lattice column lj_r of rows rr, rr + rpr, ... in the row pass; lattice cell (rr, lj_r) in the
  // column pass (first round)
  const int32_t rpr = (int32_t)blockDim.x / cfg.py;
  const int32_t rr = (int32_t)threadIdx.x / cfg.py, lj_r = (int32_t)threadIdx.x - rr * cfg.py;
  const Taps<MT> tap_y = load_taps<MT>(ty, lj_r);
  const Taps<MT> tap_fy = load_taps<MT>(fy, lj_r);
  const int32_t li_r = min(rr, cfg.px - 1);
  Taps<MT> tap_x = load_taps<MT>(tx, li_r);
  Taps<MT> tap_fx = load_taps<MT>(fx, li_r);
  ptx::mbar_wait(ptx::smem_u32(&s_bar), 0);
  for (int32_t q = threadIdx.x; q < n_quads; q += blockDim.x) {
    const CodeWord<A> cw = load_code<A>(codes, q);
    const uint32_t own = cw.byte(i) & 0xFu, oth = fold_nibbles<A>(cw, others_mask) & ~own;
    reinterpret_cast<uint32_t*>(s_own)[q] = 0x01010101u + spread_nibble(own) - spread_nibble(oth);
    const float4 o4 = reinterpret_cast<float4*>(s_map)[q];  // the staged map holds odds: to probabilities, in place
    reinterpret_cast<float4*>(s_map)[q] = make_float4(from_odds_fast(o4.x), from_odds_fast(o4.y), from_odds_fast(o4.z), from_odds_fast(o4.w));
  }
  {
    // footprint image, element (u, v) walked without per-element divisions
    const int32_t du = (int32_t)blockDim.x / w, dv = (int32_t)blockDim.x - du * w;
    int32_t u = (int32_t)threadIdx.x / w, v = (int32_t)threadIdx.x - u * w;
    for (int32_t idx = threadIdx.x; idx < h * w; idx += blockDim.x) {
      float val = 0.5f;
      if (u >= b_xl && u < b_xr && v >= b_yu && v < b_yd) {
        const int32_t cell = (me.xl + (u - b_xl)) * cfg.gy + (me.yu + (v - b_yu));
        const uint32_t byte = load_code<A>(codes, cell >> 2).byte(i);
        val = ((byte >> (4 + (cell & 3))) & 1u) ? y_hi : y_lo;
      }
      s_img[idx] = val;
      u += du;
      v += dv;
      if (v >= w) {
        v -= w;
        ++u;
      }
    }
  }
  __syncthreads();
  pool_rows<MT>(tap_y, s_map, 1.0f, s_tl, cfg.gx, cfg.gy, cfg.py, rr, rpr, lj_r);
  pool_rows<MT>(tap_y, s_own, 0.5f, s_to, cfg.gx, cfg.gy, cfg.py, rr, rpr, lj_r);
  pool_rows<MT>(tap_fy, s_img, 1.0f, s_ti, h, w, cfg.py, rr, rpr, lj_r);
  __syncthreads();
  const float budget = (float)(cfg.budget - t) / (float)cfg.budget;
  const float agent_id = (float)(i + 1) / (float)A;
  const float own_alt = (float)(me.zi + 1) / (float)(cfg.n_alt + 1);
  for (int32_t c = threadIdx.x; c < n_lat; c += blockDim.x) {
    int32_t li = rr, lj = lj_r;
    if (c != (int32_t)threadIdx.x) {
      li = c / cfg.py;
      lj = c - li * cfg.py;
      tap_x = load_taps<MT>(tx, li);
      tap_fx = load_taps<MT>(fx, li);
    }
    const float pl = pool_col<MT>(tap_x, s_tl, lj, cfg.py);
    const float pf_own = pool_col<MT>(tap_x, s_to, lj, cfg.py);
    const float pimg = pool_col<MT>(tap_fx, s_ti, lj, cfg.py);
    float pm = 1.0f;
    if (me.ix < 5 && li < 5 - me.ix) pm = 0.0f;
    if (me.iy < 5 && lj < 5 - me.iy) pm = 0.0f;
    if (me.ix > 5 && li >= cfg.px - 1 - (me.ix - 6)) pm = 0.0f;
    if (me.iy > 5 && lj >= cfg.py - 1 - (me.iy - 6)) pm = 0.0f;
    if (li == 5 && lj == 5) pm = own_alt;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == i || !((received >> j) & 1u)) continue;
      const int32_t ri = s_geo[j].ix - me.ix + 5, rj = s_geo[j].iy - me.iy + 5;
      if (ri >= 0 && ri < cfg.px && rj >= 0 && rj < cfg.px && ri == li && rj == lj)
        pm = (float)(s_geo[j].zi + 1) / (float)(cfg.n_alt + 1);
    }
    const float2 wl = w_entropy(cfg, pl);
    const float2 wf = w_entropy(cfg, pimg);
    float* o = obs_out + (((int64_t)b * A + i) * n_lat + c) * 7;
    o[0] = budget;
    o[1] = agent_id;
    o[2] = pm;
    o[3] = wl.x;
    o[4] = wf.x;
    o[5] = wl.y;
    o[6] = pf_own;
  }
}

template <int A, int MT>
__global__ void __launch_bounds__(128)
    features_critic_staged_kernel(const __grid_constant__ ipp_config cfg, const ipp_state st, const PoolTables pt,
                                  const int32_t* __restrict__ pos_in, const int32_t* __restrict__ actions,
                                  const int32_t t, const float* __restrict__ obs_in, float* __restrict__ state_out) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int32_t n_cells = cfg.gx * cfg.gy, n_quads = (n_cells + 3) >> 2;
  const int32_t n_lat = cfg.px * cfg.py;
  // layout: s_map | pad | s_tg | s_tu | column pad | s_uni (bytes) | pad | codes
  float* s_map = reinterpret_cast<float*>(fsm);     // [map_stride] global map, filled by the bulk copy
  float* s_tg = s_map + cfg.map_stride + FEAT_PAD;  // [gx][py]
  float* s_tu = s_tg + cfg.gx * cfg.py;             // [gx][py]
  float* s_cpad = s_tu + cfg.gx * cfg.py;           // [MT][py] zeros
  uint8_t* s_uni = reinterpret_cast<uint8_t*>(s_cpad + MT * cfg.py);  // [4 n_quads] 2 inside some footprint / 1
  unsigned char* codes = align16(s_uni + 4 * n_quads + FEAT_PAD);     // [code_stride]
  const int32_t b = blockIdx.x;
  __shared__ AgentGeo s_geo[A];
  __shared__ int32_t s_act[A];
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x < A) {
    s_geo[threadIdx.x] = agent_geo(cfg, pos_in + ((int64_t)b * A + threadIdx.x) * 3);
    s_act[threadIdx.x] = actions[(int64_t)b * A + threadIdx.x];
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&s_bar), 1);
    ptx::fence_mbar_init();
    const uint32_t bar = ptx::smem_u32(&s_bar), map_bytes = (uint32_t)n_quads * 16u;
    ptx::mbar_arrive_expect_tx(bar, map_bytes + (uint32_t)cfg.code_stride);
    ptx::bulk_load(ptx::smem_u32(s_map), st.global_map + (int64_t)b * cfg.map_stride, map_bytes, bar);
    ptx::bulk_load(ptx::smem_u32(codes),
                   st.meas_codes + code_row_offset(cfg, t, b), (uint32_t)cfg.code_stride, bar);
  }
  for (int32_t k = 4 * n_quads + (int32_t)threadIdx.x; k < cfg.map_stride + FEAT_PAD; k += blockDim.x) s_map[k] = 0.0f;
  for (int32_t k = threadIdx.x; k < MT * cfg.py; k += blockDim.x) s_cpad[k] = 0.0f;
  if (threadIdx.x < FEAT_PAD) s_uni[4 * n_quads + threadIdx.x] = 0;
  uint32_t all_mask[CodeWord<A>::WORDS] = {};
#pragma unroll
  for (int j = 0; j < A; ++j) all_mask[j >> 2] |= 0x0Fu << (8 * (j & 3));
  const Tab tx = get_tab(pt, 0), ty = get_tab(pt, 1);
  const int32_t rpr = (int32_t)blockDim.x / cfg.py;
  const int32_t rr = (int32_t)threadIdx.x / cfg.py, lj_r = (int32_t)threadIdx.x - rr * cfg.py;
  const Taps<MT> tap_y = load_taps<MT>(ty, lj_r);
  Taps<MT> tap_x = load_taps<MT>(tx, min(rr, cfg.px - 1));
  __syncthreads();  // s_bar initialised, s_geo / s_act visible
  ptx::mbar_wait(ptx::smem_u32(&s_bar), 0);
  for (int32_t q = threadIdx.x; q < n_quads; q += blockDim.x) {
    const uint32_t any = fold_nibbles<A>(load_code<A>(codes, q), all_mask);
    reinterpret_cast<uint32_t*>(s_uni)[q] = 0x01010101u + spread_nibble(any);
    const float4 o4 = reinterpret_cast<float4*>(s_map)[q];  // odds -> probabilities, in place
    reinterpret_cast<float4*>(s_map)[q] = make_float4(from_odds_fast(o4.x), from_odds_fast(o4.y), from_odds_fast(o4.z), from_odds_fast(o4.w));
  }
  __syncthreads();
  pool_rows<MT>(tap_y, s_map, 1.0f, s_tg, cfg.gx, cfg.gy, cfg.py, rr, rpr, lj_r);
  pool_rows<MT>(tap_y, s_uni, 0.5f, s_tu, cfg.gx, cfg.gy, cfg.py, rr, rpr, lj_r);
  __syncthreads();
  for (int32_t c = threadIdx.x; c < n_lat; c += blockDim.x) {
    int32_t li = rr, lj = lj_r;
    if (c != (int32_t)threadIdx.x) {
      li = c / cfg.py;
      lj = c - li * cfg.py;
      tap_x = load_taps<MT>(tx, li);
    }
    const float2 wg = w_entropy(cfg, pool_col<MT>(tap_x, s_tg, lj, cfg.py));
    const float pu = pool_col<MT>(tap_x, s_tu, lj, cfg.py);
    float posv = 0.0f;
#pragma unroll
    for (int j = 0; j < A; ++j)
      if (s_geo[j].ix == li && s_geo[j].iy == lj) posv = (float)(s_geo[j].zi + 1) / (float)cfg.n_alt;
#pragma unroll
    for (int i = 0; i < A; ++i) {
      float actv = 0.0f;
#pragma unroll
      for (int j = 0; j < A; ++j)
        if (j != i && s_geo[j].ix == li && s_geo[j].iy == lj) actv = (float)(s_act[j] + 1) / (float)IPP_N_ACTIONS;
      const float* o = obs_in + (((int64_t)b * A + i) * n_lat + c) * 7;
      float4* sdst = reinterpret_cast<float4*>(state_out + (((int64_t)b * A + i) * n_lat + c) * 12);  // 48 B cells
      sdst[0] = make_float4(o[0], o[1], o[2], o[3]);
      sdst[1] = make_float4(o[4], o[5], o[6], posv);
      sdst[2] = make_float4(wg.x, wg.y, pu, actv);
    }
  }
}

constexpr int FEAT_MAX_TAPS = 8;  // staged kernels are instantiated for 4 / 6 / 8 taps

static size_t staged_actor_smem(const ipp_config& cfg) {
  int hw = 0, hmax = 0;
  for (int a = 0; a < cfg.n_alt; ++a) {
    hw = std::max(hw, 4 * cfg.radius_x[a] * cfg.radius_y[a]);
    hmax = std::max(hmax, 2 * cfg.radius_y[a]);
  }
  const size_t n_quads = ((size_t)cfg.gx * cfg.gy + 3) / 4;
  return sizeof(float) * ((size_t)cfg.map_stride + FEAT_PAD + 2 * (size_t)cfg.gx * cfg.py + (size_t)hmax * cfg.py +
                          (size_t)FEAT_MAX_TAPS * cfg.py + hw + FEAT_PAD) +
         4 * n_quads + FEAT_PAD + 16 + (size_t)cfg.code_stride;
}
static size_t staged_critic_smem(const ipp_config& cfg) {
  const size_t n_quads = ((size_t)cfg.gx * cfg.gy + 3) / 4;
  return sizeof(float) * ((size_t)cfg.map_stride + FEAT_PAD + 2 * (size_t)cfg.gx * cfg.py +
                          (size_t)FEAT_MAX_TAPS * cfg.py) +
         4 * n_quads + FEAT_PAD + 16 + (size_t)cfg.code_stride;
}
constexpr size_t FEAT_SMEM_LIMIT = 100 * 1024;

template <int A, int MT>
static cudaError_t actor_staged_mt(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                   const int32_t* pos_in, const uint8_t* comm, int32_t t, float* obs_out, size_t smem,
                                   cudaStream_t s) {
  auto kern = features_actor_staged_kernel<A, MT>;
  static PerDevice attr;
  if (!attr.cur()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FEAT_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    attr.cur() = 1;
  }
  kern<<<(unsigned)cfg.n_envs * A, 128, smem, s>>>(cfg, st, pt, pos_in, comm, t, obs_out);
  return cudaGetLastError();
}

template <int A>
static cudaError_t actor_staged(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt, const int32_t* pos_in,
                                const uint8_t* comm, int32_t t, float* obs_out, size_t smem, cudaStream_t s) {
  if (pt.maxt <= 4) return actor_staged_mt<A, 4>(cfg, st, pt, pos_in, comm, t, obs_out, smem, s);
  if (pt.maxt <= 6) return actor_staged_mt<A, 6>(cfg, st, pt, pos_in, comm, t, obs_out, smem, s);
  return actor_staged_mt<A, 8>(cfg, st, pt, pos_in, comm, t, obs_out, smem, s);
}

template <int A, int MT>
static cudaError_t critic_staged_mt(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                    const int32_t* pos_in, const int32_t* actions, int32_t t, const float* obs_in,
                                    float* state_out, size_t smem, cudaStream_t s) {
  auto kern = features_critic_staged_kernel<A, MT>;
  static PerDevice attr;
  if (!attr.cur()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FEAT_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    attr.cur() = 1;
  }
  kern<<<(unsigned)cfg.n_envs, 128, smem, s>>>(cfg, st, pt, pos_in, actions, t, obs_in, state_out);
  return cudaGetLastError();
}

template <int A>
static cudaError_t critic_staged(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                 const int32_t* pos_in, const int32_t* actions, int32_t t, const float* obs_in,
                                 float* state_out, size_t smem, cudaStream_t s) {
  if (pt.maxt <= 4) return critic_staged_mt<A, 4>(cfg, st, pt, pos_in, actions, t, obs_in, state_out, smem, s);
  if (pt.maxt <= 6) return critic_staged_mt<A, 6>(cfg, st, pt, pos_in, actions, t, obs_in, state_out, smem, s);
  return critic_staged_mt<A, 8>(cfg, st, pt, pos_in, actions, t, obs_in, state_out, smem, s);
}

#define IPP_FEAT_DISPATCH(A_, CALL)                \
  switch (A_) {                                    \
    case 1: { constexpr int kA = 1; CALL; } break; \
    case 2: { constexpr int kA = 2; CALL; } break; \
    case 3: { constexpr int kA = 3; CALL; } break; \
    case 4: { constexpr int kA = 4; CALL; } break; \
    case 5: { constexpr int kA = 5; CALL; } break; \
    case 6: { constexpr int kA = 6; CALL; } break; \
    case 7: { constexpr int kA = 7; CALL; } break; \
    case 8: { constexpr int kA = 8; CALL; } break; \
    default: return cudaErrorInvalidValue;         \
  }

cudaError_t launch_features_actor(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                  const int32_t* pos_in, const uint8_t* comm, int32_t t, float* obs_out,
                                  cudaStream_t s) {
  const size_t smem = staged_actor_smem(cfg);
  if (smem <= FEAT_SMEM_LIMIT && pt.maxt <= FEAT_MAX_TAPS && cfg.py <= 128) {
    IPP_FEAT_DISPATCH(cfg.n_agents, return actor_staged<kA>(cfg, st, pt, pos_in, comm, t, obs_out, smem, s));
  }
  IPP_FEAT_DISPATCH(cfg.n_agents, (features_actor_kernel<kA><<<(unsigned)cfg.n_envs * kA, 128, 0, s>>>(
                                      cfg, st, pt, pos_in, comm, t, obs_out)));
  return cudaGetLastError();
}

cudaError_t launch_features_critic(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                   const int32_t* pos_in, const int32_t* actions, int32_t t, const float* obs_in,
                                   float* state_out, cudaStream_t s) {
  const size_t smem = staged_critic_smem(cfg);
  const bool aligned = (reinterpret_cast<uintptr_t>(state_out) & 15u) == 0;  // float4 stores of the 48 B cells
  if (smem <= FEAT_SMEM_LIMIT && pt.maxt <= FEAT_MAX_TAPS && cfg.py <= 128 && aligned) {
    IPP_FEAT_DISPATCH(cfg.n_agents,
                      return critic_staged<kA>(cfg, st, pt, pos_in, actions, t, obs_in, state_out, smem, s));
  }
  IPP_FEAT_DISPATCH(cfg.n_agents, (features_critic_kernel<kA><<<(unsigned)cfg.n_envs, 128, 0, s>>>(
                                      cfg, st, pt, pos_in, actions, t, obs_in, state_out)));
  return cudaGetLastError();
}

}  // namespace ipp

// Host-visible launch plans + launcher prototypes (kernels live in ipp_kernels.cu / ipp_step_tma.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ipp_b200.h"

namespace ipp {

constexpr int STEP_THREADS = 256;   // direct-load variant: threads per (env, chunk) block
constexpr int TMA_QPC = 640;             // TMA variant: quads per work item (20 tiles of 32 quads, 10 KB per map)
// Warps of the TMA map kernel: consumer warps pulling (item, tile) tasks + 1 producer warp + 1 finisher warp.
// A <= 4: 32 warps (64 registers per thread); A > 4: 28 warps (72 registers).
__host__ __device__ constexpr int tma_consumer_warps(int n_agents) { return n_agents <= 4 ? 30 : 26; }
__host__ __device__ constexpr int tma_threads(int n_agents) { return (tma_consumer_warps(n_agents) + 2) * 32; }

// Launch-time state that is per DEVICE (function attributes such as the dynamic shared-memory limit belong to the
// device a kernel is launched on): one slot per device ordinal, so that handles on several GPUs can live in one process.
struct PerDevice {
  int v[64] = {};
  int& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return v[d & 63];
  }
};

struct LaunchPlan {
  int32_t n_chunks;         // chunks per env map (1 => per-env reward finishes inside the block)
  int32_t quads_per_chunk;  // float4 groups of cells per chunk
};

struct TmaPlan {
  int32_t n_chunks, quads_per_chunk, slot_bytes, env_bytes, d_map, d_env, smem_bytes;
  bool ok;
};

TmaPlan plan_tma(const ipp_config& cfg, int max_smem_optin);
cudaError_t configure_step_tma(const ipp_config& cfg, const TmaPlan& plan);
cudaError_t configure_plan();

// cv2.INTER_AREA tap tables of the feature builders (device memory, owned by the handle)
struct PoolTables {
  int32_t* ints;  // [n_tabs][IPP_MAX_LATTICE][2]: first source index, tap count
  float* w;       // [n_tabs][IPP_MAX_LATTICE][maxt]
  int32_t maxt, n_tabs;
};
cudaError_t build_pool_tables(const ipp_config& cfg, PoolTables* pt);
void free_pool_tables(PoolTables* pt);
cudaError_t launch_features_actor(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                  const int32_t* pos_in, const uint8_t* comm, int32_t t, float* obs_out,
                                  cudaStream_t s);
cudaError_t launch_features_critic(const ipp_config& cfg, const ipp_state& st, const PoolTables& pt,
                                   const int32_t* pos_in, const int32_t* actions, int32_t t, const float* obs_in,
                                   float* state_out, cudaStream_t s);

// lut: device table [n_alt][256] float4 = odds multipliers of the 4 cells of a quad for a code byte
// step_meta: device scratch [n_envs][4 * n_agents] uint32 — the EnvMeta record (ipp_cell.cuh) of every env, written
// by the plan kernel (comm bits, LUT rows of the communicated / new measurements) and read by the map kernels
cudaError_t launch_plan(const ipp_config& cfg, const ipp_state& st, const ipp_step_io& io, int32_t t, int do_comm,
                        int do_move, uint32_t* step_meta, const int32_t* gt_params, cudaStream_t s);
cudaError_t launch_step_dense(const ipp_config& cfg, const ipp_state& st, const float4* lut,
                              const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                              double* partials, bool do_own, cudaStream_t s);
cudaError_t launch_step_tma(const ipp_config& cfg, const ipp_state& st, const float4* lut, const TmaPlan& plan,
                            int n_sm, const uint32_t* step_meta, int32_t t, float* reward_rel, float* reward_abs,
                            double* partials, bool do_own, cudaStream_t s);
cudaError_t launch_reward_finalize(const ipp_config& cfg, const double* partials, int32_t n_chunks, float* reward_rel,
                                   float* reward_abs, cudaStream_t s);
cudaError_t launch_own_update(const ipp_config& cfg, const ipp_state& st, const float4* lut, const int32_t* pos_out,
                              int32_t t, cudaStream_t s);
cudaError_t launch_reset(const ipp_config& cfg, const ipp_state& st, const float4* lut, const LaunchPlan& plan,
                         int32_t* pos_out, int32_t* gt_params, cudaStream_t s);

// stored odds -> probabilities (n_floats multiple of 4)
cudaError_t launch_export_beliefs(const float* src, float* dst, int64_t n_floats, cudaStream_t s);

// IG-greedy planner + evaluation metrics (ipp_planner.cu)
cudaError_t launch_ig_plan(const ipp_config& cfg, const ipp_state& st, const int32_t* pos_in, int communication,
                           int32_t* actions_out, uint8_t* mask_out, double* gains_out, double* util_out,
                           cudaStream_t s);
cudaError_t launch_eval_metrics(const ipp_config& cfg, const ipp_state& st, double* entropy_out, double* f1_out,
                                cudaStream_t s);

// single-map helpers used by the facade entry points (device pointers; dtype flags: 0 = float32, 1 = float64)
cudaError_t launch_update_cells(const ipp_config& cfg, void* x, int x_f64, const void* y, int y_f64, int y_is_scalar,
                                int64_t n, double* out, cudaStream_t s);
cudaError_t launch_fuse_map(const ipp_config& cfg, const float* own, const float* others, int n_others, int64_t n,
                            double* out, cudaStream_t s);
cudaError_t launch_measure(const ipp_config& cfg, const uint8_t* gt, const int32_t* rect, uint32_t key,
                           uint32_t thresh, float y_hi, float y_lo, float* out, cudaStream_t s);
cudaError_t launch_entropy(void* p, int is_f64, int64_t n, void* out, cudaStream_t s);
cudaError_t launch_utility_reward(const void* last, int last_f64, const void* next, int next_f64, int64_t n,
                                  double* out2, cudaStream_t s);

}  // namespace ipp

// C ABI (include/ipp_b200.h): argument validation, launch planning, scratch ownership.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ipp_cell.cuh"
#include "ipp_launch.h"

struct ipp_handle {
  ipp_config cfg;
  ipp::LaunchPlan plan;   // direct-load variant (also used by reset)
  ipp::TmaPlan tma;       // TMA-staged variant
  int variant;            // IPP_VARIANT_DIRECT / IPP_VARIANT_TMA
  int n_sm;
  int device;
  double* partials;     // [n_envs, n_chunks, 2]
  int32_t* gt_params;   // [n_envs, 4]
  uint8_t* comm;        // [n_envs, n_agents] when the caller does not ask for comm_out
  uint32_t* step_meta;  // [n_envs, n_seg, rec_words] ItemRec of every work item, handed from the plan kernel to the map kernels
  void* policy_in;      // [n_envs, n_agents, 6] float32 staging of ipp_step_host's policy input (lazily allocated)
  float4* lut;          // [n_alt, 256] odds multipliers of a quad for every measurement code byte
  ipp::PoolTables pool; // cv2.INTER_AREA tap tables of the feature builders
  // facade scratch (grown on demand)
  // CUDA graphs of ipp_run_steps, keyed by a hash of everything that is baked into them
  struct GraphSlot {
    uint64_t key;
    cudaGraphExec_t exec;
  } graphs[8];
  int n_graphs;
  void* fbuf;
  size_t fbuf_bytes;
  int64_t scratch_bytes;
  char err[256];
};

namespace {

int fail_cuda(ipp_handle* h, cudaError_t e, const char* where) {
  if (h != nullptr) snprintf(h->err, sizeof(h->err), "%s: %s", where, cudaGetErrorString(e));
  return IPP_ERR_CUDA;
}

#define IPP_CUDA(h, call)                                 \
  do {                                                    \
    cudaError_t e_ = (call);                              \
    if (e_ != cudaSuccess) return fail_cuda(h, e_, #call); \
  } while (0)

cudaError_t launch_maps(ipp_handle* h, const ipp_state* st, const ipp_step_io& io, int32_t t, bool do_own,
                        cudaStream_t s) {
  if (h->variant == IPP_VARIANT_TMA) {
    cudaError_t e = ipp::launch_step_tma(h->cfg, *st, h->lut, h->tma, h->n_sm, h->step_meta, t, io.reward_rel,
                                         io.reward_abs, h->partials, do_own, s);
    if (e != cudaSuccess) return e;
    if (h->tma.n_chunks > 1)
      e = ipp::launch_reward_finalize(h->cfg, h->partials, h->tma.n_chunks, io.reward_rel, io.reward_abs, s);
    return e;
  }
  return ipp::launch_step_dense(h->cfg, *st, h->lut, h->step_meta, t, io.reward_rel, io.reward_abs, h->partials,
                                do_own, s);
}

// Every entry point works on the device its handle was created on, whatever the caller's current device is.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const ipp_handle* h) {
    int cur = -1;
    if (h != nullptr && cudaGetDevice(&cur) == cudaSuccess && cur != h->device) {
      prev = cur;
      cudaSetDevice(h->device);
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int validate(const ipp_config* c) {
  if (c == nullptr) return IPP_ERR_INVALID_ARG;
  if (c->gx <= 0 || c->gy < 4 || c->n_envs <= 0) return IPP_ERR_INVALID_ARG;  // a quad spans <= 2 rows
  if ((int64_t)c->gx * c->gy > (1 << 28)) return IPP_ERR_UNSUPPORTED;
  if (c->map_stride < c->gx * c->gy || (c->map_stride & 3) != 0) return IPP_ERR_INVALID_ARG;
  if (c->gt_stride < c->map_stride || (c->gt_stride & 15) != 0) return IPP_ERR_INVALID_ARG;
  {
    const int64_t need = (((int64_t)c->gx * c->gy + 3) / 4) * (c->n_agents <= 4 ? 4 : 8);
    if (c->code_stride < need || (c->code_stride & 15) != 0) return IPP_ERR_INVALID_ARG;
    const int64_t quads = ((int64_t)c->gx * c->gy + 3) / 4;
    if (c->n_seg != (int32_t)((quads + IPP_FLAG_QUADS - 1) / IPP_FLAG_QUADS)) return IPP_ERR_INVALID_ARG;
  }
  if (c->n_agents < 1 || c->n_agents > IPP_MAX_AGENTS) return IPP_ERR_UNSUPPORTED;
  if (c->n_alt < 1 || c->n_alt > IPP_MAX_ALT) return IPP_ERR_UNSUPPORTED;
  if (c->px < 1 || c->px > IPP_MAX_LATTICE || c->py < 1 || c->py > IPP_MAX_LATTICE) return IPP_ERR_UNSUPPORTED;
  if (c->spacing <= 0 || c->min_altitude % c->spacing != 0) return IPP_ERR_INVALID_ARG;
  if (!(c->prior > 0.0f && c->prior < 1.0f)) return IPP_ERR_INVALID_ARG;
  if (!(c->o_min > 0.0f) || !(c->o_max > c->o_min) || !(c->p_max > c->p_min)) return IPP_ERR_INVALID_ARG;
  for (int i = 0; i < c->n_alt; ++i)
    if (c->radius_x[i] < 0 || c->radius_y[i] < 0 || !(c->k_hi[i] > 0.0f) || !(c->k_lo[i] > 0.0f))
      return IPP_ERR_INVALID_ARG;
  for (int i = 0; i < c->px; ++i)
    if (c->cell_x[i] < 0) return IPP_ERR_INVALID_ARG;
  for (int i = 0; i < c->py; ++i)
    if (c->cell_y[i] < 0) return IPP_ERR_INVALID_ARG;
  return IPP_OK;
}

int check_state(const ipp_state* st) {
  if (st == nullptr || st->local_maps == nullptr || st->global_map == nullptr || st->ground_truth == nullptr ||
      st->episodes == nullptr || st->meas_codes == nullptr || st->map_flags == nullptr)
    return IPP_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(st->local_maps) & 15) || (reinterpret_cast<uintptr_t>(st->global_map) & 15) ||
      (reinterpret_cast<uintptr_t>(st->ground_truth) & 15) || (reinterpret_cast<uintptr_t>(st->meas_codes) & 15) ||
      (reinterpret_cast<uintptr_t>(st->map_flags) & 15))
    return IPP_ERR_INVALID_ARG;
  return IPP_OK;
}

int ensure_fbuf(ipp_handle* h, size_t bytes) {
  if (h->fbuf_bytes >= bytes) return IPP_OK;
  if (h->fbuf != nullptr) cudaFree(h->fbuf);
  h->fbuf = nullptr;
  h->fbuf_bytes = 0;
  IPP_CUDA(h, cudaMalloc(&h->fbuf, bytes));
  h->fbuf_bytes = bytes;
  return IPP_OK;
}

}  // namespace

extern "C" {

const char* ipp_status_string(int status) {
  switch (status) {
    case IPP_OK: return "ok";
    case IPP_ERR_INVALID_ARG: return "invalid argument";
    case IPP_ERR_UNSUPPORTED: return "unsupported configuration";
    case IPP_ERR_CUDA: return "CUDA error (see ipp_last_error)";
    case IPP_ERR_NO_DEVICE: return "no CUDA device";
    case IPP_ERR_ALLOC: return "allocation failed";
    default: return "unknown status";
  }
}

const char* ipp_last_error(const ipp_handle* h) { return h != nullptr ? h->err : ""; }

int ipp_version(void) { return 210; }

int ipp_create(const ipp_config* cfg, ipp_handle** out) {
  if (out == nullptr) return IPP_ERR_INVALID_ARG;
  *out = nullptr;
  int rc = validate(cfg);
  if (rc != IPP_OK) return rc;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return IPP_ERR_NO_DEVICE;
  ipp_handle* h = new (std::nothrow) ipp_handle;
  if (h == nullptr) return IPP_ERR_ALLOC;
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  if (cudaGetDevice(&h->device) != cudaSuccess) {
    delete h;
    return IPP_ERR_NO_DEVICE;
  }
  const int32_t n_quads = h->cfg.map_stride >> 2;
  h->plan.quads_per_chunk = 1024;
  h->plan.n_chunks = (n_quads + h->plan.quads_per_chunk - 1) / h->plan.quads_per_chunk;
  int smem_optin = 0;
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
  h->tma = ipp::plan_tma(h->cfg, smem_optin);
  h->variant = h->tma.ok ? IPP_VARIANT_TMA : IPP_VARIANT_DIRECT;
  if (ipp::configure_plan() != cudaSuccess || (h->tma.ok && ipp::configure_step_tma(h->cfg, h->tma) != cudaSuccess)) {
    delete h;
    return IPP_ERR_CUDA;
  }
  if (const char* v = getenv("IPP_STEP_VARIANT")) {
    if (strcmp(v, "direct") == 0) h->variant = IPP_VARIANT_DIRECT;
    if (strcmp(v, "tma") == 0 && h->tma.ok) h->variant = IPP_VARIANT_TMA;
  }
  const int32_t max_chunks = h->tma.ok && h->tma.n_chunks > h->plan.n_chunks ? h->tma.n_chunks : h->plan.n_chunks;
  const size_t pb = sizeof(double) * 2 * (size_t)cfg->n_envs * max_chunks;
  const size_t gb = sizeof(int32_t) * 4 * (size_t)cfg->n_envs;
  const size_t cb = (size_t)cfg->n_envs * cfg->n_agents;
  const size_t mb = sizeof(uint32_t) * (size_t)ipp::rec_words(cfg->n_agents) * (size_t)cfg->n_envs * cfg->n_seg;
  if (cudaMalloc(&h->partials, pb) != cudaSuccess || cudaMalloc(&h->gt_params, gb) != cudaSuccess ||
      cudaMalloc(&h->comm, cb) != cudaSuccess || cudaMalloc(&h->step_meta, mb) != cudaSuccess) {
    ipp_destroy(h);
    return IPP_ERR_ALLOC;
  }
  cudaMemset(h->gt_params, 0, gb);  // split 0 until the first ipp_reset (read by the plan kernel when fix_range == 0)
  // lut[alt][byte][c]: cell c of a quad with code `byte` (low nibble: inside the footprint, high nibble:
  // seen as 1) is multiplied by k_hi / k_lo of the altitude, cells outside the footprint by k_out
  const size_t lb = sizeof(float) * 4 * 256 * (size_t)cfg->n_alt;
  {
    float* host = new (std::nothrow) float[4 * 256 * (size_t)cfg->n_alt];
    if (host == nullptr || cudaMalloc(&h->lut, lb) != cudaSuccess) {
      delete[] host;
      ipp_destroy(h);
      return IPP_ERR_ALLOC;
    }
    for (int a = 0; a < cfg->n_alt; ++a)
      for (int byte = 0; byte < 256; ++byte)
        for (int c = 0; c < 4; ++c) {
          const bool in = (byte >> c) & 1, one = (byte >> (4 + c)) & 1;
          host[((size_t)a * 256 + byte) * 4 + c] = in ? (one ? cfg->k_hi[a] : cfg->k_lo[a]) : cfg->k_out;
        }
    const cudaError_t e = cudaMemcpy(h->lut, host, lb, cudaMemcpyHostToDevice);
    delete[] host;
    if (e != cudaSuccess) {
      ipp_destroy(h);
      return IPP_ERR_CUDA;
    }
  }
  if (ipp::build_pool_tables(h->cfg, &h->pool) != cudaSuccess) {
    ipp_destroy(h);
    return IPP_ERR_ALLOC;
  }
  h->scratch_bytes = (int64_t)(pb + gb + cb + lb + mb);
  *out = h;
  return IPP_OK;
}

int ipp_destroy(ipp_handle* h) {
  if (h == nullptr) return IPP_OK;
  if (h->partials) cudaFree(h->partials);
  if (h->gt_params) cudaFree(h->gt_params);
  if (h->comm) cudaFree(h->comm);
  if (h->step_meta) cudaFree(h->step_meta);
  if (h->policy_in) cudaFree(h->policy_in);
  if (h->lut) cudaFree(h->lut);
  ipp::free_pool_tables(&h->pool);
  if (h->fbuf) cudaFree(h->fbuf);
  for (int i = 0; i < h->n_graphs; ++i) cudaGraphExecDestroy(h->graphs[i].exec);
  delete h;
  return IPP_OK;
}

int ipp_set_step_variant(ipp_handle* h, int32_t variant) {
  if (h == nullptr) return IPP_ERR_INVALID_ARG;
  if (variant == IPP_VARIANT_DIRECT) {
    h->variant = variant;
    return IPP_OK;
  }
  if (variant == IPP_VARIANT_TMA) {
    if (!h->tma.ok) return IPP_ERR_UNSUPPORTED;
    h->variant = variant;
    return IPP_OK;
  }
  return IPP_ERR_INVALID_ARG;
}

int ipp_get_step_variant(const ipp_handle* h) { return h != nullptr ? h->variant : IPP_ERR_INVALID_ARG; }

int64_t ipp_scratch_bytes(const ipp_handle* h) { return h != nullptr ? h->scratch_bytes + (int64_t)h->fbuf_bytes : 0; }

int ipp_reset(ipp_handle* h, const ipp_state* st, int32_t* pos_out, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || pos_out == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  IPP_CUDA(h, ipp::launch_reset(h->cfg, *st, h->lut, h->plan, pos_out, h->gt_params, (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_step_phases(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, int32_t phases,
                    void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || io == nullptr || io->pos_in == nullptr || io->pos_out == nullptr || io->pos_in == io->pos_out)
    return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if (t < 0 || t > 0xFFFE) return IPP_ERR_INVALID_ARG;
  ipp_step_io io2 = *io;
  if (io2.comm_out == nullptr) io2.comm_out = h->comm;
  cudaStream_t s = (cudaStream_t)stream;
  if (phases & IPP_PHASE_MOVE) IPP_CUDA(h, ipp::launch_plan(h->cfg, *st, io2, t, 1, 1, h->step_meta, h->gt_params, s));
  if (phases & IPP_PHASE_MAPS) IPP_CUDA(h, launch_maps(h, st, io2, t, true, s));
  return IPP_OK;
}

int ipp_step(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream) {
  return ipp_step_phases(h, st, t, io, IPP_PHASE_MOVE | IPP_PHASE_MAPS, stream);
}

int ipp_run_steps(ipp_handle* h, const ipp_state* st, int32_t with_reset, int32_t* reset_pos_out, int32_t t0,
                  int32_t n_steps, const ipp_step_io* ios, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || ios == nullptr || n_steps < 1 || n_steps > 4096 || (with_reset && reset_pos_out == nullptr))
    return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  // everything that is baked into the graph: the state and io pointers, the timesteps, the kernel variant
  uint64_t key = 1469598103934665603ull;
  auto mix = [&key](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) key = (key ^ b[i]) * 1099511628211ull;
  };
  mix(st, sizeof(*st));
  mix(ios, sizeof(ipp_step_io) * (size_t)n_steps);
  mix(&with_reset, sizeof(with_reset));
  mix(&reset_pos_out, sizeof(reset_pos_out));
  mix(&t0, sizeof(t0));
  mix(&n_steps, sizeof(n_steps));
  mix(&h->variant, sizeof(h->variant));
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < h->n_graphs; ++i)
    if (h->graphs[i].key == key) {
      IPP_CUDA(h, cudaGraphLaunch(h->graphs[i].exec, s));
      return IPP_OK;
    }
  // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
  cudaStream_t cs = nullptr;
  IPP_CUDA(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e == cudaSuccess) {
    if (with_reset) rc = ipp_reset(h, st, reset_pos_out, cs);
    for (int32_t i = 0; rc == IPP_OK && i < n_steps; ++i) rc = ipp_step(h, st, t0 + i, &ios[i], cs);
    e = cudaStreamEndCapture(cs, &graph);
  }
  cudaStreamDestroy(cs);
  if (rc != IPP_OK) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail_cuda(h, e, "ipp_run_steps: stream capture");
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail_cuda(h, e, "cudaGraphInstantiate");
  if (h->n_graphs == 8) {  // evict the oldest
    cudaGraphExecDestroy(h->graphs[0].exec);
    for (int i = 1; i < 8; ++i) h->graphs[i - 1] = h->graphs[i];
    h->n_graphs = 7;
  }
  h->graphs[h->n_graphs].key = key;
  h->graphs[h->n_graphs].exec = exec;
  ++h->n_graphs;
  IPP_CUDA(h, cudaGraphLaunch(exec, s));
  return IPP_OK;
}

int ipp_step_host(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, const float* probs_host,
                  const int32_t* actions_host, float* reward_rel_host, float* reward_abs_host,
                  int32_t* actions_out_host, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || io == nullptr || (probs_host == nullptr) == (actions_host == nullptr))
    return IPP_ERR_INVALID_ARG;
  if (io->reward_rel == nullptr || io->reward_abs == nullptr || io->actions_out == nullptr) return IPP_ERR_INVALID_ARG;
  const size_t n = (size_t)h->cfg.n_envs * h->cfg.n_agents;
  if (h->policy_in == nullptr) {
    IPP_CUDA(h, cudaMalloc(&h->policy_in, n * IPP_N_ACTIONS * sizeof(float)));
    h->scratch_bytes += (int64_t)(n * IPP_N_ACTIONS * sizeof(float));
  }
  cudaStream_t s = (cudaStream_t)stream;
  ipp_step_io io2 = *io;
  if (probs_host != nullptr) {
    IPP_CUDA(h, cudaMemcpyAsync(h->policy_in, probs_host, n * IPP_N_ACTIONS * sizeof(float), cudaMemcpyHostToDevice, s));
    io2.probs_in = static_cast<const float*>(h->policy_in);
    io2.actions_in = nullptr;
  } else {
    IPP_CUDA(h, cudaMemcpyAsync(h->policy_in, actions_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    io2.actions_in = static_cast<const int32_t*>(h->policy_in);
    io2.probs_in = nullptr;
  }
  int rc = ipp_step_phases(h, st, t, &io2, IPP_PHASE_MOVE | IPP_PHASE_MAPS, stream);
  if (rc != IPP_OK) return rc;
  // results: one copy when the three device outputs and the three host buffers are each contiguous in the order
  // (reward_rel, reward_abs, actions) — a device->host copy costs ~10 us of latency whatever its size
  const size_t ne = (size_t)h->cfg.n_envs;
  const bool dev_packed = io->reward_abs == io->reward_rel + ne &&
                          reinterpret_cast<const char*>(io->actions_out) == reinterpret_cast<const char*>(io->reward_abs + ne);
  const bool host_packed = reward_rel_host != nullptr && reward_abs_host == reward_rel_host + ne &&
                           reinterpret_cast<const char*>(actions_out_host) == reinterpret_cast<const char*>(reward_abs_host + ne);
  if (dev_packed && host_packed) {
    IPP_CUDA(h, cudaMemcpyAsync(reward_rel_host, io->reward_rel, 2 * ne * sizeof(float) + n * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, s));
    return IPP_OK;
  }
  if (reward_rel_host != nullptr)
    IPP_CUDA(h, cudaMemcpyAsync(reward_rel_host, io->reward_rel, ne * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (reward_abs_host != nullptr)
    IPP_CUDA(h, cudaMemcpyAsync(reward_abs_host, io->reward_abs, ne * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (actions_out_host != nullptr)
    IPP_CUDA(h, cudaMemcpyAsync(actions_out_host, io->actions_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  return IPP_OK;
}

int ipp_observe(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || io == nullptr || io->pos_in == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if (t < 0 || t > 0xFFFE) return IPP_ERR_INVALID_ARG;
  ipp_step_io io2 = *io;
  if (io2.comm_out == nullptr) io2.comm_out = h->comm;
  cudaStream_t s = (cudaStream_t)stream;
  IPP_CUDA(h, ipp::launch_plan(h->cfg, *st, io2, t, 1, 0, h->step_meta, h->gt_params, s));
  IPP_CUDA(h, launch_maps(h, st, io2, t, false, s));
  return IPP_OK;
}

int ipp_act(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || io == nullptr || io->pos_in == nullptr || io->pos_out == nullptr || io->pos_in == io->pos_out)
    return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if (t < 0 || t > 0xFFFE) return IPP_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  IPP_CUDA(h, ipp::launch_plan(h->cfg, *st, *io, t, 0, 1, h->step_meta, h->gt_params, s));
  IPP_CUDA(h, ipp::launch_own_update(h->cfg, *st, h->lut, io->pos_out, t, s));
  return IPP_OK;
}

int ipp_features_actor(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, float* obs_out,
                       void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || io == nullptr || io->pos_in == nullptr || obs_out == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if (t < 0 || t > 0xFFFE || h->cfg.px < 6 || h->cfg.py < 6) return IPP_ERR_INVALID_ARG;  // window around index 5
  const uint8_t* comm = io->comm_out != nullptr ? io->comm_out : h->comm;
  IPP_CUDA(h, ipp::launch_features_actor(h->cfg, *st, h->pool, io->pos_in, comm, t, obs_out, (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_features_critic(ipp_handle* h, const ipp_state* st, int32_t t, const int32_t* pos_in, const int32_t* actions,
                        const float* obs_in, float* state_out, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || pos_in == nullptr || actions == nullptr || obs_in == nullptr || state_out == nullptr)
    return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if (t < 0 || t > 0xFFFE) return IPP_ERR_INVALID_ARG;
  IPP_CUDA(h, ipp::launch_features_critic(h->cfg, *st, h->pool, pos_in, actions, t, obs_in, state_out,
                                          (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_export_beliefs(ipp_handle* h, const ipp_state* st, float* local_out, float* global_out, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(local_out) & 15) || (reinterpret_cast<uintptr_t>(global_out) & 15))
    return IPP_ERR_INVALID_ARG;
  const int64_t per_env = h->cfg.map_stride;
  if (local_out != nullptr)
    IPP_CUDA(h, ipp::launch_export_beliefs(st->local_maps, local_out, per_env * h->cfg.n_agents * h->cfg.n_envs,
                                           (cudaStream_t)stream));
  if (global_out != nullptr)
    IPP_CUDA(h, ipp::launch_export_beliefs(st->global_map, global_out, per_env * h->cfg.n_envs,
                                           (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_ig_plan(ipp_handle* h, const ipp_state* st, const int32_t* pos_in, int32_t communication,
                int32_t* actions_out, uint8_t* mask_out, double* gains_out, double* util_out, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || pos_in == nullptr || actions_out == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  IPP_CUDA(h, ipp::launch_ig_plan(h->cfg, *st, pos_in, communication != 0, actions_out, mask_out, gains_out, util_out,
                                  (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_eval_metrics(ipp_handle* h, const ipp_state* st, double* entropy_out, double* f1_out, void* stream) {
  DeviceGuard on_device(h);
  if (h == nullptr || entropy_out == nullptr || f1_out == nullptr) return IPP_ERR_INVALID_ARG;
  int rc = check_state(st);
  if (rc != IPP_OK) return rc;
  IPP_CUDA(h, ipp::launch_eval_metrics(h->cfg, *st, entropy_out, f1_out, (cudaStream_t)stream));
  return IPP_OK;
}

int ipp_project_fov(const ipp_handle* h, const int32_t* position, int32_t* raw, int32_t* clipped) {
  if (h == nullptr || position == nullptr || raw == nullptr || clipped == nullptr) return IPP_ERR_INVALID_ARG;
  const ipp_config& c = h->cfg;
  if (position[0] < 0 || position[1] < 0 || position[0] % c.spacing || position[1] % c.spacing ||
      position[2] % c.spacing)
    return IPP_ERR_INVALID_ARG;
  const int ix = position[0] / c.spacing, iy = position[1] / c.spacing;
  const int iz = position[2] / c.spacing - c.min_altitude / c.spacing;
  if (ix >= c.px || iy >= c.py || iz < 0 || iz >= c.n_alt) return IPP_ERR_INVALID_ARG;
  const int cx = c.cell_x[ix], cy = c.cell_y[iy], rx = c.radius_x[iz], ry = c.radius_y[iz];
  raw[0] = cy - ry;  // yu
  raw[1] = cy + ry;  // yd
  raw[2] = cx - rx;  // xl
  raw[3] = cx + rx;  // xr
  auto clip = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
  clipped[0] = clip(raw[0], c.gy - 1);
  clipped[1] = clip(raw[1], c.gy - 1);
  clipped[2] = clip(raw[2], c.gx - 1);
  clipped[3] = clip(raw[3], c.gx - 1);
  return IPP_OK;
}

int ipp_measure(ipp_handle* h, const uint8_t* gt_host, const int32_t* rect, int32_t altitude_m, uint32_t episode,
                uint32_t agent, uint32_t index, float y_hi, float y_lo, float* out_host) {
  DeviceGuard on_device(h);
  if (h == nullptr || gt_host == nullptr || rect == nullptr || out_host == nullptr) return IPP_ERR_INVALID_ARG;
  const ipp_config& c = h->cfg;
  if (rect[0] < 0 || rect[1] > c.gy || rect[2] < 0 || rect[3] > c.gx) return IPP_ERR_INVALID_ARG;
  const int64_t n = (int64_t)(rect[3] - rect[2]) * (rect[1] - rect[0]);
  if (rect[3] <= rect[2] || rect[1] <= rect[0]) return IPP_OK;  // empty footprint: nothing to write
  const int iz = altitude_m / c.spacing - c.min_altitude / c.spacing;
  // sensors/models/sensor_models.py:13-22 returns noise 0 for an unknown altitude: never measured wrongly
  const uint32_t thresh = (altitude_m % c.spacing == 0 && iz >= 0 && iz < c.n_alt) ? c.flip_thresh[iz] : 0u;
  const size_t cells = (size_t)c.gx * c.gy;
  const size_t gt_al = (cells + 15) & ~(size_t)15;
  int rc = ensure_fbuf(h, gt_al + sizeof(float) * (size_t)n);
  if (rc != IPP_OK) return rc;
  uint8_t* dgt = static_cast<uint8_t*>(h->fbuf);
  float* dout = reinterpret_cast<float*>(dgt + gt_al);
  IPP_CUDA(h, cudaMemcpy(dgt, gt_host, cells, cudaMemcpyHostToDevice));
  const uint32_t key = ipp::stream_key(c.seed, episode, agent, index, ipp::PURPOSE_NOISE);
  IPP_CUDA(h, ipp::launch_measure(c, dgt, rect, key, thresh, y_hi, y_lo, dout, 0));
  IPP_CUDA(h, cudaMemcpy(out_host, dout, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
  return IPP_OK;
}

int ipp_update_cells(ipp_handle* h, void* x_host, int32_t x_f64, const void* y_host, int32_t y_f64,
                     int32_t y_is_scalar, int64_t n, double* out_host) {
  DeviceGuard on_device(h);
  if (h == nullptr || x_host == nullptr || y_host == nullptr || out_host == nullptr || n < 0)
    return IPP_ERR_INVALID_ARG;
  if (n == 0) return IPP_OK;
  const size_t xb = (x_f64 ? 8 : 4) * (size_t)n, yb = (y_f64 ? 8 : 4) * (size_t)(y_is_scalar ? 1 : n);
  const size_t xb_al = (xb + 15) & ~(size_t)15, yb_al = (yb + 15) & ~(size_t)15;
  int rc = ensure_fbuf(h, xb_al + yb_al + 8 * (size_t)n);
  if (rc != IPP_OK) return rc;
  char* base = static_cast<char*>(h->fbuf);
  void* dx = base;
  void* dy = base + xb_al;
  double* dout = reinterpret_cast<double*>(base + xb_al + yb_al);
  IPP_CUDA(h, cudaMemcpy(dx, x_host, xb, cudaMemcpyHostToDevice));
  IPP_CUDA(h, cudaMemcpy(dy, y_host, yb, cudaMemcpyHostToDevice));
  IPP_CUDA(h, ipp::launch_update_cells(h->cfg, dx, x_f64, dy, y_f64, y_is_scalar, n, dout, 0));
  IPP_CUDA(h, cudaMemcpy(x_host, dx, xb, cudaMemcpyDeviceToHost));
  IPP_CUDA(h, cudaMemcpy(out_host, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost));
  return IPP_OK;
}

int ipp_shannon_entropy(ipp_handle* h, void* p_host, int32_t is_f64, int64_t n, void* out_host) {
  DeviceGuard on_device(h);
  if (h == nullptr || p_host == nullptr || out_host == nullptr || n < 0) return IPP_ERR_INVALID_ARG;
  if (n == 0) return IPP_OK;
  const size_t nb = (is_f64 ? 8 : 4) * (size_t)n;
  const size_t nb_al = (nb + 15) & ~(size_t)15;
  int rc = ensure_fbuf(h, 2 * nb_al);
  if (rc != IPP_OK) return rc;
  char* dp = static_cast<char*>(h->fbuf);
  char* dout = dp + nb_al;
  IPP_CUDA(h, cudaMemcpy(dp, p_host, nb, cudaMemcpyHostToDevice));
  IPP_CUDA(h, ipp::launch_entropy(dp, is_f64, n, dout, 0));
  IPP_CUDA(h, cudaMemcpy(p_host, dp, nb, cudaMemcpyDeviceToHost));
  IPP_CUDA(h, cudaMemcpy(out_host, dout, nb, cudaMemcpyDeviceToHost));
  return IPP_OK;
}

int ipp_fuse_map(ipp_handle* h, const float* own_host, const float* others_host, int32_t n_others, int64_t cells,
                 double* out_host) {
  DeviceGuard on_device(h);
  if (h == nullptr || own_host == nullptr || out_host == nullptr || cells < 0 || n_others < 0 ||
      (n_others > 0 && others_host == nullptr))
    return IPP_ERR_INVALID_ARG;
  if (cells == 0) return IPP_OK;
  const size_t nb = sizeof(float) * (size_t)cells;
  const size_t nb_al = (nb + 15) & ~(size_t)15;
  const size_t ob = nb * (size_t)n_others;
  const size_t ob_al = (ob + 15) & ~(size_t)15;
  int rc = ensure_fbuf(h, nb_al + ob_al + 8 * (size_t)cells);
  if (rc != IPP_OK) return rc;
  char* base = static_cast<char*>(h->fbuf);
  float* down = reinterpret_cast<float*>(base);
  float* doth = reinterpret_cast<float*>(base + nb_al);
  double* dout = reinterpret_cast<double*>(base + nb_al + ob_al);
  IPP_CUDA(h, cudaMemcpy(down, own_host, nb, cudaMemcpyHostToDevice));
  if (n_others > 0) IPP_CUDA(h, cudaMemcpy(doth, others_host, ob, cudaMemcpyHostToDevice));
  IPP_CUDA(h, ipp::launch_fuse_map(h->cfg, down, doth, n_others, cells, dout, 0));
  IPP_CUDA(h, cudaMemcpy(out_host, dout, 8 * (size_t)cells, cudaMemcpyDeviceToHost));
  return IPP_OK;
}

int ipp_utility_reward(ipp_handle* h, const void* last_host, int32_t last_f64, const void* next_host,
                       int32_t next_f64, int64_t cells, double* out2_host) {
  DeviceGuard on_device(h);
  if (h == nullptr || last_host == nullptr || next_host == nullptr || out2_host == nullptr || cells <= 0)
    return IPP_ERR_INVALID_ARG;
  const size_t lb = (last_f64 ? 8 : 4) * (size_t)cells, nb = (next_f64 ? 8 : 4) * (size_t)cells;
  const size_t lb_al = (lb + 15) & ~(size_t)15, nb_al = (nb + 15) & ~(size_t)15;
  const size_t rb = sizeof(double) * (2 + 2 * 296);
  int rc = ensure_fbuf(h, lb_al + nb_al + rb);
  if (rc != IPP_OK) return rc;
  char* base = static_cast<char*>(h->fbuf);
  double* dres = reinterpret_cast<double*>(base + lb_al + nb_al);
  IPP_CUDA(h, cudaMemcpy(base, last_host, lb, cudaMemcpyHostToDevice));
  IPP_CUDA(h, cudaMemcpy(base + lb_al, next_host, nb, cudaMemcpyHostToDevice));
  IPP_CUDA(h, ipp::launch_utility_reward(base, last_f64, base + lb_al, next_f64, cells, dres, 0));
  IPP_CUDA(h, cudaMemcpy(out2_host, dres, 2 * sizeof(double), cudaMemcpyDeviceToHost));
  return IPP_OK;
}

}  // extern "C"

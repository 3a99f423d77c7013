/*
 * ipp_b200.h — C ABI of the B200-native batched informative-path-planning environment.
 *
 * The reference (dmar-bonn/ipp-marl, Python/numpy) has no FFI; its "operator interface" for the
 * per-timestep environment path is the Python surface listed in SURVEY.md section 8b.  This header is
 * what a binding for that path would target (INTEGRATION.md shows the ctypes stub).  Each entry
 * point cites the reference code it replaces (paths relative to marl_framework/).
 *
 * Conventions: every function returns 0 (IPP_OK) or a negative ipp_status; nothing throws; the
 * caller owns every buffer; "device" pointers are CUDA device memory of the current device,
 * "host" pointers are ordinary host memory; `stream` is a cudaStream_t passed as void* (NULL =
 * the legacy default stream).  Distinct handles may be used from distinct threads.
 *
 * Layouts (all C-contiguous):
 *   belief maps   float32 [..., map_stride], cell (x, y) at x * gy + y  — first array axis is
 *                 world x exactly as in the reference (mapping/mappings.py:46-61); map_stride >=
 *                 gx*gy is a multiple of 4 cells so every map starts 16-byte aligned.
 *                 The resident state holds each cell's ODDS o = p / (1 - p), not p: one Bayes pass
 *                 (mapping/mappings.py:106-124) is then a clamp and a multiply per cell.  Probabilities —
 *                 what Agent.local_map holds in the reference — are produced by ipp_export_beliefs
 *                 (p = o/(1+o)); every other entry point converts on the fly where it needs p.
 *   ground truth  uint8   [n_envs, gt_stride]    (mapping/ground_truths.py:42-56 half-plane field)
 *   positions     int32   [n_envs, n_agents, 3]  metres (x, y, z), as agent/agent.py keeps them
 *   episodes      uint32  [n_envs]               episode number of each env (seeds + random streams)
 *   meas codes    uint8   [n_envs, 2, code_stride]  latest measurement of every agent, 1 byte per (quad, agent)
 */
#ifndef IPP_B200_H
#define IPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPP_MAX_AGENTS 8
#define IPP_MAX_ALT 8
#define IPP_MAX_LATTICE 128
#define IPP_N_ACTIONS 6
#define IPP_FLAG_QUADS 640 /* ipp_state.map_flags: one 32-bit word per (local map, segment of 640 quads), one bit per 32-quad tile */

typedef enum ipp_status {
  IPP_OK = 0,
  IPP_ERR_INVALID_ARG = -1,
  IPP_ERR_UNSUPPORTED = -2,
  IPP_ERR_CUDA = -3,
  IPP_ERR_NO_DEVICE = -4,
  IPP_ERR_ALLOC = -5
} ipp_status;

/*
 * Static description of one batch of environments.  The geometric / arithmetic tables are
 * computed ON THE HOST with the reference's own float64 expressions (ipp_marl_b200/geometry.py)
 * because they sit on floating-point knife edges (SURVEY.md section 7 "exact geometry").
 */
typedef struct ipp_config {
  int32_t gx, gy;          /* belief cells per side: mapping/grid_maps.py:16-50                    */
  int32_t map_stride;      /* cells between consecutive float32 maps, multiple of 4, >= gx*gy     */
  int32_t gt_stride;       /* bytes between consecutive ground-truth maps, multiple of 16, >=      */
                           /* map_stride (16-byte rows so that TMA bulk copies can stage them)     */
  int32_t code_stride;     /* bytes of measurement codes per env: roundup16(ceil(gx*gy/4) * (A<=4 ? 4 : 8)) */
  int32_t n_seg;           /* flag segments per map: ceil(ceil(gx*gy/4) / IPP_FLAG_QUADS)                  */
  int32_t px, py, n_alt;   /* agent lattice: agent/state_space.py:16-18                            */
  int32_t n_agents;        /* 1..IPP_MAX_AGENTS                                                    */
  int32_t n_envs;          /* envs owned by this handle (this GPU's shard)                         */
  int32_t spacing;         /* metres between lattice points                                        */
  int32_t min_altitude, max_altitude, x_dim_m, y_dim_m;
  int32_t budget;          /* last timestep index: coma_wrapper.py:163-164                         */
  uint32_t seed;           /* environment.seed                                                     */
  int32_t comm_d2_max;     /* largest squared distance (m^2) with sqrt(d2) <= communication_range  */
  uint32_t fail_thresh24;  /* message delivered iff (hash >> 8) >= fail_thresh24                   */
  float prior;             /* mapping.prior                                                        */
  float k_out;             /* odds multiplier of a 0.5 cell of map2communicate = exp(-logit prior) */
  float p_min, p_max;      /* float32(0.0001), float32(0.9999): mapping/mappings.py:110-111        */
  float o_min, o_max;      /* same clamp in odds space                                             */
  int32_t radius_x[IPP_MAX_ALT], radius_y[IPP_MAX_ALT]; /* sensors/cameras.py:62-67 per altitude   */
  float k_hi[IPP_MAX_ALT], k_lo[IPP_MAX_ALT]; /* odds multipliers of a cell seen as 1 / as 0       */
  float y_hi[IPP_MAX_ALT], y_lo[IPP_MAX_ALT]; /* measurement values themselves (simulations.py:47-50): */
                                              /* float32(round(1-noise, 3)) and float32(round(noise, 3)) */
  uint32_t flip_thresh[IPP_MAX_ALT];          /* cell measured wrongly iff hash < flip_thresh      */
  int32_t cell_x[IPP_MAX_LATTICE], cell_y[IPP_MAX_LATTICE]; /* floor(pos/res_x): cameras.py:66     */
  int32_t fix_range;       /* experiment.uav.fix_range; 0 = per-episode random range (communication_log.py:22-31): */
  int32_t comm_d2_table[4];/* comm_d2_max of the ranges {0, 15, 25, 100} m, indexed by the episode's first         */
                           /* randint(4) after np.random.seed(episode) (drawn by ipp_reset: call it on this handle) */
  int32_t n_meas;          /* measurement values the simulation can produce (mapping/simulations.py:47-50) and their   */
  float meas_y[16];        /* float32 logits AS NUMPY EVALUATES THEM (np.log(y / (1 - y)) on float32, mappings.py:113):  */
  float meas_ly[16];       /* y_hi / y_lo of every altitude and 0.5; used by the single-map entry points               */
  double l_prior;          /* np.log(prior / (1 - prior)) in float64: mapping/mappings.py:116 (single-map entry points) */
} ipp_config;

/* Device-resident state of the batch (struct-of-arrays form of agent/agent.py:13-38 per UAV and of
 * the accumulated global map of missions/episode_generator.py:47,53). */
typedef struct ipp_state {
  float* local_maps;   /* [n_envs, n_agents, map_stride]  Agent.local_map, as odds                 */
  float* global_map;   /* [n_envs, map_stride]            accumulated_map_knowledge, as odds       */
  uint8_t* ground_truth; /* [n_envs, gt_stride]           Mapping.simulated_map                    */
  uint32_t* episodes;  /* [n_envs]                                                                  */
  uint8_t* meas_codes; /* [n_envs, 2, code_stride] compact form of Agent.map2communicate: one byte per   */
                       /* (4-cell quad, agent): low nibble = cell inside the agent's latest footprint,   */
                       /* high nibble = cell measured as occupied.  Row (t & 1) of an env holds the measurements */
                       /* communicated at step t, the other half receives those taken after the moves.   */
  uint32_t* map_flags; /* [n_envs, n_seg, 8] (word i of a 32-byte record = local map i; 16-byte aligned):           */
                       /* bookkeeping of the reference's lazily applied clamp (mapping/mappings.py:110-111 clamps a */
                       /* map only when the next update reads it).  Bit t of a word: tile t (32 quads = 128 cells)  */
                       /* of this 640-quad segment of the local map may hold odds outside [o_min, o_max], so the    */
                       /* next fuse pass must clamp all of that tile; a clear bit lets the kernels touch only the   */
                       /* footprint cells of the tile.  Conservative (a set bit is always safe); written by reset / */
                       /* step / act.  A caller that writes odds into local_maps itself must set the words of the  */
                       /* maps it touched to 0xFFFFFFFF.                                                            */
} ipp_state;

/* Per-step inputs / outputs (device pointers; any output may be NULL). */
typedef struct ipp_step_io {
  const int32_t* pos_in;    /* [n_envs, n_agents, 3] positions at the start of the step            */
  int32_t* pos_out;         /* [n_envs, n_agents, 3] positions after the moves                     */
  const int32_t* actions_in;/* [n_envs, n_agents] injected actions (-1 = stay) or NULL             */
  const float* probs_in;    /* [n_envs, n_agents, 6] policy probabilities or NULL; with both NULL  */
                            /* the policy is uniform over the unmasked actions                     */
  int32_t greedy;           /* probs_in: 0 = sample like torch.multinomial(p*mask), 1 = argmax     */
  int32_t* actions_out;     /* [n_envs, n_agents] chosen actions                                   */
  uint8_t* mask_out;        /* [n_envs, n_agents] bit a = action a allowed (after collision mask)  */
  uint8_t* comm_out;        /* [n_envs, n_agents] bit j = agent i received agent j's message       */
  float* reward_rel;        /* [n_envs] 22*rel-0.5 : utils/reward.py:39-41                         */
  float* reward_abs;        /* [n_envs] 10*abs-0.17: utils/reward.py:38                            */
  uint8_t* stuck_out;       /* [n_envs] bit0: some agent had an all-zero mask and stayed in place;  */
                            /*          bit1: an injected action would have left the lattice (-> stay) */
} ipp_step_io;

typedef struct ipp_handle ipp_handle;

const char* ipp_status_string(int status);
/* Text of the last CUDA error seen by this handle ("" if none). */
const char* ipp_last_error(const ipp_handle* h);
int ipp_version(void);

/* Validate the config, pick launch geometry, allocate the (small) per-handle scratch. */
int ipp_create(const ipp_config* cfg, ipp_handle** out);
int ipp_destroy(ipp_handle* h);
/* Which implementation of the map kernel ipp_step / ipp_observe launch.  Default: the TMA-staged
 * persistent kernel when the (G, A) shape fits its shared-memory ring, else the direct-load kernel;
 * the environment variable IPP_STEP_VARIANT=direct|tma overrides the default at ipp_create.
 * Both produce bit-identical belief maps. */
#define IPP_VARIANT_DIRECT 0
#define IPP_VARIANT_TMA 1
int ipp_set_step_variant(ipp_handle* h, int32_t variant);
int ipp_get_step_variant(const ipp_handle* h);

/* Bytes of device scratch held by the handle (reward partials, reset parameters, the per-step item records that the
 * plan kernel hands to the map kernel: 144 bytes per env and map segment at 4 agents). */
int64_t ipp_scratch_bytes(const ipp_handle* h);

/*
 * Episode reset for the whole batch.  Replaces, per env: Mapping.__init__ -> Simulation ->
 * gaussian_random_field (mapping/ground_truths.py:42-56: MT19937 seeded with the episode number),
 * Mapping.init_priors (mapping/mappings.py:126-132), AgentStateSpace.get_random_agent_state
 * (agent/state_space.py:28-51: MT19937 seeded with seed*episode*agent) and the t == 0 measurement
 * of Agent.communicate (agent/agent.py:44-49).  `pos_out` receives the start positions.
 */
int ipp_reset(ipp_handle* h, const ipp_state* st, int32_t* pos_out, void* stream);

/*
 * One full environment timestep, fused (policy = injected actions, given probabilities or uniform):
 * CommunicationLog.get_messages (agent/communication_log.py:39-58) -> Agent.receive_messages /
 * Mapping.fuse_map "local" (agent/agent.py:62-71, mapping/mappings.py:82-89) -> fuse_map "global"
 * (coma_wrapper.py:93-95) -> get_global_reward (utils/reward.py:11-53) -> per agent, in id order:
 * get_action_mask / apply_collision_mask / action choice / action_to_position
 * (agent/action_space.py:56-70,211-223,328-344) -> Mapping.update_grid_map at the new position
 * (mapping/mappings.py:32-78).  t is the timestep index (0..budget).
 */
int ipp_step(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream);

/*
 * n_steps consecutive timesteps t0 .. t0 + n_steps - 1 — optionally preceded by ipp_reset (with_reset != 0, start
 * positions to reset_pos_out) — as ONE launch: exactly the kernels that the same sequence of ipp_reset / ipp_step
 * calls would launch, captured once into a CUDA graph that the handle keeps and replays on every later call with the
 * same arguments (same state buffers, same ios[i] contents, same t0 / n_steps / with_reset).  ios [n_steps] host
 * array; for a whole episode pass its budget + 1 structs, ios[i].pos_in = ios[i - 1].pos_out.  The policy inputs of
 * ios (actions_in / probs_in) are device pointers like in ipp_step and are read when the graph runs; results of
 * every step go where its ios[i] points, so give each step its own reward / action rows to keep them all.
 * Saves the per-launch host cost and the gaps between the 2 .. 3 kernels of a step (missions/episode_generator.py:
 * 49-79 is one such loop over the budget).
 */
int ipp_run_steps(ipp_handle* h, const ipp_state* st, int32_t with_reset, int32_t* reset_pos_out, int32_t t0,
                  int32_t n_steps, const ipp_step_io* ios, void* stream);

/*
 * ipp_step for a policy that lives on the HOST: copies the step's policy output from host memory to the device,
 * runs the timestep and copies its results back, all asynchronously on `stream` (one call, no host work between the
 * copies and the two launches).  probs_host [n_envs, n_agents, 6] float32 or actions_host [n_envs, n_agents] int32
 * (exactly one non-NULL; page-locked memory keeps the copies asynchronous); io->probs_in / io->actions_in are
 * ignored, io->reward_rel / reward_abs / actions_out must be device buffers.  reward_rel_host / reward_abs_host
 * [n_envs] float32 and actions_out_host [n_envs, n_agents] int32 may be NULL.  When the three device outputs and
 * the three host buffers are each one contiguous block in the order (reward_rel, reward_abs, actions) the results
 * travel in a single copy.  The host outputs are valid once the stream has been synchronised.
 */
int ipp_step_host(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, const float* probs_host,
                  const int32_t* actions_host, float* reward_rel_host, float* reward_abs_host,
                  int32_t* actions_out_host, void* stream);

/* ipp_step with its two launches selectable (profiling aid: lets bench.py bracket the map kernel
 * alone with CUDA events).  phases = IPP_PHASE_MOVE | IPP_PHASE_MAPS is exactly ipp_step; the MAPS
 * phase needs comm_out / pos_out of a preceding MOVE phase of the same timestep. */
#define IPP_PHASE_MOVE 1
#define IPP_PHASE_MAPS 2
int ipp_step_phases(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, int32_t phases,
                    void* stream);

/* The same timestep split around the policy network: ipp_observe = everything before the actor
 * forward (fuse local + global + reward; io->pos_in only), ipp_act = masks, action choice, move
 * and the measurement update at the new positions. */
int ipp_observe(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream);
int ipp_act(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, void* stream);

/*
 * Network-input feature builders (split mode only: call between ipp_observe and
 * ipp_act / after ipp_act of the same timestep t).
 * ipp_features_actor : actor/transformations.py:14-59 -> obs_out [n_envs, n_agents, px, py, 7] float32
 *   (budget, agent id, ego-centred position map, w-entropy of the area-pooled fused local map, w-entropy of
 *   the pooled footprint image, pooled fused local map, pooled footprint-ownership map); needs io->pos_in
 *   and the comm bits written by ipp_observe (io->comm_out, or NULL to use the handle's copy).
 * ipp_features_critic: critic/transformations.py:17-67 -> state_out [n_envs, n_agents, px, py, 12] float32
 *   = obs (7) + global position map, w-entropy / area-pooled global map, union footprint, other agents'
 *   action map; needs pos_in (pre-move positions), actions [n_envs, n_agents] and obs_in.
 * Area pooling follows cv2.resize(INTER_AREA) (utils/state.py:22-41) with host-built tap tables.
 */
int ipp_features_actor(ipp_handle* h, const ipp_state* st, int32_t t, const ipp_step_io* io, float* obs_out,
                       void* stream);
int ipp_features_critic(ipp_handle* h, const ipp_state* st, int32_t t, const int32_t* pos_in, const int32_t* actions,
                        const float* obs_in, float* state_out, void* stream);

/*
 * The belief maps as probabilities (what Agent.local_map / accumulated_map_knowledge hold in the reference:
 * agent/agent.py:35, missions/episode_generator.py:47).  local_out [n_envs, n_agents, map_stride] and
 * global_out [n_envs, map_stride] float32 device buffers; either may be NULL.  p = o/(1+o) for o < 1, else
 * 1 - 1/(1+o), in IEEE float32 (keeps the accuracy of 1-p near p = 1).
 */
int ipp_export_beliefs(ipp_handle* h, const ipp_state* st, float* local_out, float* global_out, void* stream);

/*
 * Batched information-gain greedy planner (IG_baseline.py:127-135,222-325), split mode:
 * call after ipp_observe of timestep t, then hand actions_out to ipp_act as injected actions.  Per env and
 * agent: action mask (bounds + collision rules against the CURRENT positions of the lower-id agents), expected
 * entropy reduction of the agent's fused local map over the footprint of every allowed candidate position
 * (get_individual_ig), per-agent normalisation (get_relative_ig), with `communication` != 0 the sequential
 * discount of candidates other agents can reach too (get_cell_utilities), argmax (select_action).
 * pos_in [n_envs, n_agents, 3]; actions_out [n_envs, n_agents] int32; optional (may be NULL): mask_out
 * [n_envs, n_agents] uint8, gains_out / util_out [n_envs, n_agents, 6] float64 (0 for masked actions).
 */
int ipp_ig_plan(ipp_handle* h, const ipp_state* st, const int32_t* pos_in, int32_t communication,
                int32_t* actions_out, uint8_t* mask_out, double* gains_out, double* util_out, void* stream);

/*
 * Evaluation metrics of the accumulated global map (IG_baseline.py:81-100,191-210):
 * entropy_out [n_envs] float64 = mean Shannon entropy over the ground-truth-occupied cells
 * (utils/state.py:53-121 "eval" branch), f1_out [n_envs] float64 = F1 score of class 1 of the thresholded map
 * (utils/utils.py:43-76, `get_wrmse`).
 */
int ipp_eval_metrics(ipp_handle* h, const ipp_state* st, double* entropy_out, double* f1_out, void* stream);

/* ---- single-map entry points used by the drop-in facade (host pointers, synchronous) -------- */

/* Camera.project_field_of_view (sensors/cameras.py:46-79): position[3] metres ->
 * raw[4], clipped[4] = [yu, yd, xl, xr].  Pure host arithmetic on the config tables. */
int ipp_project_fov(const ipp_handle* h, const int32_t* position, int32_t* raw, int32_t* clipped);

/* Simulation.get_measurement (mapping/simulations.py:42-65): gt_host [gx*gy] uint8, rect = clipped
 * footprint [yu, yd, xl, xr], noise stream (episode, agent, index); writes the [xr-xl, yd-yu] block of
 * measurement values (y_hi where the cell is seen as 1, y_lo otherwise) to out_host. */
int ipp_measure(ipp_handle* h, const uint8_t* gt_host, const int32_t* rect, int32_t altitude_m, uint32_t episode,
                uint32_t agent, uint32_t index, float y_hi, float y_lo, float* out_host);

/*
 * The four entry points below reproduce the reference's dtype flow under numpy >= 2: the
 * `*_f64` flags say whether a host array holds float32 (0) or float64 (1) values, exactly as the numpy arrays the
 * reference would pass; results have the dtype the reference returns.
 *
 * Mapping.update_cells / apply_update (mapping/mappings.py:106-124) on n cells: x is clamped IN PLACE (in its own
 * dtype) like the reference; y is per cell (y_is_scalar == 0) or one value (IG_baseline.py:240-245 passes a Python
 * float = float64 scalar); logit(x) and logit(y) are taken in their own dtypes, the sigmoid in float64;
 * out_host [n] float64 = updated probabilities.
 */
int ipp_update_cells(ipp_handle* h, void* x_host, int32_t x_f64, const void* y_host, int32_t y_f64,
                     int32_t y_is_scalar, int64_t n, double* out_host);

/* get_shannon_entropy (utils/state.py:118-121): p clamped IN PLACE, H (same dtype as p) written to out. */
int ipp_shannon_entropy(ipp_handle* h, void* p_host, int32_t is_f64, int64_t n, void* out_host);

/* Mapping.fuse_map (mapping/mappings.py:80-104): own [cells] (float32: the reference casts it with np.float32(...)
 * on entry) fused with n_others dense float32 maps (map2communicate arrays, [n_others, cells]) in order; the first
 * pass takes its logits in float32, every later pass runs in float64; result float64 in out (own is not modified). */
int ipp_fuse_map(ipp_handle* h, const float* own_host, const float* others_host, int32_t n_others, int64_t cells,
                 double* out_host);

/* get_utility_reward (utils/reward.py:68-82) on two dense maps: out[0] = absolute, out[1] = relative. */
int ipp_utility_reward(ipp_handle* h, const void* last_host, int32_t last_f64, const void* next_host,
                       int32_t next_f64, int64_t cells, double* out2_host);

#ifdef __cplusplus
}
#endif
#endif /* IPP_B200_H */

/*
 * dcc_b200.h — C ABI of the B200-native dynamic-coverage-control hot path.
 *
 * Drop-in boundary: the reference's Python plugin boundary for this path is the vec-env object
 * returned by make_env(cfg) (envs/make_env.py:8-49) — reset()/step(actions) over E env instances,
 * implemented there as one OS process per env behind pickled Pipes (envs/wrappers.py:97-165,
 * 203-235) around DCEnv (envs/mpe/uav_dcc.py:7-58) -> MultiAgentEnv.step (envs/mpe/multiagent/
 * environment.py:86-110) -> CoverageWorld.step (envs/mpe/multiagent/CoverageWorld.py:57-68) and
 * the Scenario callbacks (envs/mpe/multiagent/scenarios/coverage.py:64-117).  This library replaces
 * everything below that boundary with one CUDA launch over all E instances.
 * (All reference paths are relative to /root/reference/uav_dcc_control/.)
 *
 * Conventions
 *   - every function returns 0 (DCC_OK) or a negative dcc_status; nothing throws; nothing calls
 *     cudaDeviceSynchronize; launches go to the caller's stream (a CUstream / cudaStream_t passed
 *     as void*; NULL = the legacy default stream);
 *   - `d_` pointers are device memory OWNED BY THE CALLER (torch tensors' data_ptr()); `h_` pointers
 *     are host memory.  The library owns only what hangs off its handle (compact env state:
 *     UAV position/velocity in float64, PoI energies in uint8, the PoI table);
 *   - one handle = one GPU; a handle may be used from one host thread at a time;
 *   - no torch types, no C++ types: plain C, loadable with ctypes / cgo / JNI / N-API.
 */
#ifndef DCC_B200_H_
#define DCC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCC_ABI_VERSION 7
#define DCC_MAX_AGENTS 32 /* one warp lane per UAV */

typedef enum dcc_status {
    DCC_OK = 0,
    DCC_ERR_INVALID_ARG = -1,   /* NULL pointer, bad shape, unsupported configuration */
    DCC_ERR_CUDA = -2,          /* a CUDA runtime call failed; see dcc_last_cuda_error() */
    DCC_ERR_NO_DEVICE = -3,     /* no CUDA device / wrong architecture (needs sm_100) */
    DCC_ERR_ALLOC = -4,
    DCC_ERR_UNSUPPORTED = -5
} dcc_status;

typedef void *dcc_stream_t; /* cudaStream_t */

/*
 * Environment configuration.  Mirrors config/env_config/dcc.yaml + the scenario/world constants
 * (SURVEY.md Appendix A.1).  Fill with dcc_env_cfg_default() and override.
 *
 * reference_compat = 1 reproduces the SHIPPED reference, whose scenario never forwards its comm
 * arguments to the world (scenarios/coverage.py:34 builds CoverageWorld() with defaults):
 * the world then uses comm_r_scale 0.9 and contact force 0 regardless of the two fields below.
 * reference_compat = 0 passes comm_r_scale / comm_force_scale through
 * (world.contact_force = 1e2 * comm_force_scale, core.py:109 + CoverageWorld.py:16).
 */
typedef struct dcc_env_cfg {
    int32_t n_envs;            /* E: independent env instances (reference: n_rollout_threads) */
    int32_t n_agents;          /* N: UAVs, 1..32 */
    int32_t n_pois;            /* M: points of interest, 1..4096 */
    int32_t max_ep_len;        /* informational (the rollout loop enforces it, learner.py:184) */
    int32_t reference_compat;  /* see above */
    int32_t reserved0;
    double r_cover;            /* dcc.yaml:8 */
    double r_comm;             /* dcc.yaml:9 */
    double comm_r_scale;       /* dcc.yaml:10 */
    double comm_force_scale;   /* dcc.yaml:11 */
    double dt;                 /* 0.1   CoverageWorld.py:23 */
    double damping;            /* 0.25  core.py:107 */
    double max_speed;          /* 0.5   scenarios/coverage.py:53 */
    double sensitivity;        /* 5.0   environment.py:187 */
    double m_energy;           /* 5.0   scenarios/coverage.py:23 */
    double rew_cover;          /* 75    scenarios/coverage.py:25 */
    double rew_done;           /* 1500  scenarios/coverage.py:26 */
    double rew_out;            /* -100  scenarios/coverage.py:28 */
    double contact_margin;     /* 1e-3  core.py:110 */
} dcc_env_cfg;

/* Fills *cfg with the shipped defaults (4 UAV / 20 PoI / 16 envs, dcc.yaml). */
int dcc_env_cfg_default(dcc_env_cfg *cfg);

/* Observation row length D = 4 + 2(N-1) + 5M (scenarios/coverage.py:99-110). */
int dcc_env_obs_dim(int32_t n_agents, int32_t n_pois);

/*
 * Replaces: make_env(cfg) -> SubprocVecEnv.__init__ / DummyVecEnv.__init__ (envs/make_env.py:46-49,
 * envs/wrappers.py:134-154,204-212) -> DCEnv.__init__ (envs/mpe/uav_dcc.py:8-44) ->
 * Scenario.make_world (scenarios/coverage.py:33-62).
 * h_poi_xy: M x 2 float64 PoI positions (the reference loads scenarios/pos_pois.npy[0:M],
 * scenarios/coverage.py:15-17), shared by all E instances.  State starts as after reset.
 */
int dcc_env_create(const dcc_env_cfg *cfg, const double *h_poi_xy, int device, void **handle);

/* Replaces: ShareVecEnv.close / SubprocVecEnv.close (envs/wrappers.py:56-63,187-197). */
/*
 * Per-env PoI layouts ("synthetic PoI layouts", BASELINE north_star; SURVEY.md §8b / §8d stress variant).  The reference
 * gives every env instance the same table, scenarios/pos_pois.npy[0:M] (scenarios/coverage.py:15-17, with
 * `np.random.uniform(-1, 1)` per landmark as its commented-out alternative, :71); its SubprocVecEnv workers are
 * nevertheless independent Scenario objects, so distinct `pos_pois` per env is the natural generalisation.
 *   h_poi_xy  host, [n_envs, M, 2] float64 (copied; the call synchronises the stream), or NULL to go back to the
 *             shared layout passed to dcc_env_create.  Takes effect from the next reset / step; the PoI energies are
 *             not touched (call dcc_env_reset for a fresh episode on the new layouts).
 */
int dcc_env_set_poi_layouts(void *handle, const double *h_poi_xy, dcc_stream_t stream);
int dcc_env_destroy(void *handle);

/*
 * Replaces: SubprocVecEnv.reset / DummyVecEnv.reset (envs/wrappers.py:167-171,237-239) ->
 * MultiAgentEnv.reset (environment.py:112-123) -> Scenario.reset_world (scenarios/coverage.py:64-78).
 * d_obs: E x N x D float32 (may be NULL: state reset only).
 */
int dcc_env_reset(void *handle, float *d_obs, dcc_stream_t stream);

/*
 * Replaces: SubprocVecEnv.step_async + step_wait + worker('step') / DummyVecEnv.step
 * (envs/wrappers.py:102-110,156-165,214-235) -> DCEnv.step (uav_dcc.py:46-49) ->
 * MultiAgentEnv.step (environment.py:86-110) -> CoverageWorld.step (CoverageWorld.py:57-68) and
 * Scenario.observation/reward/done (scenarios/coverage.py:80-117), including the wrapper's
 * auto-reset (obs replaced by the reset obs; reward/done/info are the terminal step's).
 *
 *   d_actions   E x N x 2 float32, NOT mutated (the reference scales the caller's array by 5 in place,
 *               environment.py:186-190)
 *   d_obs       E x N x D float32 = float32(reference float64 obs); share_obs is the same memory
 *               viewed as E x (N*D) (learner.py:219-220)
 *   d_rew       E x N float32, the N entries of an env are equal (environment.py:106-108)       [opt]
 *   d_done      E x N uint8,   the N entries of an env are equal (scenarios/coverage.py:112-117) [opt]
 *   d_coverage  E float32: info["coverage_rate"] of the (possibly terminal) step (uav_dcc.py:48)  [opt]
 *   d_connect   E uint8: bit0 = world.connect, bit1 = world.connect_ (CoverageWorld.py:92-93)    [opt]
 *   d_adj, d_adj_s  E x N uint32 row bitmasks of adj_mat / adj_mat_ (CoverageWorld.py:72-83)    [opt]
 * [opt] pointers may be NULL.
 */
int dcc_env_step(void *handle, const float *d_actions, float *d_obs, float *d_rew, uint8_t *d_done,
                 float *d_coverage, uint8_t *d_connect, uint32_t *d_adj, uint32_t *d_adj_s,
                 dcc_stream_t stream);

/*
 * Same call with HOST buffers (the shape of the reference's numpy interface): copies h_actions to the
 * device, steps, copies the results back and synchronises the stream before returning.  Pinned host
 * memory (dcc_host_alloc) makes the copies asynchronous DMA.  Optional outputs may be NULL.
 */
int dcc_env_step_host(void *handle, const float *h_actions, float *h_obs, float *h_rew, uint8_t *h_done,
                      float *h_coverage, dcc_stream_t stream);
int dcc_env_reset_host(void *handle, float *h_obs, dcc_stream_t stream);

/*
 * Parity / checkpoint access to the compact state (no reference equivalent: the reference state
 * lives in Python objects, agent.state.p_pos/p_vel and landmark.energy/done).
 *   pos_vel: E x N x 4 float64 (px, py, vx, vy);  energy: E x M uint8 (done_j == energy_j >= m_energy)
 * Host pointers; the call synchronises the stream.
 */
int dcc_env_get_state(void *handle, double *h_pos_vel, uint8_t *h_energy, dcc_stream_t stream);
int dcc_env_set_state(void *handle, const double *h_pos_vel, const uint8_t *h_energy, dcc_stream_t stream);

/* Device pointers to the live compact state (for the rollout storage and for in-place checkpoints). */
int dcc_env_state_ptrs(void *handle, double **d_pos_vel, uint8_t **d_energy);

/* Device-to-device copy of the live compact state into caller-owned rollout storage (slot t+1 of the compact rollout,
 * see "compact-state learner path" below): d_pos_vel_out [E, N, 4] float64, d_energy_out [E, M] uint8.  Asynchronous on
 * `stream`.  Replaces the obs / share_obs copies of SharedReplayBuffer.insert (buffer/shared_buffer.py:72-80). */
int dcc_env_snapshot_state(void *handle, double *d_pos_vel_out, uint8_t *d_energy_out, dcc_stream_t stream);

/* Launch geometry knobs (tuning / tests).  warps_per_cta in {1,2,4,8,16}; ctas <= 0 = auto. */
int dcc_env_set_launch(void *handle, int warps_per_cta, int ctas);
/* (N, M) shapes with a compile-time specialised kernel (4/20, 8/64, 16/256) use it by default; enable = 0 forces the
 * generic runtime-shape kernel (tests compare the two).  Returns 1 if the specialised kernel is in use, 0 if not. */
int dcc_env_use_specialized(void *handle, int enable);
/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches). */
int64_t dcc_env_launch_count(void *handle);

/* Pinned host memory helpers for the *_host entry points. */
int dcc_host_alloc(void **ptr, size_t bytes);
int dcc_host_free(void *ptr);


/* =====================================================================================================
 * MAPPO learner path.
 *
 * Replaces, below the reference's algo boundary (algos/mappo.py: MAPPOPolicy.get_actions / get_values /
 * evaluate_actions, MAPPOTrainer.train / ppo_update / cal_value_loss; buffer/shared_buffer.py:
 * compute_returns; utils/valuenorm.py), the torch-CPU networks, the NumPy GAE loop and the autograd graph
 * with CUDA kernels.  The Python host (dcc_b200/algos, dcc_b200/buffer) keeps the reference's class and
 * method names.
 *
 * Memory conventions
 *   - parameters, gradients and Adam moments of each net are CALLER-owned flat float32 device buffers in the
 *     reference's state_dict order without the never-used fc_h block (mlp.py:21-23):
 *       base.feature_norm.weight[in] .bias[in]  base.mlp.fc1.0.weight[H,in] .bias[H]  base.mlp.fc1.2.weight[H] .bias[H]
 *       base.mlp.fc2.0.0.weight[H,H] .bias[H]  base.mlp.fc2.0.2.weight[H] .bias[H]
 *       head.weight[out,H] head.bias[out]      (actor: act.action_out.fc_mean, out = 2; critic: v_out, out = 1)
 *       actor only: act.action_out.logstd._bias[2]
 *     (in = D for the actor, N*D for the critic);
 *   - the rollout is stored per ENV-STEP row (t,e): the N agent rows of an env share reward, mask, value and
 *     return (shared reward environment.py:106-108; identical critic input learner.py:219-220), so those
 *     arrays are [T(+1), E]; obs is [T+1, E, N, D] and doubles as the critic input [T+1, E, N*D];
 *   - ValueNorm state = 3 float32 {running_mean, running_mean_sq, debiasing_term} (valuenorm.py:24-26);
 *   - the handle owns only activation scratch for `chunk_rows` env-step rows; bigger batches are streamed in
 *     chunks with gradient accumulation == the reference's single minibatch (num_mini_batch 1, mappo.yaml:43).
 * ===================================================================================================== */
typedef struct dcc_mappo_cfg {
    int32_t n_agents;        /* N */
    int32_t obs_dim;         /* D (actor input); the critic input is N*D */
    int32_t hidden;          /* algo_hidden_size, <= 256 (mappo.yaml: 256) */
    int32_t act_dim;         /* 2 (Box(2), environment.py:52) */
    int32_t chunk_rows;      /* env-step rows per activation chunk; 0 = auto (~2.5 GB of scratch, whole 148-SM waves for both nets) */
    int32_t gemm_backend;    /* 0 = auto, 1 = SIMT fp32 FFMA, 2 = tcgen05 split precision: 3xTF32, fp16 hi/lo on LayerNorm outputs (hidden == 256 only) */
    float clip_param;        /* 0.2     mappo.yaml */
    float entropy_coef;      /* 0.01 */
    float value_loss_coef;   /* 1.0 */
    float huber_delta;       /* 10.0 */
    float max_grad_norm;     /* 10.0 */
    float gamma;             /* 0.99 */
    float gae_lambda;        /* 0.95 */
    float opti_eps;          /* 1e-5    Adam eps */
    float adam_beta1;        /* 0.9 */
    float adam_beta2;        /* 0.999 */
    double vn_beta;          /* 0.99999 valuenorm.py:11 (double: torch forms 1.0 - beta in double before the float32 op) */
    /* mappo.yaml switches of the update path (0 / 1; dcc_mappo_cfg_default sets the shipped values) */
    int32_t use_huber_loss;          /* 1: one-sided Huber (util.py:36-39); 0: mse_loss = e^2/2 (util.py:42-43)   mappo.py:113-118 */
    int32_t use_clipped_value_loss;  /* 1: max(original, clipped); 0: original only                               mappo.py:120-123 */
    int32_t use_max_grad_norm;       /* 1: clip_grad_norm_(max_grad_norm); 0: norm reported, gradients unscaled   mappo.py:176-181 */
    int32_t use_valuenorm;           /* 1: ValueNorm on returns; 0: value_normalizer = None (raw returns)         mappo.py:96-101 */
    int32_t use_gae;                 /* 1: GAE; 0: discounted returns bootstrapped from the RAW next value        shared_buffer.py:199-212 */
    int32_t use_feature_normalization; /* 1: LayerNorm on the raw input (mlp.py:44-45,52-53); 0: the flat parameter buffers have no
                                          feature_norm.weight / .bias entries and fc1 sees the raw observation */
    float weight_decay;              /* 0: torch.optim.Adam L2 term, grad += weight_decay * param (after the clip) mappo.py:30-37 */
    int32_t use_relu;                /* 1: ReLU trunk; 0: tanh (mlp.py:13 `[nn.Tanh(), nn.ReLU()][use_ReLU]`) */
    int32_t layer_N;                 /* 1..3: number of fc2 blocks after fc1 (mlp.py:23,27-28); the flat buffers then hold
                                        base.mlp.fc2.{i}.0.weight/.bias, .2.weight/.bias for i < layer_N, in order */
    int32_t recurrent_N;             /* 0: MLP policy (shipped).  1..4: use_recurrent_policy / use_naive_recurrent_policy — an RNNLayer
                                        (torch.nn.GRU x recurrent_N + LayerNorm, algos/algo_utils/rnn.py:8-22) between the trunk and the
                                        head of both nets (r_actor_critic.py:36-37,102-103); the flat buffers then hold, after the last
                                        fc2 block: rnn.rnn.weight_ih_l{i} [3H, H], weight_hh_l{i} [3H, H], bias_ih_l{i} [3H],
                                        bias_hh_l{i} [3H] for i < recurrent_N, then rnn.norm.weight / .bias [H] */
} dcc_mappo_cfg;

int dcc_mappo_cfg_default(dcc_mappo_cfg *cfg);
int dcc_mappo_create(const dcc_mappo_cfg *cfg, int device, void **handle);
int dcc_mappo_destroy(void *handle);
/* Flat parameter count of a net: which = 0 actor, 1 critic. */
int64_t dcc_mappo_param_count(void *handle, int which);
int dcc_mappo_chunk_rows(void *handle);
/* GEMM backend actually in use: 1 = SIMT fp32, 2 = tcgen05 3xTF32. */
int dcc_mappo_gemm_backend(void *handle);
int64_t dcc_mappo_launch_count(void *handle);

/*
 * Replaces MAPPOPolicy.get_actions (algos/mappo.py:43-49) -> R_Actor.forward (r_actor_critic.py:43-57) +
 * R_Critic.forward (:111-121) on one vec-env step.
 *   d_obs     [n_envs, N, D] float32 (the env's obs buffer; the critic reads it as [n_envs, N*D])
 *   d_actions [n_envs*N, 2]  a = mu + exp(logstd) * eps, eps ~ N(0,1) from Philox4x32-10 keyed by (seed, offset,
 *             agent row); deterministic != 0 returns mu (FixedNormal.mode)
 *   d_logp    [n_envs*N]     sum over the 2 action dims of Normal.log_prob (distributions.py:33-35)     [opt]
 *   d_values  [n_envs]       critic output in ValueNorm-normalised units, one per env (the reference
 *             evaluates N identical rows per env)                                                     [opt]
 * d_actor / d_critic may be NULL to skip that net (get_values = actor NULL; act = critic NULL).
 */
int dcc_mappo_act(void *handle, const float *d_actor, const float *d_critic, const float *d_obs, int n_envs,
                  uint64_t seed, uint64_t offset, int deterministic, float *d_actions, float *d_logp, float *d_values,
                  dcc_stream_t stream);

/*
 * Replaces MAPPOPolicy.evaluate_actions (algos/mappo.py:55-61), forward only: log-prob of GIVEN actions, the
 * mean, and the values.  Shapes as dcc_mappo_act; d_mu [n_envs*N, 2] optional.
 */
int dcc_mappo_evaluate(void *handle, const float *d_actor, const float *d_critic, const float *d_obs,
                       const float *d_actions, int n_envs, float *d_logp, float *d_values, float *d_mu,
                       dcc_stream_t stream);

/*
 * ---- compact-state learner path ("compact-state storage + obs regeneration") ----------------------
 * The observation rows SharedReplayBuffer stores (buffer/shared_buffer.py:15-70: obs [T+1, E, N, D] and share_obs
 * [T+1, E, N, N*D]) are an affine function of the env's compact state (scenarios/coverage.py:99-110), so the rollout
 * keeps only that state — pos_vel [rows, N, 4] float64 and energy [rows, M] uint8, 32 N + M bytes per env step instead
 * of 4 N D — and the first layer of both nets is evaluated from it directly: xhat W1^T = f (W1 A)^T with f the
 * LayerNorm-scaled state features (csrc/dcc_compact.cuh; exact algebra, pinned by oracle/compact_oracle.py).
 *   dcc_mappo_set_env_layout    gives the learner handle the env's PoI table (host, M x 2 float64) and m_energy; required
 *                               before any *_state call.  DCC_ERR_UNSUPPORTED unless obs_dim == 2N + 2 + 5M and the critic
 *                               is centralised (use_centralized_V), the only layouts the identity covers.
 *   dcc_mappo_act_state         dcc_mappo_act            with (d_pos_vel, d_energy) in place of d_obs
 *   dcc_mappo_evaluate_state    dcc_mappo_evaluate       likewise
 *   dcc_mappo_epoch_grads_state dcc_mappo_epoch_grads    likewise (rows = T*E env steps, time-major like the obs buffer)
 *   dcc_mappo_minibatch_grads_state  dcc_mappo_minibatch_grads likewise (num_mini_batch > 1)
 *   dcc_obs_from_state          regenerates the observation rows [n_rows, N, D] float32, bit-identical to dcc_env_step's
 *                               (the reference-shaped view `buffer.obs[t]` of a compact rollout)
 */
int dcc_mappo_set_env_layout(void *handle, int n_pois, const double *h_poi_xy, double m_energy);
int dcc_mappo_act_state(void *handle, const float *d_actor, const float *d_critic, const double *d_pos_vel,
                        const uint8_t *d_energy, int n_envs, uint64_t seed, uint64_t offset, int deterministic,
                        float *d_actions, float *d_logp, float *d_values, dcc_stream_t stream);
int dcc_mappo_evaluate_state(void *handle, const float *d_actor, const float *d_critic, const double *d_pos_vel,
                             const uint8_t *d_energy, const float *d_actions, int n_envs, float *d_logp, float *d_values,
                             float *d_mu, dcc_stream_t stream);
int dcc_mappo_epoch_grads_state(void *handle, const float *d_actor, const float *d_critic, float *d_grad_actor,
                                float *d_grad_critic, const double *d_pos_vel, const uint8_t *d_energy,
                                const float *d_actions, const float *d_logp_old, const float *d_values,
                                const float *d_returns, float *d_vn_state, const double *d_stats4, double n_rows_global,
                                int T, int E, double *d_epoch_stats, dcc_stream_t stream);
int dcc_obs_from_state(void *handle, const double *d_pos_vel, const uint8_t *d_energy, int n_rows, float *d_obs,
                       dcc_stream_t stream);
/* dcc_mappo_minibatch_grads on a compact rollout: the rows of the minibatch (agent-row indices of the generator's
 * permutation, buffer/shared_buffer.py:238-262) are gathered from the stored state — the actor features of agent
 * idx % N of state row idx / N, the critic's centralised features of that state row once per sampled agent row. */
int dcc_mappo_minibatch_grads_state(void *handle, const float *d_actor, const float *d_critic, float *d_grad_actor,
                                    float *d_grad_critic, const double *d_pos_vel, const uint8_t *d_energy,
                                    const float *d_actions, const float *d_logp_old, const float *d_values,
                                    const float *d_returns, float *d_vn_state, const double *d_stats4, double n_rows_global,
                                    const int64_t *d_row_index, int64_t n_index, const double *d_ret_sums,
                                    double n_index_global, double *d_epoch_stats, dcc_stream_t stream);

/*
 * Replaces the per-step bookkeeping of Learner.insert (learner.py:254-276): reward of the env (entry 0 of the N
 * equal entries) and masks = 1 - done.  d_rew_in [E,N] float32, d_done_in [E,N] uint8 -> d_rew_out [E], d_mask_out [E].
 */
int dcc_rollout_insert(const float *d_rew_in, const uint8_t *d_done_in, int n_envs, int n_agents, float *d_rew_out,
                       float *d_mask_out, dcc_stream_t stream);

/*
 * Replaces SharedReplayBuffer.compute_returns, live branch (buffer/shared_buffer.py:199-208) with
 * ValueNorm.denormalize folded in (valuenorm.py:68-79).
 *   d_rewards [T,E], d_values [T+1,E] (d_values[T] = bootstrap value), d_masks [T+1,E], d_vn_state[3]
 *   -> d_returns [T+1,E] (rows 0..T-1 written)
 * cfg.use_gae = 0 selects the discounted-return branch (:209-212; row T = the raw bootstrap value is written too);
 * cfg.use_valuenorm = 0 drops the denormalisation (d_vn_state may then be NULL, here and in the calls below).
 */
int dcc_mappo_gae(void *handle, const float *d_rewards, const float *d_values, const float *d_masks,
                  const float *d_vn_state, int T, int E, float *d_returns, dcc_stream_t stream);

/*
 * MAPPOTrainer.train prologue (algos/mappo.py:189-198): advantage = returns - denormalize(value_preds) and its
 * global mean / population std, plus the batch statistics ValueNorm.update needs every epoch.
 * d_stats_out: 4 float64 {sum adv, sum adv^2, sum ret, sum ret^2} over this rank's T*E env-step rows — a
 * caller-visible buffer so that a multi-GPU host can all-reduce (SUM) it before the epochs.
 * Also snapshots the ValueNorm state (advantages use the state BEFORE this update's ValueNorm.update calls).
 */
int dcc_mappo_train_begin(void *handle, const float *d_returns, const float *d_values, const float *d_vn_state, int T,
                          int E, double *d_stats_out, dcc_stream_t stream);

/*
 * One PPO epoch up to and including total_loss.backward() (algos/mappo.py:133-174) on this rank's rollout:
 * ValueNorm.update (valuenorm.py:38-55, in place on d_vn_state), evaluate_actions, clipped-ratio policy loss
 * (2x: two equal log-prob columns, shared_buffer.py:61-62), Gaussian entropy, clipped one-sided-Huber value
 * loss (utils/util.py:36-39), and the gradients of both nets.
 *   d_obs [T+1,E,N,D] (first T used), d_actions [T,E,N,2], d_logp_old [T,E,N], d_values / d_returns [T+1,E]
 *   d_stats4        the (all-reduced) output of dcc_mappo_train_begin; n_rows_global = global T*E
 *   d_grad_actor / d_grad_critic   flat, zeroed here then accumulated: SUM over this rank's rows of the
 *                   per-row gradient already divided by the GLOBAL row count, so a SUM all-reduce across ranks
 *                   yields the big-batch gradient (the entropy-bonus term is added once, in dcc_mappo_apply)
 *   d_epoch_stats   4 float64, += {policy_loss_sum, value_loss_sum, ratio_sum (over agent rows), dist_entropy}
 */
int dcc_mappo_epoch_grads(void *handle, const float *d_actor, const float *d_critic, float *d_grad_actor,
                          float *d_grad_critic, const float *d_obs, const float *d_actions, const float *d_logp_old,
                          const float *d_values, const float *d_returns, float *d_vn_state, const double *d_stats4,
                          double n_rows_global, int T, int E, double *d_epoch_stats, dcc_stream_t stream);

/*
 * num_mini_batch > 1 — SharedReplayBuffer.feed_forward_generator (buffer/shared_buffer.py:219-279) + ppo_update
 * (algos/mappo.py:133-187) for ONE minibatch.  The reference draws `torch.randperm(T*E*N)` over AGENT rows
 * (row = (t*E + e)*N + n) once per epoch and cuts it into num_mini_batch equal index lists; the caller passes one
 * such list (device int64, `n_index` entries; in a multi-GPU job: this rank's share).  Rows are gathered on the
 * fly — observations through the index inside the input-LayerNorm kernel, per-row scalars inside the loss kernels —
 * so no minibatch copy of the rollout is ever materialised.  The critic is evaluated per agent row here, as the
 * reference does (the N agent rows of an env step fall into different minibatches).
 *   dcc_mappo_minibatch_stats   d_ret_sums_out[2] = {sum, sum of squares} of the returns of the indexed rows: the
 *                               batch statistics ValueNorm.update(return_batch) needs (mappo.py:107); all-reduce
 *                               (SUM) across ranks before the next call
 *   dcc_mappo_minibatch_grads   as dcc_mappo_epoch_grads on the indexed rows.  n_rows_global = global T*E (the
 *                               advantage statistics in d_stats4 stay those of the WHOLE rollout, mappo.py:189-198);
 *                               n_index_global = global agent rows of this minibatch (the loss denominators)
 */
int dcc_mappo_minibatch_stats(void *handle, const float *d_returns, const int64_t *d_row_index, int64_t n_index,
                              double *d_ret_sums_out, dcc_stream_t stream);
int dcc_mappo_minibatch_grads(void *handle, const float *d_actor, const float *d_critic, float *d_grad_actor,
                              float *d_grad_critic, const float *d_obs, const float *d_actions, const float *d_logp_old,
                              const float *d_values, const float *d_returns, float *d_vn_state, const double *d_stats4,
                              double n_rows_global, const int64_t *d_row_index, int64_t n_index,
                              const double *d_ret_sums, double n_index_global, double *d_epoch_stats,
                              dcc_stream_t stream);

/*
 * clip_grad_norm_(max_grad_norm) + torch.optim.Adam.step for one net (algos/mappo.py:176-185), fused over the
 * flat buffers, on the (all-reduced) gradient.  which = 0 actor (adds the entropy-bonus gradient first), 1 critic.
 * step = 1-based Adam step count; lr = the decayed learning rate (utils/util.py:29-33).
 * d_grad_norm_sq_out (1 float64, optional) receives the squared pre-clip gradient norm.
 */
int dcc_mappo_apply(void *handle, int which, float *d_params, float *d_grads, float *d_adam_m, float *d_adam_v,
                    float lr, int64_t step, double *d_grad_norm_sq_out, dcc_stream_t stream);

/*
 * ---- recurrent policies (mappo.yaml use_recurrent_policy / use_naive_recurrent_policy; dcc_mappo_cfg.recurrent_N >= 1) --------
 * Handles created with recurrent_N >= 1 use these entry points instead of dcc_mappo_act / _evaluate / _epoch_grads /
 * _minibatch_grads (which return DCC_ERR_UNSUPPORTED for them), on materialised observation rows.
 *
 * dcc_mappo_act_rnn replaces MAPPOPolicy.get_actions / get_values / act / evaluate_actions for ONE vec-env step
 * (algos/mappo.py:43-65 -> R_Actor.forward / R_Critic.forward with the RNNLayer, r_actor_critic.py:43-57,111-121,
 * rnn.py:24-29): hidden state *= mask, one GRU step per layer, LayerNorm, head.
 *   d_h_actor  [n_envs*N, recurrent_N, H]  hidden states in (learner.py:233-236: buffer.rnn_states[step])
 *   d_h_critic [n_envs, recurrent_N, H]    ONE per env: the N agent rows of an env carry identical inputs, masks and states
 *   d_masks    [n_envs]                    buffer.masks[step] (identical for the agents of an env, learner.py:266-267)
 *   mode       0 = sample / deterministic mode (writes d_actions, d_logp optional); 1 = log-prob of the GIVEN d_actions
 *   d_h_actor_out / d_h_critic_out        new hidden states, same shapes (may alias nothing of the inputs; optional)
 * The caller zeroes the new states of finished episodes before storing them (learner.py:258-265).
 *
 * dcc_mappo_seq_grads replaces one ppo_update (algos/mappo.py:133-187) on a minibatch drawn by recurrent_generator
 * (buffer/shared_buffer.py:378-470, chunks of data_chunk_length steps) or naive_recurrent_generator (:281-376, whole
 * episodes): n_seq sequences of seq_len consecutive entries of the rollout flattened in (env, agent, time) order, each
 * started from the stored hidden state of its first entry, BPTT inside the sequence.
 *   d_row_index [n_seq * seq_len] int64  agent-row indices (t*E*N + e*N + a) of the minibatch, arranged in passes of
 *               P = dcc_mappo_rnn_pass_seqs(handle, seq_len) sequences: pass p holds sequences [p*P, min((p+1)*P, n_seq)),
 *               time-major inside the pass (entry l * S_p + s = step l of the pass's s-th sequence)
 *   d_h_actor   [(T+1)*E*N, recurrent_N, H], d_h_critic [(T+1)*E, recurrent_N, H], d_masks [(T+1)*E]: the rollout's stored
 *               hidden states and masks; the other arguments as dcc_mappo_minibatch_grads (n_index_global = n_seq * seq_len
 *               summed over ranks; d_ret_sums from dcc_mappo_minibatch_stats over the same index list).
 */
int dcc_mappo_rnn_pass_seqs(void *handle, int seq_len);
int dcc_mappo_act_rnn(void *handle, const float *d_actor, const float *d_critic, const float *d_obs, int n_envs,
                      const float *d_h_actor, const float *d_h_critic, const float *d_masks, int mode, uint64_t seed,
                      uint64_t offset, int deterministic, float *d_actions, float *d_logp, float *d_values,
                      float *d_h_actor_out, float *d_h_critic_out, dcc_stream_t stream);
int dcc_mappo_seq_grads(void *handle, const float *d_actor, const float *d_critic, float *d_grad_actor, float *d_grad_critic,
                        const float *d_obs, const float *d_h_actor, const float *d_h_critic, const float *d_masks,
                        const float *d_actions, const float *d_logp_old, const float *d_values, const float *d_returns,
                        float *d_vn_state, const double *d_stats4, double n_rows_global, const int64_t *d_row_index,
                        int64_t n_seq, int seq_len, const double *d_ret_sums, double n_index_global, double *d_epoch_stats,
                        dcc_stream_t stream);

/* Kernel-level test hook: C[M,N] (+)= op(A)[M,K] op(B)[K,N], row-major with leading dimensions; ta/tb = operand
 * stored transposed.  backend as dcc_mappo_cfg.gemm_backend (2 requires the shapes the tcgen05 kernels cover); 3 = the
 * fp16 hi/lo split forward kernel on its own (ta == 0, N == 256, |A| < 65504, |B| < 255). */
int dcc_op_gemm(void *handle, int backend, int ta, int tb, int M, int N, int K, const float *d_A, int lda,
                const float *d_B, int ldb, float *d_C, int ldc, int accumulate, dcc_stream_t stream);

const char *dcc_status_string(int status);
const char *dcc_last_cuda_error(void);
int dcc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DCC_B200_H_ */

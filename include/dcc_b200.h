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

#define DCC_ABI_VERSION 1
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

const char *dcc_status_string(int status);
const char *dcc_last_cuda_error(void);
int dcc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DCC_B200_H_ */

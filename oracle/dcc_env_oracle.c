/*
 * dcc_env_oracle.c — CPU restatement of the reference env step.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may load
 * this library, and only as the checker or the timed CPU baseline.  The product path
 * (dynamic-coverage-control_b200/) never links, imports or calls it.
 *
 * What it restates (paths relative to /root/reference/uav_dcc_control/):
 *   envs/mpe/multiagent/environment.py:86-110,153-190   MultiAgentEnv.step / _set_action
 *   envs/mpe/multiagent/CoverageWorld.py:57-174          CoverageWorld.step and its five phases
 *   envs/mpe/multiagent/scenarios/coverage.py:64-117     reset_world / reward / observation / done
 *   envs/wrappers.py:222-235                             DummyVecEnv auto-reset rule
 *
 * Pinning: the reference ships no tests or golden vectors.  This file is pinned against
 * tests/golden/env_*.npz, which tests/golden/make_golden.py produced by running the UNMODIFIED
 * reference in the build container (tests/test_oracle_env.py: obs/state/flags bit-exact, reward
 * to 1e-12 relative).
 *
 * Floating-point contract (what "bit-exact" is anchored to):
 *   - all state is IEEE binary64; every operation below is a separately rounded +,-,*,/ or sqrt,
 *     in the reference's evaluation order (compile with -ffp-contract=off);
 *   - np.linalg.norm of a 2-vector is sqrt(ddot(x,x)); the OpenBLAS ddot in this image's numpy
 *     evaluates x0*x0 rounded, then fma(x1,x1,.) (probe: 0/200000 mismatches vs that form,
 *     8.3 % vs the unfused form), so norm2() below uses exactly that;
 *   - the force actions are float32 (learner.py:241,250); `u *= 5.0` and `(u/mass)*dt` are float32
 *     products (NumPy keeps float32 against Python scalars), `p_force[a] += f` rounds the float64
 *     sum back to float32 (CoverageWorld.py:115-116), and `p_vel += ...` promotes to float64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define DCC_ORACLE_MAX_AGENTS 32

typedef struct dcc_oracle_cfg {
    int32_t n_agents;       /* N <= 32 */
    int32_t n_pois;         /* M */
    double r_cover;         /* dcc.yaml:8 */
    double r_comm;          /* dcc.yaml:9 */
    double comm_r_scale;    /* world value: shipped 0.9 (CoverageWorld.py:7), generalised = cfg */
    double contact_force;   /* world value: 1e2 * comm_force_scale (core.py:109, CoverageWorld.py:16) */
    double contact_margin;  /* 1e-3, core.py:110 */
    double dt;              /* 0.1, CoverageWorld.py:23 */
    double damping;         /* 0.25, core.py:107 */
    double max_speed;       /* 0.5, coverage.py:53 */
    double sensitivity;     /* 5.0, environment.py:187 */
    double m_energy;        /* 5.0, coverage.py:23 */
    double rew_cover;       /* 75, coverage.py:25 */
    double rew_done;        /* 1500, coverage.py:26 */
    double rew_out;         /* -100, coverage.py:28 */
} dcc_oracle_cfg;

/* np.linalg.norm on a 2-vector as this image's numpy/OpenBLAS evaluates it (see header). */
static inline double norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }

/* np.logaddexp(0, z) (numpy npymath: npy_logaddexp) */
static inline double logaddexp0(double z) {
    if (z == 0.0) return 0.0 + 0.6931471805599453094172321214581766;
    double tmp = 0.0 - z;
    if (tmp > 0) return 0.0 + log1p(exp(-tmp));
    return z + log1p(exp(tmp));
}

/* CoverageWorld.get_connect_force, CoverageWorld.py:129-140.  f_a = -F, f_b = +F */
static void connect_force(const dcc_oracle_cfg *c, const double *pv, int a, int b, double F[2]) {
    if (a == b) { F[0] = F[1] = 0.0; return; }
    double dx = pv[4 * a + 0] - pv[4 * b + 0];
    double dy = pv[4 * a + 1] - pv[4 * b + 1];
    double dist = norm2(dx, dy);
    double dist_max = (c->r_comm + c->r_comm) * c->comm_r_scale;
    double k = c->contact_margin;
    double pen = logaddexp0((dist - dist_max) / k) * k;
    F[0] = c->contact_force * dx / dist * pen;
    F[1] = c->contact_force * dy / dist * pen;
}

/* Scenario.observation, coverage.py:99-110; float32 cast = SharedReplayBuffer dtype (shared_buffer.py:40) */
static void write_obs(const dcc_oracle_cfg *c, const double *poi, const double *pv, const uint8_t *energy,
                      float *obs) {
    const int N = c->n_agents, M = c->n_pois;
    const int D = 4 + 2 * (N - 1) + 5 * M;
    for (int i = 0; i < N; ++i) {
        float *o = obs + (size_t)i * D;
        double px = pv[4 * i + 0], py = pv[4 * i + 1];
        *o++ = (float)pv[4 * i + 2];
        *o++ = (float)pv[4 * i + 3];
        *o++ = (float)px;
        *o++ = (float)py;
        for (int k = 0; k < N; ++k) {
            if (k == i) continue;
            *o++ = (float)(pv[4 * k + 0] - px);
            *o++ = (float)(pv[4 * k + 1] - py);
        }
        for (int j = 0; j < M; ++j) {
            *o++ = (float)(poi[2 * j + 0] - px);
            *o++ = (float)(poi[2 * j + 1] - py);
            *o++ = (float)energy[j];
            *o++ = (float)c->m_energy;
            *o++ = ((double)energy[j] >= c->m_energy) ? 1.0f : 0.0f;
        }
    }
}

/* One env, one step.  pv: N x (px,py,vx,vy) f64 in/out; energy: M u8 in/out (done_j == energy_j >= m_energy,
 * CoverageWorld.py:160-161 freezes the energy of a done PoI).  Optional outputs may be NULL. */
static void step_one(const dcc_oracle_cfg *c, const double *poi, const float *act, double *pv, uint8_t *energy,
                     float *obs, double *rew64, uint8_t *done_out, double *cov_out, uint8_t *connect_out,
                     uint32_t *adj_rows, uint32_t *adjs_rows, double *pv_pre, uint8_t *energy_pre) {
    const int N = c->n_agents, M = c->n_pois;
    float u[DCC_ORACLE_MAX_AGENTS][2];
    double dist[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS];
    double adj[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS], adjs[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS];
    int connect = 0, connect_s = 0;

    /* 1. _set_action: u = float32(a) * 5.0 in float32 (environment.py:186-190) */
    const float sens = (float)c->sensitivity;
    for (int i = 0; i < N; ++i) {
        u[i][0] = act[2 * i + 0] * sens;
        u[i][1] = act[2 * i + 1] * sens;
    }

    /* 2. update_connect (CoverageWorld.py:70-93), pre-move positions */
    memset(adj, 0, sizeof adj);
    memset(adjs, 0, sizeof adjs);
    if (c->comm_r_scale > 0) {
        const double thr = c->r_comm + c->r_comm;
        const double thr_s = c->comm_r_scale * (c->r_comm + c->r_comm);
        for (int a = 0; a < N; ++a) {
            for (int b = 0; b < N; ++b) {
                dist[a][b] = norm2(pv[4 * a + 0] - pv[4 * b + 0], pv[4 * a + 1] - pv[4 * b + 1]);
                if (dist[a][b] < thr) {
                    adj[a][b] = 1;
                    if (dist[a][b] < thr_s) adjs[a][b] = 1;
                }
            }
            dist[a][a] = 1e5;
            adj[a][a] = 0;
            adjs[a][a] = 0;
        }
        /* connect  = all(I + sum_{k=1}^{N-1} adj^k        > 0)
         * connect_ = all(I + sum_{k=1}^{N-1} adj^k . adj_ > 0)   (the connect_mat[-1] quirk, :90) */
        static _Thread_local double P[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS],
            Q[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS], S[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS],
            Ss[DCC_ORACLE_MAX_AGENTS][DCC_ORACLE_MAX_AGENTS];
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) P[a][b] = S[a][b] = Ss[a][b] = (a == b) ? 1.0 : 0.0;
        for (int it = 0; it < N - 1; ++it) {
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    double s = 0;
                    for (int k = 0; k < N; ++k) s += P[a][k] * adj[k][b];
                    Q[a][b] = s;
                }
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    P[a][b] = Q[a][b]; /* walk counts: <= 31^31 ~ 1.7e46 for N <= 32, exact sign in float64 */
                    S[a][b] += P[a][b];
                }
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    double s = 0;
                    for (int k = 0; k < N; ++k) s += P[a][k] * adjs[k][b];
                    Ss[a][b] += s;
                }
        }
        connect = connect_s = 1;
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) {
                if (!(S[a][b] > 0)) connect = 0;
                if (!(Ss[a][b] > 0)) connect_s = 0;
            }
    }

    /* 3. apply_connect_force (CoverageWorld.py:100-127): increments land in the float32 action array */
    if (c->contact_force > 0 && !connect_s) {
        int n_iso = 0;
        for (int a = 0; a < N; ++a) {
            double col = 0;
            for (int b = 0; b < N; ++b) col += adjs[b][a];
            if (col == 0) {
                ++n_iso;
                int bmin = 0;
                for (int b = 1; b < N; ++b)
                    if (dist[a][b] < dist[a][bmin]) bmin = b; /* np.argmin: first minimum */
                double F[2];
                connect_force(c, pv, a, bmin, F);
                u[a][0] = (float)((double)u[a][0] + (-F[0]));
                u[a][1] = (float)((double)u[a][1] + (-F[1]));
                u[bmin][0] = (float)((double)u[bmin][0] + F[0]);
                u[bmin][1] = (float)((double)u[bmin][1] + F[1]);
            }
        }
        if (n_iso == 0) {
            const double lim = c->comm_r_scale * 2 * c->r_comm;
            int am = 0, bm = 0;
            double best = 0;
            int first = 1;
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    double d = dist[a][b] < lim ? 1e5 : dist[a][b];
                    if (first || d < best) { best = d; am = a; bm = b; first = 0; }
                }
            double F[2];
            connect_force(c, pv, am, bm, F);
            u[am][0] = (float)((double)u[am][0] + (-F[0]));
            u[am][1] = (float)((double)u[am][1] + (-F[1]));
            u[bm][0] = (float)((double)u[bm][0] + F[0]);
            u[bm][1] = (float)((double)u[bm][1] + F[1]);
        }
    }

    /* 4. integrate_state (CoverageWorld.py:142-155) */
    const float dt32 = (float)c->dt;
    for (int i = 0; i < N; ++i) {
        double vx = pv[4 * i + 2] * (1 - c->damping);
        double vy = pv[4 * i + 3] * (1 - c->damping);
        vx += (double)(float)((float)(u[i][0] / 1.0f) * dt32);
        vy += (double)(float)((float)(u[i][1] / 1.0f) * dt32);
        double speed = sqrt(vx * vx + vy * vy);
        if (speed > c->max_speed) {
            double s = sqrt(vx * vx + vy * vy);
            vx = vx / s * c->max_speed;
            vy = vy / s * c->max_speed;
        }
        pv[4 * i + 2] = vx;
        pv[4 * i + 3] = vy;
        pv[4 * i + 0] += vx * c->dt;
        pv[4 * i + 1] += vy * c->dt;
    }

    /* 5. update_energy (CoverageWorld.py:157-174) */
    int num_done = 0, n_just = 0;
    uint8_t just[1024];
    uint8_t *justv = (M <= (int)sizeof just) ? just : (uint8_t *)malloc((size_t)M);
    for (int j = 0; j < M; ++j) {
        justv[j] = 0;
        if ((double)energy[j] >= c->m_energy) { ++num_done; continue; }
        for (int i = 0; i < N; ++i) {
            double d = norm2(poi[2 * j + 0] - pv[4 * i + 0], poi[2 * j + 1] - pv[4 * i + 1]);
            if (d <= c->r_cover) energy[j] += 1;
        }
        if ((double)energy[j] >= c->m_energy) { justv[j] = 1; ++n_just; ++num_done; }
    }
    const double coverage_rate = (double)num_done / (double)M;

    /* 6. obs (pre-reset), 7. reward: Scenario.reward called once per agent, `just` consumed by the first
     *    call (coverage.py:80-97), then np.sum over agents and broadcast (environment.py:106-108) */
    double total = 0;
    for (int call = 0; call < N; ++call) {
        double rew = 0.0;
        int all_done = 1;
        for (int j = 0; j < M; ++j) {
            if (!((double)energy[j] >= c->m_energy)) {
                all_done = 0;
                double mn = 0;
                for (int i = 0; i < N; ++i) {
                    double d = norm2(pv[4 * i + 0] - poi[2 * j + 0], pv[4 * i + 1] - poi[2 * j + 1]);
                    if (i == 0 || d < mn) mn = d;
                }
                rew -= mn;
            } else if (call == 0 && justv[j]) {
                rew += c->rew_cover;
            }
        }
        if (all_done) rew += c->rew_done;
        for (int i = 0; i < N; ++i) {
            double ax = fabs(pv[4 * i + 0]), ay = fabs(pv[4 * i + 1]);
            double s = 0;
            if (ax > 1) s += ax - 1;
            if (ay > 1) s += ay - 1;
            rew += s * c->rew_out;
            if (ax > 1.5 || ay > 1.5) rew += c->rew_out;
        }
        total += rew;
    }
    if (justv != just) free(justv);

    /* 8. Scenario.done (coverage.py:112-117) */
    int done = 0;
    for (int i = 0; i < N; ++i)
        if (fabs(pv[4 * i + 0]) > 1.5 || fabs(pv[4 * i + 1]) > 1.5) done = 1;
    if (num_done == M) done = 1;

    if (pv_pre) memcpy(pv_pre, pv, sizeof(double) * 4 * N);
    if (energy_pre) memcpy(energy_pre, energy, M);

    /* 9. wrapper auto-reset (wrappers.py:104-109,226-232): obs replaced, reward/done/info kept */
    if (done) {
        memset(pv, 0, sizeof(double) * 4 * N);
        memset(energy, 0, M);
    }
    if (obs) write_obs(c, poi, pv, energy, obs);
    if (rew64) *rew64 = total;
    if (done_out) *done_out = (uint8_t)done;
    if (cov_out) *cov_out = coverage_rate;
    if (connect_out) *connect_out = (uint8_t)((connect ? 1 : 0) | (connect_s ? 2 : 0));
    for (int a = 0; a < N; ++a) {
        uint32_t r = 0, rs = 0;
        for (int b = 0; b < N; ++b) {
            if (adj[a][b] != 0) r |= 1u << b;
            if (adjs[a][b] != 0) rs |= 1u << b;
        }
        if (adj_rows) adj_rows[a] = r;
        if (adjs_rows) adjs_rows[a] = rs;
    }
}

/* ---- exported batch API (ctypes) ------------------------------------------------------------- */

int dcc_oracle_obs_dim(int n_agents, int n_pois) { return 4 + 2 * (n_agents - 1) + 5 * n_pois; }

/* Scenario.reset_world for E envs + reset observation. */
/* poi_stride: doubles between the PoI tables of consecutive envs — 0 = one table shared by all envs (the reference's
 * scenarios/pos_pois.npy), 2*M = every env instance has its own `pos_pois` (each SubprocVecEnv worker owns an
 * independent Scenario object, wrappers.py:141-146; coverage.py:15-17,71). */
int dcc_oracle_reset_layouts(const dcc_oracle_cfg *c, int n_envs, const double *poi, int poi_stride, double *pos_vel,
                             uint8_t *energy, float *obs) {
    const int N = c->n_agents, M = c->n_pois;
    const int D = dcc_oracle_obs_dim(N, M);
    if (N < 1 || N > DCC_ORACLE_MAX_AGENTS || M < 1) return -1;
    for (int e = 0; e < n_envs; ++e) {
        memset(pos_vel + (size_t)e * 4 * N, 0, sizeof(double) * 4 * N);
        memset(energy + (size_t)e * M, 0, M);
        if (obs)
            write_obs(c, poi + (size_t)e * poi_stride, pos_vel + (size_t)e * 4 * N, energy + (size_t)e * M,
                      obs + (size_t)e * N * D);
    }
    return 0;
}

int dcc_oracle_reset(const dcc_oracle_cfg *c, int n_envs, const double *poi, double *pos_vel, uint8_t *energy,
                     float *obs) {
    return dcc_oracle_reset_layouts(c, n_envs, poi, 0, pos_vel, energy, obs);
}

/* E independent envs, one step each.  Arrays are env-major; optional outputs may be NULL.
 * rew64: (E) float64 — shared by the N agents of an env; done: (E); coverage_rate: (E) f64 (terminal-step value);
 * connect: (E) bit0 = connect, bit1 = connect_; adj/adjs: (E,N) row bitmasks; *_pre: state before the auto-reset.
 * n_threads > 1 splits the env axis over POSIX threads (envs are independent, wrappers.py:141-146 runs one
 * process per env). */
typedef struct step_job {
    const dcc_oracle_cfg *c;
    int e0, e1;
    const double *poi;
    int poi_stride;
    const float *actions;
    double *pos_vel;
    uint8_t *energy;
    float *obs;
    double *rew64;
    uint8_t *done;
    double *coverage_rate;
    uint8_t *connect;
    uint32_t *adj, *adjs;
    double *pos_vel_pre;
    uint8_t *energy_pre;
} step_job;

static void *step_range(void *arg) {
    const step_job *j = (const step_job *)arg;
    const int N = j->c->n_agents, M = j->c->n_pois;
    const int D = 4 + 2 * (N - 1) + 5 * M;
    for (int e = j->e0; e < j->e1; ++e) {
        step_one(j->c, j->poi + (size_t)e * j->poi_stride, j->actions + (size_t)e * 2 * N, j->pos_vel + (size_t)e * 4 * N,
                 j->energy + (size_t)e * M, j->obs ? j->obs + (size_t)e * N * D : NULL,
                 j->rew64 ? j->rew64 + e : NULL, j->done ? j->done + e : NULL,
                 j->coverage_rate ? j->coverage_rate + e : NULL, j->connect ? j->connect + e : NULL,
                 j->adj ? j->adj + (size_t)e * N : NULL, j->adjs ? j->adjs + (size_t)e * N : NULL,
                 j->pos_vel_pre ? j->pos_vel_pre + (size_t)e * 4 * N : NULL,
                 j->energy_pre ? j->energy_pre + (size_t)e * M : NULL);
    }
    return NULL;
}

int dcc_oracle_step_layouts(const dcc_oracle_cfg *c, int n_envs, const double *poi, int poi_stride, const float *actions,
                            double *pos_vel, uint8_t *energy, float *obs, double *rew64, uint8_t *done,
                            double *coverage_rate, uint8_t *connect, uint32_t *adj, uint32_t *adjs, double *pos_vel_pre,
                            uint8_t *energy_pre, int n_threads) {
    const int N = c->n_agents, M = c->n_pois;
    if (N < 1 || N > DCC_ORACLE_MAX_AGENTS || M < 1 || n_envs < 0) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads > n_envs) n_threads = n_envs > 0 ? n_envs : 1;
    step_job jobs[256];
    pthread_t tids[256];
    for (int t = 0; t < n_threads; ++t) {
        step_job j = {c, (int)((long long)n_envs * t / n_threads), (int)((long long)n_envs * (t + 1) / n_threads),
                      poi, poi_stride, actions, pos_vel, energy, obs, rew64, done, coverage_rate, connect, adj, adjs,
                      pos_vel_pre, energy_pre};
        jobs[t] = j;
    }
    for (int t = 1; t < n_threads; ++t)
        if (pthread_create(&tids[t], NULL, step_range, &jobs[t]) != 0) return -2;
    step_range(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(tids[t], NULL);
    return 0;
}

int dcc_oracle_step(const dcc_oracle_cfg *c, int n_envs, const double *poi, const float *actions, double *pos_vel,
                    uint8_t *energy, float *obs, double *rew64, uint8_t *done, double *coverage_rate,
                    uint8_t *connect, uint32_t *adj, uint32_t *adjs, double *pos_vel_pre, uint8_t *energy_pre,
                    int n_threads) {
    return dcc_oracle_step_layouts(c, n_envs, poi, 0, actions, pos_vel, energy, obs, rew64, done, coverage_rate, connect,
                                   adj, adjs, pos_vel_pre, energy_pre, n_threads);
}

int dcc_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

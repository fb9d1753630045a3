// dcc_env.cu — the env reset/step hot path as one sm_100a kernel over all E env instances.
//
// One WARP per env instance (lane i < N owns UAV i; PoIs are strided over the 32 lanes):
//   phase 0  load compact state (UAV pos/vel float64, PoI energy uint8) and the float32 actions
//   phase 1  comm-radius adjacency bitmasks + warp connected-components (REDUX.OR label propagation)
//            -> connect / connect_                       CoverageWorld.update_connect   (:70-93)
//   phase 2  rule-based connectivity pull force          CoverageWorld.apply_connect_force (:100-140)
//   phase 3  integrate UAV kinematics                    CoverageWorld.integrate_state  (:142-155)
//   phase 4  N x M coverage distances, energy update, reward, done
//            CoverageWorld.update_energy (:157-174), Scenario.reward/done (coverage.py:80-117)
//   phase 5  auto-reset (wrappers.py:104-109), state write-back
//   phase 6  observation rows (coverage.py:99-110) built in shared memory and written to HBM with
//            cp.async.bulk (TMA bulk store), ping-pong staged when an env block exceeds the stage.
// Float64 state math uses separately rounded operations in the reference's order (dcc_common.cuh);
// threshold compares on sqrt(x) are replaced by exactly equivalent compares on x (host-computed bounds).
//
// Reference paths are relative to /root/reference/uav_dcc_control/envs/mpe/multiagent/.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "dcc_common.cuh"

namespace dcc {

struct EnvKParams {
    int E, N, M, D, H;  // H = 2N+2 = head length of an obs row
    int rows_per_chunk, n_chunks, n_buf, use_bulk;
    int stage_floats;  // rows_per_chunk * D
    int stage_stride;  // floats between the two stage buffers (16 B aligned)
    int slots;         // ceil(M / 32)
    int e_thr;         // smallest integer energy that counts as done
    int force_on;
    int pw_bytes;      // shared bytes per warp
    int poi_bytes;     // shared bytes of the PoI table
    float m_energy_f, sens, dt32;
    double thr2_adj, thr2_adjs, cover2, speed2;
    double keep, dt, max_speed;
    double lim_force, dist_max, margin, contact_force;
    double rew_cover, rew_done, rew_out;
    const double *poi;
    const double *poi_env;   // per-env PoI layouts [E, M, 2] (dcc_env_set_poi_layouts), or nullptr: every env uses `poi`
    double *pos_vel;
    uint8_t *energy;
    const float *actions;
    float *obs;
    float *rew;
    uint8_t *done;
    float *cov;
    uint8_t *connect;
    uint32_t *adj;
    uint32_t *adjs;
};

// Observation staging: chunks of R rows whose byte size is a multiple of 16 (cp.async.bulk granularity) and fits the
// per-warp stage budget.  Shared by the host-side planner and the compile-time specialisations.
constexpr size_t STAGE_BUDGET = 12 * 1024;        // upper bound of a chunk (shared-memory planning)
constexpr size_t SPEC_STAGE_BUDGET = 6000;        // chunk size actually used: single buffer, more resident warps
__host__ __device__ constexpr bool env_block_bulk_ok(int N, size_t row_bytes) { return ((size_t)N * row_bytes) % 16 == 0; }
__host__ __device__ constexpr int rows_per_chunk(int N, size_t row_bytes, size_t budget = STAGE_BUDGET) {
    const bool bulk = env_block_bulk_ok(N, row_bytes);
    const int unit = !bulk ? 1 : ((row_bytes % 16 == 0) ? 1 : ((row_bytes % 8 == 0) ? 2 : 4));
    int R = unit;
    for (int r = unit; r <= N; r += unit)
        if (N % r == 0 && r * row_bytes <= budget) R = r;
    return R;
}

// np.logaddexp(0, z)
__device__ __forceinline__ double logaddexp0(double z) {
    if (z == 0.0) return 0.6931471805599453094172321214581766;
    const double tmp = dsub(0.0, z);
    if (tmp > 0) return log1p(exp(-tmp));
    return dadd(z, log1p(exp(tmp)));
}

// CoverageWorld.get_connect_force (:129-140): returns F; f_a = -F, f_b = +F
__device__ __forceinline__ void connect_force(const EnvKParams &p, const double *s_pv, int a, int b, double &Fx,
                                              double &Fy) {
    if (a == b) { Fx = 0.0; Fy = 0.0; return; }
    const double dx = dsub(s_pv[a * 4 + 0], s_pv[b * 4 + 0]);
    const double dy = dsub(s_pv[a * 4 + 1], s_pv[b * 4 + 1]);
    const double dist = dsqrt(sqnorm2(dx, dy));
    const double pen = dmul(logaddexp0(ddiv(dsub(dist, p.dist_max), p.margin)), p.margin);
    Fx = dmul(ddiv(dmul(p.contact_force, dx), dist), pen);
    Fy = dmul(ddiv(dmul(p.contact_force, dy), dist), pen);
}

// lexicographic (value, index) warp argmin: np.argmin returns the first minimum
__device__ __forceinline__ void warp_argmin(double &v, int &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL_MASK, v, o);
        const int oi = __shfl_xor_sync(FULL_MASK, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

template <bool STEP>
__global__ void __launch_bounds__(512) dcc_env_kernel(const EnvKParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const int N = p.N, M = p.M, D = p.D, H = p.H;

    double *s_poi = reinterpret_cast<double *>(smem);
    unsigned char *wbase = smem + p.poi_bytes + (size_t)warp * p.pw_bytes;
    double *s_pv = reinterpret_cast<double *>(wbase);             // N x (px,py,vx,vy)
    uint8_t *s_en = wbase + (size_t)N * 32;                        // M energies
    float *stage = reinterpret_cast<float *>(wbase + (size_t)N * 32 + align_up((size_t)M, 16));

    for (int i = threadIdx.x; i < 2 * M; i += blockDim.x) s_poi[i] = p.poi[i];
    __syncthreads();

    const unsigned full = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
    int bufsel = 0;

    for (int e = blockIdx.x * wpc + warp; e < p.E; e += gridDim.x * wpc) {
        // PoI table of this env: the shared one (in shared memory) or its own layout (global, L1/L2-cached)
        const double *poi_e = p.poi_env ? p.poi_env + (size_t)e * 2 * M : s_poi;
        double px = 0.0, py = 0.0, vx = 0.0, vy = 0.0;
        if (STEP) {
            // ---- phase 0: load -----------------------------------------------------------------
            float ux = 0.f, uy = 0.f;
            if (lane < N) {
                const double2 *g = reinterpret_cast<const double2 *>(p.pos_vel + ((size_t)e * N + lane) * 4);
                const double2 a = g[0], b = g[1];
                px = a.x; py = a.y; vx = b.x; vy = b.y;
                const float2 act = reinterpret_cast<const float2 *>(p.actions)[(size_t)e * N + lane];
                ux = __fmul_rn(act.x, p.sens);  // _set_action: u = a * 5.0 in float32 (environment.py:186-190)
                uy = __fmul_rn(act.y, p.sens);
                s_pv[lane * 4 + 0] = px;
                s_pv[lane * 4 + 1] = py;
            }
            for (int j = lane; j < M; j += 32) s_en[j] = p.energy[(size_t)e * M + j];
            __syncwarp();

            // ---- phase 1: adjacency + connectivity on PRE-move positions (:70-93) ----------------
            unsigned rows = 0, rows_s = 0;
            if (lane < N) {
                for (int b = 0; b < N; ++b) {
                    const double dx = dsub(px, s_pv[b * 4 + 0]);
                    const double dy = dsub(py, s_pv[b * 4 + 1]);
                    const double d2 = sqnorm2(dx, dy);
                    const bool a1 = (b != lane) && (d2 < p.thr2_adj);
                    const bool a2 = a1 && (d2 < p.thr2_adjs);
                    rows |= (unsigned)a1 << b;
                    rows_s |= (unsigned)a2 << b;
                }
            }
            // connect <=> graph(adj) connected: label propagation from UAV 0, <= N-1 rounds
            unsigned reach = 1u;
            for (int it = 0; it < N - 1; ++it) {
                const unsigned contrib = (lane < N && ((reach >> lane) & 1u)) ? rows : 0u;
                const unsigned nr = reach | __reduce_or_sync(FULL_MASK, contrib);
                if (nr == reach) break;
                reach = nr;
            }
            const bool connect = (reach == full);
            // connect_ (the connect_mat[-1] quirk, :90) <=> connected AND every UAV has an adj_ neighbour (N>=3);
            // N == 2: always False; N == 1: True
            const bool all_nb = __all_sync(FULL_MASK, lane >= N || rows_s != 0u);
            const bool connect_s = (N == 1) ? true : ((N == 2) ? false : (connect && all_nb));

            // ---- phase 2: connectivity pull force (:100-127), float32 accumulation into u ---------
            if (p.force_on && !connect_s) {
                unsigned iso = __ballot_sync(FULL_MASK, lane < N && rows_s == 0u);
                if (iso) {
                    while (iso) {
                        const int a = __ffs(iso) - 1;
                        iso &= iso - 1;
                        double d = INFINITY;
                        if (lane < N) {
                            d = 1e5;
                            if (lane != a)
                                d = dsqrt(sqnorm2(dsub(s_pv[a * 4 + 0], px), dsub(s_pv[a * 4 + 1], py)));
                        }
                        int b = lane;
                        warp_argmin(d, b);
                        double Fx, Fy;
                        connect_force(p, s_pv, a, b, Fx, Fy);
                        if (lane == a) {
                            ux = __double2float_rn(dadd((double)ux, -Fx));
                            uy = __double2float_rn(dadd((double)uy, -Fy));
                        }
                        if (lane == b) {
                            ux = __double2float_rn(dadd((double)ux, Fx));
                            uy = __double2float_rn(dadd((double)uy, Fy));
                        }
                    }
                } else {
                    double best = INFINITY;
                    int bb = 0;
                    if (lane < N) {
                        for (int b = 0; b < N; ++b) {
                            double d = 1e5;
                            if (b != lane) d = dsqrt(sqnorm2(dsub(px, s_pv[b * 4 + 0]), dsub(py, s_pv[b * 4 + 1])));
                            if (d < p.lim_force) d = 1e5;
                            if (b == 0 || d < best) { best = d; bb = b; }
                        }
                    }
                    int a = lane;
                    warp_argmin(best, a);
                    const int b = __shfl_sync(FULL_MASK, bb, a);
                    double Fx, Fy;
                    connect_force(p, s_pv, a, b, Fx, Fy);
                    if (lane == a) {
                        ux = __double2float_rn(dadd((double)ux, -Fx));
                        uy = __double2float_rn(dadd((double)uy, -Fy));
                    }
                    if (lane == b) {
                        ux = __double2float_rn(dadd((double)ux, Fx));
                        uy = __double2float_rn(dadd((double)uy, Fy));
                    }
                }
            }
            __syncwarp();  // every lane is done reading pre-move positions

            // ---- phase 3: integrate (:142-155) ----------------------------------------------------
            bool hard_out = false;
            double bound_term = 0.0;
            if (lane < N) {
                vx = dmul(vx, p.keep);
                vy = dmul(vy, p.keep);
                vx = dadd(vx, (double)__fmul_rn(ux, p.dt32));  // (u / mass) * dt evaluated in float32
                vy = dadd(vy, (double)__fmul_rn(uy, p.dt32));
                const double s2 = dadd(dmul(vx, vx), dmul(vy, vy));
                if (s2 > p.speed2) {  // <=> sqrt(s2) > max_speed
                    const double s = dsqrt(s2);
                    vx = dmul(ddiv(vx, s), p.max_speed);
                    vy = dmul(ddiv(vy, s), p.max_speed);
                }
                px = dadd(px, dmul(vx, p.dt));
                py = dadd(py, dmul(vy, p.dt));
                s_pv[lane * 4 + 0] = px;
                s_pv[lane * 4 + 1] = py;
                s_pv[lane * 4 + 2] = vx;
                s_pv[lane * 4 + 3] = vy;
                // bounds part of Scenario.reward (coverage.py:92-96) and Scenario.done (:113-116)
                const double ax = fabs(px), ay = fabs(py);
                double s = 0.0;
                if (ax > 1.0) s = dadd(s, dsub(ax, 1.0));
                if (ay > 1.0) s = dadd(s, dsub(ay, 1.0));
                bound_term = dmul(s, p.rew_out);
                hard_out = (ax > 1.5) || (ay > 1.5);
                if (hard_out) bound_term = dadd(bound_term, p.rew_out);
            }
            __syncwarp();

            // ---- phase 4: coverage / energy (:157-174) and reward distances (coverage.py:82-86) ---
            double sumd = 0.0;
            int n_done = 0, n_just = 0;
            for (int s = 0; s < p.slots; ++s) {
                const int j = lane + 32 * s;
                const bool valid = j < M;
                bool now_done = false, just = false;
                if (valid) {
                    const double qx = poi_e[2 * j], qy = poi_e[2 * j + 1];
                    int en = s_en[j];
                    int cnt = 0;
                    double mind2 = INFINITY;
                    for (int i = 0; i < N; ++i) {
                        const double dx = dsub(qx, s_pv[i * 4 + 0]);
                        const double dy = dsub(qy, s_pv[i * 4 + 1]);
                        const double d2 = sqnorm2(dx, dy);
                        cnt += (d2 <= p.cover2) ? 1 : 0;  // <=> norm <= r_cover
                        mind2 = fmin(mind2, d2);
                    }
                    now_done = en >= p.e_thr;
                    if (!now_done) {
                        en += cnt;
                        now_done = en >= p.e_thr;
                        just = now_done;
                        s_en[j] = (uint8_t)en;
                    }
                    if (!now_done) sumd = dadd(sumd, dsqrt(mind2));  // min_i sqrt(.) == sqrt(min_i .)
                }
                n_done += __popc(__ballot_sync(FULL_MASK, valid && now_done));
                n_just += __popc(__ballot_sync(FULL_MASK, just));
            }
            // reward: N calls of Scenario.reward summed and shared (environment.py:106-108) = N*base + 75*n_just
            const bool all_done = (n_done == M);
            double base = warp_sum(dsub(bound_term, sumd));
            if (all_done) base = dadd(base, p.rew_done);
            const double R = dadd(dmul((double)N, base), dmul(p.rew_cover, (double)n_just));
            const bool done = __any_sync(FULL_MASK, hard_out) || all_done;

            if (lane < N) {
                if (p.rew) p.rew[(size_t)e * N + lane] = (float)R;
                if (p.done) p.done[(size_t)e * N + lane] = done ? 1 : 0;
                if (p.adj) p.adj[(size_t)e * N + lane] = rows;
                if (p.adjs) p.adjs[(size_t)e * N + lane] = rows_s;
            }
            if (lane == 0) {
                if (p.cov) p.cov[e] = (float)((double)n_done / (double)M);
                if (p.connect) p.connect[e] = (uint8_t)((connect ? 1 : 0) | (connect_s ? 2 : 0));
            }

            // ---- phase 5: wrapper auto-reset (wrappers.py:104-109) --------------------------------
            if (done) {
                __syncwarp();
                px = py = vx = vy = 0.0;
                if (lane < N) {
                    s_pv[lane * 4 + 0] = 0.0; s_pv[lane * 4 + 1] = 0.0;
                    s_pv[lane * 4 + 2] = 0.0; s_pv[lane * 4 + 3] = 0.0;
                }
                for (int j = lane; j < M; j += 32) s_en[j] = 0;
            }
        } else {
            // Scenario.reset_world (coverage.py:64-78)
            if (lane < N) {
                s_pv[lane * 4 + 0] = 0.0; s_pv[lane * 4 + 1] = 0.0;
                s_pv[lane * 4 + 2] = 0.0; s_pv[lane * 4 + 3] = 0.0;
            }
            for (int j = lane; j < M; j += 32) s_en[j] = 0;
        }
        __syncwarp();

        // state write-back
        if (lane < N) {
            double2 *g = reinterpret_cast<double2 *>(p.pos_vel + ((size_t)e * N + lane) * 4);
            g[0] = make_double2(px, py);
            g[1] = make_double2(vx, vy);
        }
        for (int j = lane; j < M; j += 32) p.energy[(size_t)e * M + j] = s_en[j];

        // ---- phase 6: observation rows (coverage.py:99-110) -> shared stage -> HBM ----------------
        if (p.obs) {
            const int R = p.rows_per_chunk;
            for (int c = 0; c < p.n_chunks; ++c) {
                float *buf = stage + (size_t)bufsel * p.stage_stride;
                if (p.use_bulk) {  // the bulk store that last read this buffer must have drained it
                    if (lane == 0) {
                        if (p.n_buf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                    }
                    __syncwarp();
                }
                const int r0 = c * R;
                // heads: [v_i, p_i, p_k - p_i (k != i)]
                for (int idx = lane; idx < R * H; idx += 32) {
                    const int r = idx / H;
                    const int cc = idx - r * H;
                    const int i = r0 + r;
                    double v;
                    if (cc < 2) v = s_pv[i * 4 + 2 + cc];
                    else if (cc < 4) v = s_pv[i * 4 + cc - 2];
                    else {
                        const int t = (cc - 4) >> 1, comp = (cc - 4) & 1;
                        const int k = t + (t >= i ? 1 : 0);
                        v = dsub(s_pv[k * 4 + comp], s_pv[i * 4 + comp]);
                    }
                    buf[r * D + cc] = (float)v;
                }
                // PoI sections: [q_j - p_i, energy_j, m_energy, done_j]
                for (int s = 0; s < p.slots; ++s) {
                    const int j = lane + 32 * s;
                    if (j < M) {
                        const double qx = poi_e[2 * j], qy = poi_e[2 * j + 1];
                        const int en = s_en[j];
                        const float fe = (float)en;
                        const float fd = (en >= p.e_thr) ? 1.f : 0.f;
                        float *o = buf + H + 5 * j;
                        for (int r = 0; r < R; ++r, o += D) {
                            o[0] = (float)dsub(qx, s_pv[(r0 + r) * 4 + 0]);
                            o[1] = (float)dsub(qy, s_pv[(r0 + r) * 4 + 1]);
                            o[2] = fe;
                            o[3] = p.m_energy_f;
                            o[4] = fd;
                        }
                    }
                }
                float *gdst = p.obs + ((size_t)e * N + r0) * D;
                if (p.use_bulk) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        bulk_store_s2g(gdst, buf, (uint32_t)(p.stage_floats * 4));
                        bulk_commit();
                    }
                    bufsel ^= (p.n_buf - 1);
                } else {
                    __syncwarp();
                    for (int idx = lane; idx < R * D; idx += 32) gdst[idx] = buf[idx];
                    __syncwarp();
                }
            }
        }
    }
    if (p.use_bulk && lane == 0) bulk_wait_all<0>();
}

// ---- compile-time specialisation ------------------------------------------------------------------
// Same algorithm and the same separately-rounded arithmetic as dcc_env_kernel, with N and M known at compile
// time: PoI coordinates and energies live in registers across the persistent env loop, every loop is fully
// unrolled (no index arithmetic, immediate shared-memory offsets), adjacency is evaluated one UNORDERED UAV
// pair per lane (N(N-1)/2 pairs over 32 lanes) with label propagation directly on the pair masks, obs-row
// heads are written one (i,k) pair per lane, and the constant m_energy column of the stage is written once
// per kernel instead of once per env.  The generic kernel stays the fallback for every other (N, M).
template <int N, int M>
struct EnvSpec {
    static constexpr int D = 4 + 2 * (N - 1) + 5 * M;
    static constexpr int H = 2 * N + 2;
    static constexpr int SLOTS = (M + 31) / 32;
    static constexpr int P = N * (N - 1) / 2;
    static constexpr int PPL = (P + 31) / 32 > 0 ? (P + 31) / 32 : 1;
    static constexpr size_t ROW_BYTES = (size_t)D * 4;
    static constexpr bool BULK = env_block_bulk_ok(N, ROW_BYTES);
    static constexpr int R = rows_per_chunk(N, ROW_BYTES, SPEC_STAGE_BUDGET);
    static constexpr int NCHUNK = N / R;
    // ONE stage buffer per warp and small chunks: the kernel is latency-bound, so resident warps matter more than
    // hiding the bulk store behind a second buffer (the next chunk waits for the TMA engine to have READ the stage).
    // Measured (µs per step, launch_bounds in brackets):
    //   8/64, 65 536 envs:   8-row chunk, 1 buffer, 20 warps/SM (92 regs) 125.5 | 4-row chunks, 1 buffer, 24 warps (80 regs)
    //                        122.7 | 32 warps (64 regs) 123.1 | 2-row chunks, 2 buffers, 32 warps 122.7
    //   16/256, 32 768 envs: 2 buffers, 10 warps/SM 547 | 1 buffer with 12 / 16 / 20 / 24 / 32 warps/SM (168 / 128 / 96 / 80 /
    //                        64 registers) 486 / 476 / 519 / 582 / 686
    static constexpr bool BIG_STAGE = (size_t)R * ROW_BYTES > 8 * 1024;
    static constexpr int NBUF = 1;
    static constexpr int STAGE_FLOATS = R * D;
    static constexpr int STAGE_STRIDE = (int)((((size_t)STAGE_FLOATS * 4 + 127) / 128 * 128) / 4);
    static constexpr int PW_BYTES = (int)(((size_t)N * 32 + (size_t)STAGE_STRIDE * 4 * NBUF + 127) / 128 * 128);
    // warps per CTA: 4, or 2 when a warp's stage is so large that 4-warp CTAs would leave shared memory unused
    static constexpr int WPC = (PW_BYTES * 4 > 64 * 1024) ? 2 : 4;
    static constexpr int SMEM = PW_BYTES * WPC;
    static constexpr int FIT = (int)(233472 / (SMEM + 1024));
    static constexpr int WANT_BLOCKS = BIG_STAGE ? 4 : 6;    // 4 x 4 warps at 128 registers, 6 x 4 warps at 80
    static constexpr int MIN_BLOCKS = FIT < 1 ? 1 : (FIT > WANT_BLOCKS ? WANT_BLOCKS : FIT);
    static constexpr int HP = R * (N - 1);               // ordered (row, other) head pairs per chunk
    static constexpr int HPL = (HP + 31) / 32 > 0 ? (HP + 31) / 32 : 1;
    static_assert(BULK, "specialisations require a 16-byte-multiple env block");
    static_assert(SMEM <= 227 * 1024, "stage does not fit");
};

template <int PPL>
__device__ __forceinline__ unsigned pick_word(const unsigned (&w)[PPL], int idx) {
    unsigned r = w[0];
#pragma unroll
    for (int t = 1; t < PPL; ++t) r = (idx == t) ? w[t] : r;
    return r;
}

template <int N, int M, bool STEP>
__global__ void __launch_bounds__(EnvSpec<N, M>::WPC * 32, EnvSpec<N, M>::MIN_BLOCKS)
dcc_env_spec_kernel(const EnvKParams p) {
    using S = EnvSpec<N, M>;
    constexpr int D = S::D, H = S::H, SLOTS = S::SLOTS, R = S::R;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem + (size_t)warp * S::PW_BYTES;
    double *s_pv = reinterpret_cast<double *>(wbase);
    float *stage = reinterpret_cast<float *>(wbase + (size_t)N * 32);
    constexpr unsigned full = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);

    // PoIs of this lane (j = lane + 32 s), kept in registers for the whole launch
    double qx[SLOTS], qy[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int j = lane + 32 * s;
        const bool valid = (M % 32 == 0) || (j < M);
        qx[s] = valid ? p.poi[2 * j] : 1e30;   // padding lanes: far away, never covered, masked below
        qy[s] = valid ? p.poi[2 * j + 1] : 1e30;
    }
    // constant m_energy column of every staged row: written once
#pragma unroll
    for (int b = 0; b < S::NBUF; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int j = lane + 32 * s;
                if ((M % 32 == 0) || (j < M)) stage[b * S::STAGE_STRIDE + r * D + H + 5 * j + 3] = p.m_energy_f;
            }
    // unordered UAV pairs of this lane: q -> (a < b)
    unsigned pm[S::PPL];   // (1<<a)|(1<<b), 0 for padding
    int pa[S::PPL], pb[S::PPL];
#pragma unroll
    for (int t = 0; t < S::PPL; ++t) {
        const int q = lane + 32 * t;
        int a = 0, b = 0;
#pragma unroll
        for (int aa = 0; aa < N - 1; ++aa) {
            const int start = aa * N - aa * (aa + 1) / 2;
            if (q >= start && q < start + (N - 1 - aa)) { a = aa; b = aa + 1 + (q - start); }
        }
        pa[t] = a; pb[t] = b;
        pm[t] = (q < S::P) ? ((1u << a) | (1u << b)) : 0u;
    }
    __syncwarp();
    int bufsel = 0;

    for (int e = blockIdx.x * S::WPC + warp; e < p.E; e += gridDim.x * S::WPC) {
        if (p.poi_env) {   // per-env PoI layout: this env's coordinates replace the shared ones in the registers
            const double *pe = p.poi_env + (size_t)e * 2 * M;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int j = lane + 32 * s;
                if ((M % 32 == 0) || (j < M)) { qx[s] = pe[2 * j]; qy[s] = pe[2 * j + 1]; }
            }
        }
        double px = 0.0, py = 0.0, vx = 0.0, vy = 0.0;
        int en[SLOTS];
        if (STEP) {
            // ---- phase 0: load ---------------------------------------------------------------------
            float ux = 0.f, uy = 0.f;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int j = lane + 32 * s;
                en[s] = ((M % 32 == 0) || (j < M)) ? (int)p.energy[(size_t)e * M + j] : 0;
            }
            if (lane < N) {
                const double2 *g = reinterpret_cast<const double2 *>(p.pos_vel + ((size_t)e * N + lane) * 4);
                const double2 a = g[0], b = g[1];
                px = a.x; py = a.y; vx = b.x; vy = b.y;
                const float2 act = reinterpret_cast<const float2 *>(p.actions)[(size_t)e * N + lane];
                ux = __fmul_rn(act.x, p.sens);
                uy = __fmul_rn(act.y, p.sens);
                *reinterpret_cast<double2 *>(s_pv + lane * 4) = a;
            }
            __syncwarp();

            // ---- phase 1: adjacency per unordered pair + connectivity (:70-93) -------------------------
            bool a1[S::PPL], a2[S::PPL];
#pragma unroll
            for (int t = 0; t < S::PPL; ++t) {
                const double2 A = *reinterpret_cast<const double2 *>(s_pv + pa[t] * 4);
                const double2 B = *reinterpret_cast<const double2 *>(s_pv + pb[t] * 4);
                const double d2 = sqnorm2(dsub(A.x, B.x), dsub(A.y, B.y));
                a1[t] = (pm[t] != 0u) && (d2 < p.thr2_adj);
                a2[t] = a1[t] && (d2 < p.thr2_adjs);
            }
            unsigned reach = 1u;
#pragma unroll 1
            for (int it = 0; it < N - 1; ++it) {
                unsigned contrib = 0u;
#pragma unroll
                for (int t = 0; t < S::PPL; ++t) contrib |= (a1[t] && (reach & pm[t])) ? pm[t] : 0u;
                const unsigned nr = reach | __reduce_or_sync(FULL_MASK, contrib);
                if (nr == reach) break;
                reach = nr;
            }
            const bool connect = (reach == full);
            unsigned nbc = 0u;
#pragma unroll
            for (int t = 0; t < S::PPL; ++t) nbc |= a2[t] ? pm[t] : 0u;
            const unsigned nbm = __reduce_or_sync(FULL_MASK, nbc);   // UAVs with at least one adj_ neighbour
            const bool connect_s = (N == 1) ? true : ((N == 2) ? false : (connect && nbm == full));

            if (p.adj || p.adjs) {   // optional row-bitmask outputs, rebuilt from the pair ballots
                unsigned bal1[S::PPL], bal2[S::PPL];
#pragma unroll
                for (int t = 0; t < S::PPL; ++t) {
                    bal1[t] = __ballot_sync(FULL_MASK, a1[t]);
                    bal2[t] = __ballot_sync(FULL_MASK, a2[t]);
                }
                if (lane < N) {
                    unsigned rows = 0u, rows_s = 0u;
#pragma unroll
                    for (int b = 0; b < N; ++b) {
                        if (b == lane) continue;
                        const int lo = min(lane, b), hi = max(lane, b);
                        const int q = lo * N - lo * (lo + 1) / 2 + (hi - lo - 1);
                        rows |= ((pick_word<S::PPL>(bal1, q >> 5) >> (q & 31)) & 1u) << b;
                        rows_s |= ((pick_word<S::PPL>(bal2, q >> 5) >> (q & 31)) & 1u) << b;
                    }
                    if (p.adj) p.adj[(size_t)e * N + lane] = rows;
                    if (p.adjs) p.adjs[(size_t)e * N + lane] = rows_s;
                }
            }

            // ---- phase 2: connectivity pull force (:100-127) ---------------------------------------------
            if (p.force_on && !connect_s) {
                unsigned iso = full & ~nbm;
                if (iso) {
                    while (iso) {
                        const int a = __ffs(iso) - 1;
                        iso &= iso - 1;
                        double d = INFINITY;
                        if (lane < N) {
                            d = 1e5;
                            if (lane != a)
                                d = dsqrt(sqnorm2(dsub(s_pv[a * 4 + 0], px), dsub(s_pv[a * 4 + 1], py)));
                        }
                        int b = lane;
                        warp_argmin(d, b);
                        double Fx, Fy;
                        connect_force(p, s_pv, a, b, Fx, Fy);
                        if (lane == a) {
                            ux = __double2float_rn(dadd((double)ux, -Fx));
                            uy = __double2float_rn(dadd((double)uy, -Fy));
                        }
                        if (lane == b) {
                            ux = __double2float_rn(dadd((double)ux, Fx));
                            uy = __double2float_rn(dadd((double)uy, Fy));
                        }
                    }
                } else {
                    double best = INFINITY;
                    int bb = 0;
                    if (lane < N) {
#pragma unroll 1
                        for (int b = 0; b < N; ++b) {
                            double d = 1e5;
                            if (b != lane) d = dsqrt(sqnorm2(dsub(px, s_pv[b * 4 + 0]), dsub(py, s_pv[b * 4 + 1])));
                            if (d < p.lim_force) d = 1e5;
                            if (b == 0 || d < best) { best = d; bb = b; }
                        }
                    }
                    int a = lane;
                    warp_argmin(best, a);
                    const int b = __shfl_sync(FULL_MASK, bb, a);
                    double Fx, Fy;
                    connect_force(p, s_pv, a, b, Fx, Fy);
                    if (lane == a) {
                        ux = __double2float_rn(dadd((double)ux, -Fx));
                        uy = __double2float_rn(dadd((double)uy, -Fy));
                    }
                    if (lane == b) {
                        ux = __double2float_rn(dadd((double)ux, Fx));
                        uy = __double2float_rn(dadd((double)uy, Fy));
                    }
                }
            }
            __syncwarp();

            // ---- phase 3: integrate (:142-155) -------------------------------------------------------------
            bool hard_out = false;
            double bound_term = 0.0;
            if (lane < N) {
                vx = dmul(vx, p.keep);
                vy = dmul(vy, p.keep);
                vx = dadd(vx, (double)__fmul_rn(ux, p.dt32));
                vy = dadd(vy, (double)__fmul_rn(uy, p.dt32));
                const double s2 = dadd(dmul(vx, vx), dmul(vy, vy));
                if (s2 > p.speed2) {
                    const double s = dsqrt(s2);
                    vx = dmul(ddiv(vx, s), p.max_speed);
                    vy = dmul(ddiv(vy, s), p.max_speed);
                }
                px = dadd(px, dmul(vx, p.dt));
                py = dadd(py, dmul(vy, p.dt));
                *reinterpret_cast<double2 *>(s_pv + lane * 4) = make_double2(px, py);
                *reinterpret_cast<double2 *>(s_pv + lane * 4 + 2) = make_double2(vx, vy);
                const double ax = fabs(px), ay = fabs(py);
                double s = 0.0;
                if (ax > 1.0) s = dadd(s, dsub(ax, 1.0));
                if (ay > 1.0) s = dadd(s, dsub(ay, 1.0));
                bound_term = dmul(s, p.rew_out);
                hard_out = (ax > 1.5) || (ay > 1.5);
                if (hard_out) bound_term = dadd(bound_term, p.rew_out);
            }
            __syncwarp();

            // ---- phase 4: coverage / energy (:157-174), reward distances (coverage.py:82-86) ---------------
            int cnt[SLOTS];
            double mind2[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) { cnt[s] = 0; mind2[s] = INFINITY; }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double2 pi = *reinterpret_cast<const double2 *>(s_pv + i * 4);
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const double d2 = sqnorm2(dsub(qx[s], pi.x), dsub(qy[s], pi.y));
                    cnt[s] += (d2 <= p.cover2) ? 1 : 0;
                    mind2[s] = fmin(mind2[s], d2);
                }
            }
            double sumd = 0.0;
            int n_done = 0, n_just = 0;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const bool valid = (M % 32 == 0) || (lane + 32 * s < M);
                bool now_done = en[s] >= p.e_thr, just = false;
                if (valid && !now_done) {
                    en[s] += cnt[s];
                    now_done = en[s] >= p.e_thr;
                    just = now_done;
                }
                if (valid && !now_done) sumd = dadd(sumd, dsqrt(mind2[s]));
                n_done += __popc(__ballot_sync(FULL_MASK, valid && now_done));
                n_just += __popc(__ballot_sync(FULL_MASK, just));
            }
            const bool all_done = (n_done == M);
            double base = warp_sum(dsub(bound_term, sumd));
            if (all_done) base = dadd(base, p.rew_done);
            const double Rw = dadd(dmul((double)N, base), dmul(p.rew_cover, (double)n_just));
            const bool done = __any_sync(FULL_MASK, hard_out) || all_done;

            if (lane < N) {
                if (p.rew) p.rew[(size_t)e * N + lane] = (float)Rw;
                if (p.done) p.done[(size_t)e * N + lane] = done ? 1 : 0;
            }
            if (lane == 0) {
                if (p.cov) p.cov[e] = (float)((double)n_done / (double)M);
                if (p.connect) p.connect[e] = (uint8_t)((connect ? 1 : 0) | (connect_s ? 2 : 0));
            }

            // ---- phase 5: wrapper auto-reset (wrappers.py:104-109) -------------------------------------------
            if (done) {
                px = py = vx = vy = 0.0;
                if (lane < N) {
                    *reinterpret_cast<double2 *>(s_pv + lane * 4) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2 *>(s_pv + lane * 4 + 2) = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) en[s] = 0;
            }
        } else {
            if (lane < N) {
                *reinterpret_cast<double2 *>(s_pv + lane * 4) = make_double2(0.0, 0.0);
                *reinterpret_cast<double2 *>(s_pv + lane * 4 + 2) = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) en[s] = 0;
        }
        __syncwarp();

        // state write-back
        if (lane < N) {
            double2 *g = reinterpret_cast<double2 *>(p.pos_vel + ((size_t)e * N + lane) * 4);
            g[0] = make_double2(px, py);
            g[1] = make_double2(vx, vy);
        }
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int j = lane + 32 * s;
            if ((M % 32 == 0) || (j < M)) p.energy[(size_t)e * M + j] = (uint8_t)en[s];
        }

        // ---- phase 6: observation rows (coverage.py:99-110) -> stage -> HBM (TMA bulk store) -------------------
        if (p.obs) {
            float fe[SLOTS], fd[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) { fe[s] = (float)en[s]; fd[s] = (en[s] >= p.e_thr) ? 1.f : 0.f; }
#pragma unroll
            for (int c = 0; c < S::NCHUNK; ++c) {
                float *buf = stage + (size_t)bufsel * S::STAGE_STRIDE;
                if (lane == 0) {
                    if (S::NBUF == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                }
                __syncwarp();
                constexpr int dummy = 0; (void)dummy;
                const int r0 = c * R;
                // heads: lanes < R write [v_i, p_i]; one ordered (i,k) pair per lane writes p_k - p_i
                if (lane < R) {
                    const double2 pp = *reinterpret_cast<const double2 *>(s_pv + (r0 + lane) * 4);
                    const double2 vv = *reinterpret_cast<const double2 *>(s_pv + (r0 + lane) * 4 + 2);
                    float *o = buf + lane * D;
                    o[0] = (float)vv.x; o[1] = (float)vv.y; o[2] = (float)pp.x; o[3] = (float)pp.y;
                }
                if (N > 1) {
#pragma unroll
                    for (int t = 0; t < S::HPL; ++t) {
                        const int q = lane + 32 * t;
                        if (q < S::HP) {
                            const int il = q / (N - 1);
                            const int ti = q - il * (N - 1);
                            const int i = r0 + il;
                            const int k = ti + (ti >= i ? 1 : 0);
                            const double2 A = *reinterpret_cast<const double2 *>(s_pv + k * 4);
                            const double2 B = *reinterpret_cast<const double2 *>(s_pv + i * 4);
                            float *o = buf + il * D + 4 + 2 * ti;
                            o[0] = (float)dsub(A.x, B.x);
                            o[1] = (float)dsub(A.y, B.y);
                        }
                    }
                }
                // PoI sections: [q_j - p_i, energy_j, (m_energy prefilled), done_j]
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double2 pi = *reinterpret_cast<const double2 *>(s_pv + (r0 + r) * 4);
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) {
                        const int j = lane + 32 * s;
                        if ((M % 32 == 0) || (j < M)) {
                            float *o = buf + r * D + H + 5 * j;
                            o[0] = (float)dsub(qx[s], pi.x);
                            o[1] = (float)dsub(qy[s], pi.y);
                            o[2] = fe[s];
                            o[4] = fd[s];
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    bulk_store_s2g(p.obs + ((size_t)e * N + r0) * D, buf, (uint32_t)(S::STAGE_FLOATS * 4));
                    bulk_commit();
                }
                bufsel ^= (S::NBUF - 1);
            }
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// launch table of the specialisations
typedef void (*env_kernel_fn)(const EnvKParams);
struct EnvSpecEntry {
    int N, M, smem, wpc, min_blocks;
    env_kernel_fn step, reset;
};
template <int N, int M>
constexpr EnvSpecEntry make_spec_entry() {
    return EnvSpecEntry{N, M, EnvSpec<N, M>::SMEM, EnvSpec<N, M>::WPC, EnvSpec<N, M>::MIN_BLOCKS,
                        dcc_env_spec_kernel<N, M, true>, dcc_env_spec_kernel<N, M, false>};
}
static const EnvSpecEntry g_env_specs[] = {
    make_spec_entry<4, 20>(),    // shipped default (dcc.yaml:5-6)
    make_spec_entry<8, 64>(),    // BASELINE configs[1], [3], [4]
    make_spec_entry<16, 256>(),  // BASELINE configs[2]
};

// ---- host side -----------------------------------------------------------------------------------

// smallest s with sqrt(s) >= t:  sqrt(x) < t  <=>  x < s   (sqrt correctly rounded, monotone)
static double sq_lt_bound(double t) {
    if (!(t > 0)) return 0.0;
    double s = t * t;
    while (sqrt(s) >= t && s > 0) s = nextafter(s, 0.0);
    while (sqrt(s) < t) s = nextafter(s, INFINITY);
    return s;
}
// largest s with sqrt(s) <= r:  sqrt(x) <= r  <=>  x <= s  and  sqrt(x) > r  <=>  x > s
static double sq_le_bound(double r) {
    if (!(r >= 0)) return -1.0;
    double s = r * r;
    while (sqrt(s) <= r) s = nextafter(s, INFINITY);
    while (sqrt(s) > r) s = nextafter(s, 0.0);
    return s;
}

struct EnvHandle {
    uint32_t magic;
    dcc_env_cfg cfg;
    int device;
    int sm_count;
    int D;
    double world_comm_r_scale, world_contact_force;
    double *d_poi;
    double *d_poi_env;   // [E, M, 2] per-env layouts (dcc_env_set_poi_layouts) or nullptr
    double *d_pos_vel;
    uint8_t *d_energy;
    EnvKParams kp;
    int warps_per_cta, ctas_override;
    int smem_bytes, ctas_step, ctas_reset;
    const EnvSpecEntry *spec;   // compile-time specialisation for this (N, M), or NULL
    int spec_enabled, spec_ctas;
    int64_t launches;
    // device staging for the *_host entry points
    float *hs_actions, *hs_obs, *hs_rew, *hs_cov;
    uint8_t *hs_done;
};
constexpr uint32_t ENV_MAGIC = 0xDCCE0001u;

static EnvHandle *as_env(void *h) {
    EnvHandle *e = static_cast<EnvHandle *>(h);
    return (e && e->magic == ENV_MAGIC) ? e : nullptr;
}

static int configure_launch(EnvHandle *h) {
    EnvKParams &k = h->kp;
    const int wpc = h->warps_per_cta;
    k.poi_bytes = (int)align_up((size_t)k.M * 16, 128);
    const size_t stage_bytes = align_up((size_t)k.stage_floats * 4, 128);
    k.stage_stride = (int)(stage_bytes / 4);
    k.pw_bytes = (int)align_up((size_t)k.N * 32 + align_up((size_t)k.M, 16) + stage_bytes * k.n_buf, 128);
    h->smem_bytes = k.poi_bytes + k.pw_bytes * wpc;
    if (h->smem_bytes > 227 * 1024) return DCC_ERR_UNSUPPORTED;
    {   // the attribute is per device and shared by every handle on it: only ever RAISE it, so that a later handle with
        // a smaller footprint cannot pull it under an earlier live handle that needs more than 48 KB
        static std::mutex mu;
        static int max_smem[64] = {0};
        std::lock_guard<std::mutex> lk(mu);
        int &mx = max_smem[h->device & 63];
        if (h->smem_bytes > mx) {
            DCC_CUDA_TRY(cudaFuncSetAttribute(dcc_env_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
            DCC_CUDA_TRY(cudaFuncSetAttribute(dcc_env_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
            mx = h->smem_bytes;
        }
    }
    int occ_step = 0, occ_reset = 0;
    DCC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_step, dcc_env_kernel<true>, wpc * 32, h->smem_bytes));
    DCC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_reset, dcc_env_kernel<false>, wpc * 32, h->smem_bytes));
    if (occ_step < 1 || occ_reset < 1) return DCC_ERR_UNSUPPORTED;
    const int need = (k.E + wpc - 1) / wpc;
    auto pick = [&](int occ) {
        int c = h->ctas_override > 0 ? h->ctas_override : h->sm_count * occ;
        if (c > need) c = need;
        return c < 1 ? 1 : c;
    };
    h->ctas_step = pick(occ_step);
    h->ctas_reset = pick(occ_reset);
    if (h->spec) {
        const EnvSpecEntry *sp = h->spec;
        DCC_CUDA_TRY(cudaFuncSetAttribute(sp->step, cudaFuncAttributeMaxDynamicSharedMemorySize, sp->smem));
        DCC_CUDA_TRY(cudaFuncSetAttribute(sp->reset, cudaFuncAttributeMaxDynamicSharedMemorySize, sp->smem));
        int occ = 0;
        DCC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sp->step, sp->wpc * 32, sp->smem));
        if (occ < 1) { h->spec = nullptr; return DCC_OK; }
        // default: one env per warp (non-persistent grid).  The hardware block scheduler then balances the tail
        // dynamically; measured 125.5 us vs 134.7-140.8 us per step for persistent grids at 8/64/65536.
        const int need_s = (k.E + sp->wpc - 1) / sp->wpc;
        int c = h->ctas_override > 0 ? h->ctas_override : need_s;
        if (c > need_s) c = need_s;
        h->spec_ctas = c < 1 ? 1 : c;
    }
    return DCC_OK;
}

}  // namespace dcc

using namespace dcc;

extern "C" {

int dcc_env_cfg_default(dcc_env_cfg *c) {
    if (!c) return DCC_ERR_INVALID_ARG;
    memset(c, 0, sizeof *c);
    c->n_envs = 16; c->n_agents = 4; c->n_pois = 20; c->max_ep_len = 150; c->reference_compat = 1;
    c->r_cover = 0.2; c->r_comm = 0.4; c->comm_r_scale = 0.95; c->comm_force_scale = 0.0;
    c->dt = 0.1; c->damping = 0.25; c->max_speed = 0.5; c->sensitivity = 5.0; c->m_energy = 5.0;
    c->rew_cover = 75.0; c->rew_done = 1500.0; c->rew_out = -100.0; c->contact_margin = 1e-3;
    return DCC_OK;
}

int dcc_env_obs_dim(int32_t n_agents, int32_t n_pois) { return 4 + 2 * (n_agents - 1) + 5 * n_pois; }

int dcc_env_create(const dcc_env_cfg *cfg, const double *h_poi_xy, int device, void **handle) {
    if (!cfg || !h_poi_xy || !handle) return DCC_ERR_INVALID_ARG;
    *handle = nullptr;
    if (cfg->n_envs < 1 || cfg->n_agents < 1 || cfg->n_agents > DCC_MAX_AGENTS || cfg->n_pois < 1 ||
        cfg->n_pois > 4096)
        return DCC_ERR_INVALID_ARG;
    if (!(cfg->r_cover >= 0) || !(cfg->r_comm > 0) || !(cfg->max_speed > 0) || !(cfg->m_energy > 0) ||
        cfg->m_energy > 200.0 || !(cfg->contact_margin > 0))
        return DCC_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return DCC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(device);
    cudaDeviceProp prop;
    DCC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return DCC_ERR_NO_DEVICE;  // sm_100a only: no fallback path exists

    EnvHandle *h = new (std::nothrow) EnvHandle();
    if (!h) return DCC_ERR_ALLOC;
    memset(h, 0, sizeof *h);
    h->magic = ENV_MAGIC;
    h->cfg = *cfg;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    const int N = cfg->n_agents, M = cfg->n_pois, E = cfg->n_envs;
    h->D = dcc_env_obs_dim(N, M);
    // the shipped scenario never forwards its comm args to the world
    h->world_comm_r_scale = cfg->reference_compat ? 0.9 : cfg->comm_r_scale;
    h->world_contact_force = cfg->reference_compat ? 0.0 : 1e2 * cfg->comm_force_scale;
    if (!(h->world_comm_r_scale > 0)) { delete h; return DCC_ERR_UNSUPPORTED; }  // update_connect would be skipped (:59)

    EnvKParams &k = h->kp;
    k.E = E; k.N = N; k.M = M; k.D = h->D; k.H = 2 * N + 2;
    k.slots = (M + 31) / 32;
    k.e_thr = (int)ceil(cfg->m_energy);
    k.force_on = h->world_contact_force > 0 ? 1 : 0;
    k.m_energy_f = (float)cfg->m_energy;
    k.sens = (float)cfg->sensitivity;
    k.dt32 = (float)cfg->dt;
    const double two_r = cfg->r_comm + cfg->r_comm;               // agent_a.r_comm + agent_b.r_comm (:77)
    k.thr2_adj = sq_lt_bound(two_r);
    k.thr2_adjs = sq_lt_bound(h->world_comm_r_scale * two_r);     // (:79)
    k.cover2 = sq_le_bound(cfg->r_cover);                         // (:165)
    k.speed2 = sq_le_bound(cfg->max_speed);                       // (:149-150)
    k.keep = 1 - cfg->damping;
    k.dt = cfg->dt;
    k.max_speed = cfg->max_speed;
    k.lim_force = h->world_comm_r_scale * 2 * cfg->r_comm;        // (:120)
    k.dist_max = two_r * h->world_comm_r_scale;                   // (:134)
    k.margin = cfg->contact_margin;
    k.contact_force = h->world_contact_force;
    k.rew_cover = cfg->rew_cover; k.rew_done = cfg->rew_done; k.rew_out = cfg->rew_out;

    // observation staging: chunks of R rows whose byte size is a multiple of 16 (cp.async.bulk granularity)
    const size_t row_bytes = (size_t)h->D * 4;
    k.use_bulk = env_block_bulk_ok(N, row_bytes) ? 1 : 0;
    // Small chunks through ONE stage buffer per warp, as in the specialisations: resident warps beat a second buffer.
    // Measured with the runtime-shape kernel (µs per step, 12 KB chunks + 2 buffers -> 6000 B chunks + 1 buffer):
    // 8/64 x 65 536: 254 -> 199; 16/256 x 32 768: 906 -> 572; 6/41 x 65 536: 196 -> 196 (one chunk either way).
    size_t budget = SPEC_STAGE_BUDGET;
    int nbuf_multi = 1;
    if (const char *e = getenv("DCC_ENV_GENERIC_BUDGET")) budget = (size_t)atoi(e);     // tuning knobs (tools/env_generic_probe.py)
    if (const char *e = getenv("DCC_ENV_GENERIC_NBUF")) nbuf_multi = atoi(e) == 2 ? 2 : 1;
    const int R = rows_per_chunk(N, row_bytes, budget);
    k.rows_per_chunk = R;
    k.n_chunks = N / R;
    k.n_buf = (k.use_bulk && k.n_chunks > 1) ? nbuf_multi : 1;
    k.stage_floats = R * h->D;

    h->spec = nullptr;
    h->spec_enabled = 1;
    for (const EnvSpecEntry &sp : g_env_specs)
        if (sp.N == N && sp.M == M) h->spec = &sp;
    h->warps_per_cta = 4;
    h->ctas_override = 0;
    int rc = DCC_OK;
    for (;;) {
        rc = configure_launch(h);
        if (rc == DCC_ERR_UNSUPPORTED && h->warps_per_cta > 1) { h->warps_per_cta >>= 1; continue; }
        break;
    }
    if (rc != DCC_OK) { delete h; return rc; }

    cudaError_t ce;
    if ((ce = cudaMalloc(&h->d_poi, sizeof(double) * 2 * M)) != cudaSuccess ||
        (ce = cudaMalloc(&h->d_pos_vel, sizeof(double) * 4 * (size_t)N * E)) != cudaSuccess ||
        (ce = cudaMalloc(&h->d_energy, (size_t)M * E)) != cudaSuccess) {
        set_last_cuda_error(ce, "cudaMalloc(env state)", __FILE__, __LINE__);
        dcc_env_destroy(h);
        return DCC_ERR_ALLOC;
    }
    DCC_CUDA_TRY(cudaMemcpy(h->d_poi, h_poi_xy, sizeof(double) * 2 * M, cudaMemcpyHostToDevice));
    DCC_CUDA_TRY(cudaMemset(h->d_pos_vel, 0, sizeof(double) * 4 * (size_t)N * E));
    DCC_CUDA_TRY(cudaMemset(h->d_energy, 0, (size_t)M * E));
    k.poi = h->d_poi; k.pos_vel = h->d_pos_vel; k.energy = h->d_energy;
    *handle = h;
    return DCC_OK;
}

int dcc_env_set_poi_layouts(void *handle, const double *h_poi_xy, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    EnvKParams &k = h->kp;
    if (!h_poi_xy) {                     // back to the shared layout given at creation
        k.poi_env = nullptr;
        return DCC_OK;
    }
    const size_t bytes = sizeof(double) * 2 * (size_t)k.M * k.E;
    if (!h->d_poi_env) {
        cudaError_t ce = cudaMalloc(&h->d_poi_env, bytes);
        if (ce != cudaSuccess) { set_last_cuda_error(ce, "cudaMalloc(per-env PoI layouts)", __FILE__, __LINE__); return DCC_ERR_ALLOC; }
    }
    DCC_CUDA_TRY(cudaMemcpyAsync(h->d_poi_env, h_poi_xy, bytes, cudaMemcpyHostToDevice, s));
    DCC_CUDA_TRY(cudaStreamSynchronize(s));
    k.poi_env = h->d_poi_env;
    return DCC_OK;
}

int dcc_env_destroy(void *handle) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaFree(h->d_poi); cudaFree(h->d_poi_env); cudaFree(h->d_pos_vel); cudaFree(h->d_energy);
    cudaFree(h->hs_actions); cudaFree(h->hs_obs); cudaFree(h->hs_rew); cudaFree(h->hs_cov); cudaFree(h->hs_done);
    h->magic = 0;
    delete h;
    return DCC_OK;
}

int dcc_env_set_launch(void *handle, int warps_per_cta, int ctas) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    if (warps_per_cta != 1 && warps_per_cta != 2 && warps_per_cta != 4 && warps_per_cta != 8 && warps_per_cta != 16)
        return DCC_ERR_INVALID_ARG;
    const int old_w = h->warps_per_cta, old_c = h->ctas_override;
    h->warps_per_cta = warps_per_cta;
    h->ctas_override = ctas;
    const int rc = configure_launch(h);
    if (rc != DCC_OK) { h->warps_per_cta = old_w; h->ctas_override = old_c; configure_launch(h); }
    return rc;
}

int dcc_env_use_specialized(void *handle, int enable) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    h->spec_enabled = enable ? 1 : 0;
    return (h->spec && h->spec_enabled) ? 1 : 0;
}

int64_t dcc_env_launch_count(void *handle) {
    EnvHandle *h = as_env(handle);
    return h ? h->launches : -1;
}

int dcc_env_reset(void *handle, float *d_obs, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    EnvKParams k = h->kp;
    k.actions = nullptr; k.obs = d_obs; k.rew = nullptr; k.done = nullptr; k.cov = nullptr; k.connect = nullptr;
    k.adj = nullptr; k.adjs = nullptr;
    const bool aligned = !(d_obs && (reinterpret_cast<uintptr_t>(d_obs) & 15));
    if (!aligned) k.use_bulk = 0, k.n_buf = 1;
    if (h->spec && h->spec_enabled && aligned)
        h->spec->reset<<<h->spec_ctas, h->spec->wpc * 32, h->spec->smem, static_cast<cudaStream_t>(stream)>>>(k);
    else
        dcc_env_kernel<false><<<h->ctas_reset, h->warps_per_cta * 32, h->smem_bytes, static_cast<cudaStream_t>(stream)>>>(k);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

int dcc_env_step(void *handle, const float *d_actions, float *d_obs, float *d_rew, uint8_t *d_done,
                 float *d_coverage, uint8_t *d_connect, uint32_t *d_adj, uint32_t *d_adj_s, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h || !d_actions) return DCC_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(d_actions) & 7) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    EnvKParams k = h->kp;
    k.actions = d_actions; k.obs = d_obs; k.rew = d_rew; k.done = d_done; k.cov = d_coverage; k.connect = d_connect;
    k.adj = d_adj; k.adjs = d_adj_s;
    const bool aligned = !(d_obs && (reinterpret_cast<uintptr_t>(d_obs) & 15));
    if (!aligned) k.use_bulk = 0, k.n_buf = 1;
    if (h->spec && h->spec_enabled && aligned)
        h->spec->step<<<h->spec_ctas, h->spec->wpc * 32, h->spec->smem, static_cast<cudaStream_t>(stream)>>>(k);
    else
        dcc_env_kernel<true><<<h->ctas_step, h->warps_per_cta * 32, h->smem_bytes, static_cast<cudaStream_t>(stream)>>>(k);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

static int ensure_host_staging(EnvHandle *h) {
    if (h->hs_actions) return DCC_OK;
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents;
    DCC_CUDA_TRY(cudaMalloc(&h->hs_actions, E * N * 2 * sizeof(float)));
    DCC_CUDA_TRY(cudaMalloc(&h->hs_obs, E * N * (size_t)h->D * sizeof(float)));
    DCC_CUDA_TRY(cudaMalloc(&h->hs_rew, E * N * sizeof(float)));
    DCC_CUDA_TRY(cudaMalloc(&h->hs_done, E * N));
    DCC_CUDA_TRY(cudaMalloc(&h->hs_cov, E * sizeof(float)));
    return DCC_OK;
}

int dcc_env_step_host(void *handle, const float *h_actions, float *h_obs, float *h_rew, uint8_t *h_done,
                      float *h_coverage, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h || !h_actions) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    int rc = ensure_host_staging(h);
    if (rc != DCC_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents;
    DCC_CUDA_TRY(cudaMemcpyAsync(h->hs_actions, h_actions, E * N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = dcc_env_step(h, h->hs_actions, h_obs ? h->hs_obs : nullptr, h_rew ? h->hs_rew : nullptr,
                      h_done ? h->hs_done : nullptr, h_coverage ? h->hs_cov : nullptr, nullptr, nullptr, nullptr, stream);
    if (rc != DCC_OK) return rc;
    if (h_obs) DCC_CUDA_TRY(cudaMemcpyAsync(h_obs, h->hs_obs, E * N * (size_t)h->D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_rew) DCC_CUDA_TRY(cudaMemcpyAsync(h_rew, h->hs_rew, E * N * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_done) DCC_CUDA_TRY(cudaMemcpyAsync(h_done, h->hs_done, E * N, cudaMemcpyDeviceToHost, s));
    if (h_coverage) DCC_CUDA_TRY(cudaMemcpyAsync(h_coverage, h->hs_cov, E * sizeof(float), cudaMemcpyDeviceToHost, s));
    DCC_CUDA_TRY(cudaStreamSynchronize(s));
    return DCC_OK;
}

int dcc_env_reset_host(void *handle, float *h_obs, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    int rc = ensure_host_staging(h);
    if (rc != DCC_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = dcc_env_reset(h, h_obs ? h->hs_obs : nullptr, stream);
    if (rc != DCC_OK) return rc;
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents;
    if (h_obs) DCC_CUDA_TRY(cudaMemcpyAsync(h_obs, h->hs_obs, E * N * (size_t)h->D * sizeof(float), cudaMemcpyDeviceToHost, s));
    DCC_CUDA_TRY(cudaStreamSynchronize(s));
    return DCC_OK;
}

int dcc_env_get_state(void *handle, double *h_pos_vel, uint8_t *h_energy, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents, M = h->cfg.n_pois;
    if (h_pos_vel) DCC_CUDA_TRY(cudaMemcpyAsync(h_pos_vel, h->d_pos_vel, sizeof(double) * 4 * N * E, cudaMemcpyDeviceToHost, s));
    if (h_energy) DCC_CUDA_TRY(cudaMemcpyAsync(h_energy, h->d_energy, M * E, cudaMemcpyDeviceToHost, s));
    DCC_CUDA_TRY(cudaStreamSynchronize(s));
    return DCC_OK;
}

int dcc_env_set_state(void *handle, const double *h_pos_vel, const uint8_t *h_energy, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents, M = h->cfg.n_pois;
    if (h_pos_vel) DCC_CUDA_TRY(cudaMemcpyAsync(h->d_pos_vel, h_pos_vel, sizeof(double) * 4 * N * E, cudaMemcpyHostToDevice, s));
    if (h_energy) DCC_CUDA_TRY(cudaMemcpyAsync(h->d_energy, h_energy, M * E, cudaMemcpyHostToDevice, s));
    DCC_CUDA_TRY(cudaStreamSynchronize(s));
    return DCC_OK;
}

int dcc_env_snapshot_state(void *handle, double *d_pos_vel_out, uint8_t *d_energy_out, dcc_stream_t stream) {
    EnvHandle *h = as_env(handle);
    if (!h || !d_pos_vel_out || !d_energy_out) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t E = h->cfg.n_envs, N = h->cfg.n_agents, M = h->cfg.n_pois;
    DCC_CUDA_TRY(cudaMemcpyAsync(d_pos_vel_out, h->d_pos_vel, sizeof(double) * 4 * N * E, cudaMemcpyDeviceToDevice, s));
    DCC_CUDA_TRY(cudaMemcpyAsync(d_energy_out, h->d_energy, M * E, cudaMemcpyDeviceToDevice, s));
    return DCC_OK;
}

int dcc_env_state_ptrs(void *handle, double **d_pos_vel, uint8_t **d_energy) {
    EnvHandle *h = as_env(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    if (d_pos_vel) *d_pos_vel = h->d_pos_vel;
    if (d_energy) *d_energy = h->d_energy;
    return DCC_OK;
}

}  // extern "C"

// dcc_compact.cuh — the learner's first layer evaluated from the env's COMPACT STATE instead of observation rows
// ("compact-state storage + obs regeneration").
//
// The reference's observation row of agent i (envs/mpe/multiagent/scenarios/coverage.py:99-110)
//     x = [ v_i, p_i, p_k - p_i (k != i),  then per PoI j:  q_j - p_i, e_j, 5.0, [e_j >= 5] ]          (D = 2N + 2 + 5M)
// and the critic's centralised input (the env's N rows concatenated, learner.py:219-220) are AFFINE in the compact
// state (UAV positions / velocities, PoI energies: 32 N + M bytes per env step instead of 4 N D).  With the input
// LayerNorm (algos/algo_utils/mlp.py:44-58) written as xhat = rstd * x - rstd * mean, xhat is LINEAR in
//     f = rstd * [ own_0 .. own_{nb-1},  e_1..e_M,  d_1..d_M,  1,  -mean ],      own_i = x_i[0 : 2N + 2]
// (nb = 1 for an actor row, N for a critic row):  xhat = A f  with a constant sparse A built from the PoI table.  So
//     xhat (W1 * gamma0)^T = f Wt^T,  Wt = (W1 * gamma0) A          (fold_compact_kernel, once per optimiser step)
//     G = dz1^T xhat       = (dz1^T f) A^T                           (unfold_compact_grad_kernel, once per optimiser step)
// exactly (oracle/compact_oracle.py pins the identities in float64).  The layer-1 GEMMs then reduce over
// K = 2N + 4 + 2M (148 at 8/64, was 338) for the actor and N (2N + 2) + 2M + 2 (274, was 2704) for the critic, the rollout
// stores 320 B per env step instead of 10.8 KB, and no observation row is ever read back during the update.
// The own-block values are float32 of the float64 state expressions, exactly as the env kernel writes them; the
// LayerNorm statistics come from closed forms of the same state (see compact_features_kernel).
#pragma once
#include <cuda_fp16.h>

#include "dcc_ops.cuh"

namespace dcc {

// rstd = 1 / sqrt(v) for the LayerNorm statistics: rsqrtf + one Newton step (<= 1 ulp of the float result) instead of a float64
// sqrt and division (two ~170-instruction sequences per env step in an issue-bound kernel)
__device__ __forceinline__ float rstd_from_var(double var_plus_eps) {
    const float v = (float)var_plus_eps;
    float r = rsqrtf(v);
    r = r * fmaf(-0.5f * v, r * r, 1.5f);
    return r;
}

struct CompactDims {
    int N, M, D, OWN;      // OWN = 2N + 2: the [v_i, p_i, p_k - p_i] head of an observation row
    int Ka, Kc;            // feature counts: OWN + 2M + 2, N * OWN + 2M + 2
    int lda, ldc;          // padded leading dimensions (multiples of 32; pads are zero)
    float m_energy;        // the constant 5.0 column and the done threshold (coverage.py:107-109)
    int e_thr;             // done_j = energy_j >= e_thr
    double Qx, Qy, Qxx, Qyy;   // sums of q_jx, q_jy, q_jx^2, q_jy^2 over the PoI table (closed-form LayerNorm statistics)
    __host__ void set_poi(const double *poi) {
        Qx = Qy = Qxx = Qyy = 0.0;
        for (int j = 0; j < M; ++j) { Qx += poi[2 * j]; Qy += poi[2 * j + 1]; Qxx += poi[2 * j] * poi[2 * j]; Qyy += poi[2 * j + 1] * poi[2 * j + 1]; }
    }
    __host__ void init(int n, int m, double me) {
        N = n; M = m; D = 4 + 2 * (n - 1) + 5 * m; OWN = 2 * n + 2;
        Ka = OWN + 2 * m + 2; Kc = n * OWN + 2 * m + 2;
        lda = (Ka + 31) / 32 * 32; ldc = (Kc + 31) / 32 * 32;
        m_energy = (float)me; e_thr = (int)ceil(me);
    }
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// shared memory of compact_features_kernel for `warps` warps per block (layout in the kernel's header comment)
__host__ __device__ static inline size_t compact_features_smem(const CompactDims &cd, int warps) {
    const int ownp = (cd.OWN + 7) & ~7;
    const size_t per = (size_t)cd.N * 16 + (size_t)(cd.ldc + cd.lda + cd.N * ownp + 2 * cd.N) * 4;
    return ((per + 15) & ~(size_t)15) * warps;
}

// One warp per env-step row; lane i < N owns UAV i.  Shared memory per warp (compact_features_smem):
//   s_p   [N][2] float64   positions
//   rowC  [ldc]            the UN-scaled critic row: own blocks of all agents ([v_i, p_i, p_k - p_i] as float32 of the float64
//                          expressions, coverage.py:100-105), then the shared tail [e_1..e_M, d_1..d_M, 1, 0 (slot of -mean), 0 pad]
//   rowT  [lda]            the same tail at the ACTOR row's column positions (entries [OWN, lda); [0, OWN) unused)
//   O     [N][OWNP]        per agent the first OWNP = round_up(OWN, 8) columns of its actor row (own block + head of the tail)
//   stat  [2N]             per-agent (mean, rstd)
// so that every 8-column group of an output row is two aligned LDS.128 from ONE base (O_i, rowT or rowC), eight multiplies by
// the row's rstd, one fp16 hi/lo split and two 16-byte stores.  Measured (ncu, 37 888 env steps, both outputs): 77 us, 63 M warp
// instructions = 1 670 per env step, issue-bound (75 % issue utilisation) — 5 x 102 instructions in the actor output loop, ~350
// in the two float64 1/sqrt sequences, ~150 in the neighbour loop; the element-wise select version before it took the same time.
//
// LayerNorm statistics: the row's sum and sum of squares have CLOSED FORMS in the state — sum_j (q_jx - p_ix) =
// Qx - M p_ix, sum_j (q_jx - p_ix)^2 = Qxx - 2 p_ix Qx + M p_ix^2 (Qx .. Qyy: constants of the PoI table, CompactDims),
// likewise over the other UAVs, plus integer sums of e_j, e_j^2, d_j — evaluated in float64 per agent lane, so no pass
// over the N * M relative positions is needed and var = E[x^2] - mean^2 has no cancellation problem.  They differ from
// the statistics of the float32-rounded observation values by ~1e-8 relative (each value moves by <= 2^-24 relative).
// Outputs (either may be NULL): Fa [rows * N, lda] actor features, Fc [rows, ldc] critic features.
__global__ void __launch_bounds__(256) compact_features_kernel(const double *__restrict__ pos_vel, const uint8_t *__restrict__ energy,
                                                               float *__restrict__ Fa, float *__restrict__ Fc, int rows,
                                                               CompactDims cd, int normalize, size_t lo_a, size_t lo_c,
                                                               const long long *__restrict__ ridx = nullptr) {
    // ridx (optional, minibatch path: feed_forward_generator draws AGENT rows, shared_buffer.py:238-262): output row k is
    // built from state row ridx[k] / N; the actor row is that of agent ridx[k] % N alone (Fa [rows, lda]) and the critic row
    // the centralised row of that env step (Fc [rows, ldc], one per sampled agent row, as the reference evaluates it).
    // lo_a / lo_c != 0: the rows leave PRE-SPLIT as fp16 hi | lo (hi halves at Fa / Fc, lo halves lo_a / lo_c halves behind
    // them, row pitch lda / ldc halves) for the TMA-fed GEMMs (tc::TcfParams::a_split); 0: plain float32 rows
    extern __shared__ __align__(16) unsigned char cf_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = cd.N, M = cd.M, D = cd.D, OWN = cd.OWN;
    const int nown = N * OWN, OWNP = (OWN + 7) & ~7;
    const size_t per_warp = compact_features_smem(cd, 1);
    unsigned char *base = cf_smem + per_warp * wib;
    double *s_p = reinterpret_cast<double *>(base);            // [N][2] positions
    float *rowC = reinterpret_cast<float *>(base + (size_t)N * 16);
    float *rowT = rowC + cd.ldc;
    float *O = rowT + cd.lda;
    float *s_stat = O + N * OWNP;                              // [N] mean, [N] rstd
    // constant part of the tail in both copies: 1 at 2M, zeros behind it (the slot of -mean is filled per output row)
    for (int c = nown + 2 * M + lane; c < cd.ldc; c += 32) rowC[c] = (c == nown + 2 * M) ? 1.f : 0.f;
    for (int c = OWN + 2 * M + lane; c < cd.lda; c += 32) rowT[c] = (c == OWN + 2 * M) ? 1.f : 0.f;
    const int nv = cd.lda >> 3;                                // 8-column groups of an actor row
    const int q_il0 = lane / nv, q_g0 = lane - q_il0 * nv, q_dil = 32 / nv, q_dg = 32 - q_dil * nv;   // walk of q = lane + 32 it
    const int cm_a = OWN + 2 * M + 1, cm_c = nown + 2 * M + 1;   // column of -mean * rstd
    for (int ro = blockIdx.x * wpb + wib; ro < rows; ro += gridDim.x * wpb) {
        __syncwarp();
        // r = state row read; ro = output row; ag = the one agent whose actor row is wanted (-1: all N, whole-rollout path)
        const size_t r = ridx ? (size_t)(ridx[ro] / N) : (size_t)ro;
        const int ag = ridx ? (int)(ridx[ro] % N) : -1;
        double px = 0.0, py = 0.0, vx = 0.0, vy = 0.0;
        if (lane < N) {
            const double2 pp = *reinterpret_cast<const double2 *>(pos_vel + ((size_t)r * N + lane) * 4);
            const double2 vv = *reinterpret_cast<const double2 *>(pos_vel + ((size_t)r * N + lane) * 4 + 2);
            px = pp.x; py = pp.y; vx = vv.x; vy = vv.y;
            s_p[2 * lane] = px; s_p[2 * lane + 1] = py;
        }
        // PoI energies: tail values (both copies) + exact integer sums of e, e^2, d
        int se = 0, see = 0, sd = 0;
        for (int j = lane; j < M; j += 32) {
            const int e = energy[(size_t)r * M + j];
            const int d = e >= cd.e_thr ? 1 : 0;
            const float ef = (float)e, df = (float)d;
            rowC[nown + j] = ef; rowC[nown + M + j] = df;
            rowT[OWN + j] = ef; rowT[OWN + M + j] = df;
            se += e; see += e * e; sd += d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(FULL_MASK, se, o); see += __shfl_xor_sync(FULL_MASK, see, o); sd += __shfl_xor_sync(FULL_MASK, sd, o);
        }
        __syncwarp();
        double rs = 0.0, rq = 0.0;      // this agent's row sum / sum of squares
        if (lane < N) {
            float *own = rowC + lane * OWN;
            own[0] = (float)vx; own[1] = (float)vy; own[2] = (float)px; own[3] = (float)py;
            double s = vx + vy + px + py, q = vx * vx + vy * vy + px * px + py * py;
            for (int k = 0; k < N; ++k) {       // branch-free: k == lane contributes exact zeros to the sums and stores nothing
                const double dx = dsub(s_p[2 * k], px), dy = dsub(s_p[2 * k + 1], py);
                const int t = k - (k > lane ? 1 : 0);
                if (k != lane) { own[4 + 2 * t] = (float)dx; own[5 + 2 * t] = (float)dy; }
                s += dx + dy; q = fma(dx, dx, q); q = fma(dy, dy, q);
            }
            const double m = (double)M, me = (double)cd.m_energy;
            s += (cd.Qx - m * px) + (cd.Qy - m * py) + (double)se + m * me + (double)sd;
            q += (cd.Qxx - 2.0 * px * cd.Qx + m * px * px) + (cd.Qyy - 2.0 * py * cd.Qy + m * py * py) + (double)see + m * me * me + (double)sd;
            rs = s; rq = q;
            float mean = 0.f, rstd = 1.f;
            if (normalize) {
                const double mu = s / (double)D;
                const double var = fmax(q / (double)D - mu * mu, 0.0);
                mean = (float)mu;
                rstd = rstd_from_var(var + (double)LN_EPS);
            }
            s_stat[lane] = mean; s_stat[N + lane] = rstd;
        }
        __syncwarp();
        if (Fa) {
            // head of every agent's actor row: own block, then the first tail columns up to the next multiple of 8
            for (int idx = lane; idx < N * OWNP; idx += 32) {
                const int i = idx / OWNP, c = idx - i * OWNP;
                O[idx] = c < OWN ? rowC[i * OWN + c] : rowT[c];
            }
            __syncwarp();
            const int na = ag < 0 ? N : 1;                       // actor rows written for this output row
            int il = q_il0, g = q_g0;
            for (int q8 = lane; q8 < na * nv; q8 += 32) {
                const int c0 = g << 3;
                const int i = ag < 0 ? il : ag;
                const float mean = s_stat[i], rstd = s_stat[N + i];
                const float *src = c0 < OWNP ? O + i * OWNP + c0 : rowT + c0;
                const float4 a0 = *reinterpret_cast<const float4 *>(src), a1 = *reinterpret_cast<const float4 *>(src + 4);
                float v[8] = {a0.x * rstd, a0.y * rstd, a0.z * rstd, a0.w * rstd, a1.x * rstd, a1.y * rstd, a1.z * rstd, a1.w * rstd};
                if ((unsigned)(cm_a - c0) < 8u) {
                    const float nm = -mean * rstd;
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = (e == cm_a - c0) ? nm : v[e];
                }
                const size_t orow = (ag < 0 ? (size_t)ro * N + i : (size_t)ro) * cd.lda + c0;
                if (lo_a) {
                    uint4 hi, lo;
                    tc_split_pair(v[0], v[1], hi.x, lo.x); tc_split_pair(v[2], v[3], hi.y, lo.y);
                    tc_split_pair(v[4], v[5], hi.z, lo.z); tc_split_pair(v[6], v[7], hi.w, lo.w);
                    __half *hp = reinterpret_cast<__half *>(Fa) + orow;
                    *reinterpret_cast<uint4 *>(hp) = hi;
                    *reinterpret_cast<uint4 *>(hp + lo_a) = lo;
                } else {
                    *reinterpret_cast<float4 *>(Fa + orow) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(Fa + orow + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                g += q_dg; il += q_dil;
                if (g >= nv) { g -= nv; ++il; }
            }
        }
        if (Fc) {
            float mean_c = 0.f, rstd_c = 1.f;
            if (normalize) {
                const double S = warp_sum_d(rs), Q = warp_sum_d(rq);       // lanes >= N contribute 0
                const double mu = S / ((double)N * D);
                const double var = fmax(Q / ((double)N * D) - mu * mu, 0.0);
                mean_c = (float)mu;
                rstd_c = rstd_from_var(var + (double)LN_EPS);
            }
            const float nm = -mean_c * rstd_c;
            for (int c0 = lane << 3; c0 < cd.ldc; c0 += 256) {
                const float4 a0 = *reinterpret_cast<const float4 *>(rowC + c0), a1 = *reinterpret_cast<const float4 *>(rowC + c0 + 4);
                float v[8] = {a0.x * rstd_c, a0.y * rstd_c, a0.z * rstd_c, a0.w * rstd_c, a1.x * rstd_c, a1.y * rstd_c, a1.z * rstd_c,
                              a1.w * rstd_c};
                if ((unsigned)(cm_c - c0) < 8u) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = (e == cm_c - c0) ? nm : v[e];
                }
                if (lo_c) {
                    uint4 hi, lo;
                    tc_split_pair(v[0], v[1], hi.x, lo.x); tc_split_pair(v[2], v[3], hi.y, lo.y);
                    tc_split_pair(v[4], v[5], hi.z, lo.z); tc_split_pair(v[6], v[7], hi.w, lo.w);
                    __half *hp = reinterpret_cast<__half *>(Fc) + (size_t)ro * cd.ldc + c0;
                    *reinterpret_cast<uint4 *>(hp) = hi;
                    *reinterpret_cast<uint4 *>(hp + lo_c) = lo;
                } else {
                    float *out = Fc + (size_t)ro * cd.ldc + c0;
                    *reinterpret_cast<float4 *>(out) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(out + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
    }
}

// Wt[h, :] (ld = ldk, zero padded) and b1g[h] from fc1 (W1 [H, nb*D], b1) and the input LayerNorm affine (g0, be0 or
// NULL): Wg = W1 * g0, b1g = b1 + W1 be0, Wt = Wg A (see the file header; oracle/compact_oracle.py::fold_weights).
// One warp per output unit h; sums in float64.
__global__ void fold_compact_kernel(const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ g0,
                                    const float *__restrict__ be0, const double *__restrict__ poi, float *__restrict__ Wt,
                                    float *__restrict__ b1g, int H, CompactDims cd, int nb, int ldk) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= H) return;
    const int M = cd.M, D = cd.D, OWN = cd.OWN, F = nb * D;
    const float *w = W1 + (size_t)h * F;
    float *out = Wt + (size_t)h * ldk;
    auto wg = [&](int c) -> double { return g0 ? (double)w[c] * (double)g0[c] : (double)w[c]; };
    double sb = 0.0, sall = 0.0;
    for (int c = lane; c < F; c += 32) {
        if (be0) sb += (double)w[c] * (double)be0[c];
        sall += wg(c);
    }
    sb = warp_sum_d(sb); sall = warp_sum_d(sall);
    double cst = 0.0;
    for (int i = 0; i < nb; ++i) {
        const int o = i * D + OWN;
        double sx = 0.0, sy = 0.0;
        for (int j = lane; j < M; j += 32) {
            const double wx = wg(o + 5 * j), wy = wg(o + 5 * j + 1);
            sx += wx; sy += wy;
            cst += wx * poi[2 * j] + wy * poi[2 * j + 1] + (double)cd.m_energy * wg(o + 5 * j + 3);
        }
        sx = warp_sum_d(sx); sy = warp_sum_d(sy);
        for (int k = lane; k < OWN; k += 32) {
            double v = wg(i * D + k);
            if (k == 2) v -= sx;
            if (k == 3) v -= sy;
            out[i * OWN + k] = (float)v;
        }
    }
    cst = warp_sum_d(cst);
    for (int j = lane; j < M; j += 32) {
        double se = 0.0, sd = 0.0;
        for (int i = 0; i < nb; ++i) { se += wg(i * D + OWN + 5 * j + 2); sd += wg(i * D + OWN + 5 * j + 4); }
        out[nb * OWN + j] = (float)se;
        out[nb * OWN + M + j] = (float)sd;
    }
    const int kc = nb * OWN + 2 * M;
    if (lane == 0) {
        out[kc] = (float)cst;
        out[kc + 1] = (float)sall;
        b1g[h] = (float)((double)b1[h] + sb);
    }
    for (int c = kc + 2 + lane; c < ldk; c += 32) out[c] = 0.f;
}

// G[h, c] (the fc1 weight-gradient slot, [H, nb*D]) = sum_f Gt[h, f] A[c, f]  (oracle/compact_oracle.py::unfold_grad).
// One thread per (h, c).
__global__ void unfold_compact_grad_kernel(const float *__restrict__ Gt, const double *__restrict__ poi, float *__restrict__ G,
                                           int H, CompactDims cd, int nb, int ldk) {
    const int M = cd.M, D = cd.D, OWN = cd.OWN, F = nb * D;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)H * F) return;
    const int h = (int)(idx / F), c = (int)(idx - (size_t)h * F);
    const float *g = Gt + (size_t)h * ldk;
    const int kc = nb * OWN + 2 * M;
    const double gc = g[kc], gm = g[kc + 1];
    const int i = c / D, k = c - i * D;
    double v;
    if (k < OWN) v = g[i * OWN + k];
    else {
        const int j = (k - OWN) / 5, t = (k - OWN) - 5 * j;
        if (t == 0) v = gc * poi[2 * j] - (double)g[i * OWN + 2];
        else if (t == 1) v = gc * poi[2 * j + 1] - (double)g[i * OWN + 3];
        else if (t == 2) v = g[nb * OWN + j];
        else if (t == 3) v = (double)cd.m_energy * gc;
        else v = g[nb * OWN + M + j];
    }
    G[idx] = (float)(v + gm);
}

// Observation rows regenerated from the compact state, bit-identical to what the env kernel writes
// (dcc_env.cu phase 6; coverage.py:99-110): the reference-shaped view of a compact rollout (`buffer.obs[t]`).
// One warp per env-step row.
__global__ void __launch_bounds__(256) obs_from_state_kernel(const double *__restrict__ pos_vel, const uint8_t *__restrict__ energy,
                                                             const double *__restrict__ poi, float *__restrict__ obs, int rows,
                                                             CompactDims cd) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int N = cd.N, M = cd.M, D = cd.D, OWN = cd.OWN;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const double *pv = pos_vel + (size_t)r * N * 4;
        const uint8_t *en = energy + (size_t)r * M;
        for (int i = 0; i < N; ++i) {
            float *o = obs + ((size_t)r * N + i) * D;
            const double pix = pv[i * 4], piy = pv[i * 4 + 1];
            for (int k = lane; k < OWN; k += 32) {
                float v;
                if (k < 2) v = (float)pv[i * 4 + 2 + k];
                else if (k < 4) v = (float)pv[i * 4 + (k - 2)];
                else {
                    const int t = (k - 4) >> 1, c = (k - 4) & 1;
                    const int ot = t + (t >= i ? 1 : 0);
                    v = (float)dsub(pv[ot * 4 + c], pv[i * 4 + c]);
                }
                o[k] = v;
            }
            for (int j = lane; j < M; j += 32) {
                const int e = en[j];
                float *p5 = o + OWN + 5 * j;
                p5[0] = (float)dsub(poi[2 * j], pix);
                p5[1] = (float)dsub(poi[2 * j + 1], piy);
                p5[2] = (float)e;
                p5[3] = cd.m_energy;
                p5[4] = (e >= cd.e_thr) ? 1.f : 0.f;
            }
        }
    }
}

}  // namespace dcc

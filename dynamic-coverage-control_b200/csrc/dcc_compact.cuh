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
// The LayerNorm statistics are taken over exactly the float32 observation values the reference normalises (rebuilt in
// registers from the float64 state the same way the env kernel builds them).
#pragma once
#include "dcc_ops.cuh"

namespace dcc {

struct CompactDims {
    int N, M, D, OWN;      // OWN = 2N + 2: the [v_i, p_i, p_k - p_i] head of an observation row
    int Ka, Kc;            // feature counts: OWN + 2M + 2, N * OWN + 2M + 2
    int lda, ldc;          // padded leading dimensions (multiples of 32; pads are zero)
    float m_energy;        // the constant 5.0 column and the done threshold (coverage.py:107-109)
    int e_thr;             // done_j = energy_j >= e_thr
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

// One warp per env-step row.  Shared memory per warp: N * 4 doubles (state) + (N * OWN + 2M) floats (un-normalised own
// blocks of all agents, then e_j, then d_j) + 2N floats (per-agent mean, rstd).
// Outputs (either may be NULL): Fa [rows * N, lda] actor features, Fc [rows, ldc] critic features.
// sidx (optional): output row r is built from state row sidx[r] (unused by the whole-rollout path).
__global__ void __launch_bounds__(256) compact_features_kernel(const double *__restrict__ pos_vel, const uint8_t *__restrict__ energy,
                                                               const double *__restrict__ poi, float *__restrict__ Fa,
                                                               float *__restrict__ Fc, int rows, CompactDims cd, int normalize) {
    extern __shared__ __align__(16) unsigned char cf_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = cd.N, M = cd.M, D = cd.D, OWN = cd.OWN;
    const int ush = N * OWN + 2 * M;                      // floats of the un-normalised feature stage
    const size_t per_warp = (size_t)N * 32 + align_up((size_t)(ush + 2 * N) * 4, 16);
    unsigned char *base = cf_smem + per_warp * wib;
    double *s_pv = reinterpret_cast<double *>(base);
    float *s_u = reinterpret_cast<float *>(base + (size_t)N * 32);
    float *s_stat = s_u + ush;                            // [N] mean, [N] rstd
    for (int r = blockIdx.x * wpb + wib; r < rows; r += gridDim.x * wpb) {
        __syncwarp();
        for (int i = lane; i < N * 4; i += 32) s_pv[i] = pos_vel[(size_t)r * N * 4 + i];
        for (int j = lane; j < M; j += 32) {
            const int e = energy[(size_t)r * M + j];
            s_u[N * OWN + j] = (float)e;
            s_u[N * OWN + M + j] = (e >= cd.e_thr) ? 1.f : 0.f;
        }
        __syncwarp();
        // own blocks: [v_i, p_i, p_k - p_i] as float32 of the float64 expressions (coverage.py:100-105)
        for (int q = lane; q < N * OWN; q += 32) {
            const int i = q / OWN, k = q - i * OWN;
            float v;
            if (k < 2) v = (float)s_pv[i * 4 + 2 + k];
            else if (k < 4) v = (float)s_pv[i * 4 + (k - 2)];
            else {
                const int t = (k - 4) >> 1, c = (k - 4) & 1;
                const int o = t + (t >= i ? 1 : 0);
                v = (float)dsub(s_pv[o * 4 + c], s_pv[i * 4 + c]);
            }
            s_u[q] = v;
        }
        __syncwarp();
        // LayerNorm statistics per agent row over its D float32 observation values (two-pass, like the materialised path)
        float csum = 0.f, cq = 0.f;       // critic: sum of the row sums, sum of the per-agent centred squares
        for (int i = 0; i < N; ++i) {
            const double pix = s_pv[i * 4], piy = s_pv[i * 4 + 1];
            float mean = 0.f, rstd = 1.f, rowsum = 0.f, qsum = 0.f;
            if (normalize) {
                float s = 0.f;
                for (int k = lane; k < OWN; k += 32) s += s_u[i * OWN + k];
                for (int j = lane; j < M; j += 32) {
                    const float dx = (float)dsub(poi[2 * j], pix), dy = (float)dsub(poi[2 * j + 1], piy);
                    s += (dx + dy) + (s_u[N * OWN + j] + cd.m_energy + s_u[N * OWN + M + j]);
                }
                rowsum = warp_sum_f(s);
                mean = rowsum / (float)D;
                float q = 0.f;
                for (int k = lane; k < OWN; k += 32) { const float d = s_u[i * OWN + k] - mean; q = fmaf(d, d, q); }
                for (int j = lane; j < M; j += 32) {
                    const float dx = (float)dsub(poi[2 * j], pix) - mean, dy = (float)dsub(poi[2 * j + 1], piy) - mean;
                    const float de = s_u[N * OWN + j] - mean, dm = cd.m_energy - mean, dd = s_u[N * OWN + M + j] - mean;
                    q = fmaf(dx, dx, q); q = fmaf(dy, dy, q); q = fmaf(de, de, q); q = fmaf(dm, dm, q); q = fmaf(dd, dd, q);
                }
                qsum = warp_sum_f(q);
                rstd = rsqrtf(qsum / (float)D + LN_EPS);
            }
            if (lane == 0) { s_stat[i] = mean; s_stat[N + i] = rstd; }
            csum += rowsum; cq += qsum;
        }
        __syncwarp();
        if (Fa) {
            for (int i = 0; i < N; ++i) {
                const float mean = s_stat[i], rstd = s_stat[N + i];
                float *out = Fa + ((size_t)r * N + i) * cd.lda;
                for (int c = lane; c < cd.lda; c += 32) {
                    float v;
                    if (c < OWN) v = s_u[i * OWN + c] * rstd;
                    else if (c < OWN + 2 * M) v = s_u[N * OWN + (c - OWN)] * rstd;
                    else if (c == OWN + 2 * M) v = rstd;
                    else if (c == OWN + 2 * M + 1) v = -mean * rstd;
                    else v = 0.f;
                    out[c] = v;
                }
            }
        }
        if (Fc) {
            float mean_c = 0.f, rstd_c = 1.f;
            if (normalize) {
                // all N rows have D elements: mean_c = sum of row sums / (N D);
                // sum (x - mean_c)^2 = sum_i [ q_i + D (mean_i - mean_c)^2 ]   (exact identity, no third pass)
                mean_c = csum / (float)(N * D);
                float extra = 0.f;
                for (int i = 0; i < N; ++i) { const float d = s_stat[i] - mean_c; extra = fmaf(d, d, extra); }
                rstd_c = rsqrtf((cq + (float)D * extra) / (float)(N * D) + LN_EPS);
            }
            float *out = Fc + (size_t)r * cd.ldc;
            const int nown = N * OWN;
            for (int c = lane; c < cd.ldc; c += 32) {
                float v;
                if (c < nown + 2 * M) v = s_u[c] * rstd_c;
                else if (c == nown + 2 * M) v = rstd_c;
                else if (c == nown + 2 * M + 1) v = -mean_c * rstd_c;
                else v = 0.f;
                out[c] = v;
            }
        }
    }
}

static inline size_t compact_features_smem(const CompactDims &cd, int warps) {
    return ((size_t)cd.N * 32 + align_up((size_t)(cd.N * cd.OWN + 2 * cd.M + 2 * cd.N) * 4, 16)) * warps;
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

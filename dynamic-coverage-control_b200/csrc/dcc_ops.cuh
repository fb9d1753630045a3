// dcc_ops.cuh — device kernels of the MAPPO learner path (actor/critic MLP forward+backward, Gaussian head,
// PPO loss, GAE, grad-norm clip + Adam).  Launch wrappers live in dcc_mappo.cu.
//
// Reference math (paths relative to /root/reference/uav_dcc_control/): algos/algo_utils/mlp.py:19-29,46-58
// (LayerNorm -> [Linear, ReLU, LayerNorm] x 2), algos/algo_utils/distributions.py:33-41,83-92 (diagonal Gaussian),
// algos/mappo.py:103-187 (PPO update), buffer/shared_buffer.py:199-208 (GAE), utils/valuenorm.py:32-79.
#pragma once
#include <cuda_fp16.h>
#include "dcc_common.cuh"

namespace dcc {

constexpr float LN_EPS = 1e-5f;
constexpr float LOG_2PI = 1.8378770664093453f;

// ---- SIMT fp32 GEMM ---------------------------------------------------------------------------------
// C[M,N] (+)= sum_k A(m,k) * B(k,n);  A(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  B(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n].
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, register-staged double buffering.  blockIdx.z splits K
// (results combined with float atomics when ATOMIC).  fp32 FFMA accumulate = the reference's own precision.
constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_PAD = 4;

template <bool TRANS>
__device__ __forceinline__ void gemm_fetch(const float *__restrict__ P, int ld, int r0, int k0, int R, int Kend, int t,
                                           float (&v)[8]) {
    // fetch this thread's 8 elements of a [128 (r) x 16 (k)] operand tile
    if (!TRANS) {  // element (r,k) at P[r*ld + k]: k contiguous.  thread -> row t/2, k segment (t%2)*8
        const int r = r0 + (t >> 1), kb = k0 + (t & 1) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (r < R && kb + i < Kend) ? P[(size_t)r * ld + kb + i] : 0.f;
    } else {       // element (r,k) at P[k*ld + r]: r contiguous.  thread -> k t/16, r segment (t%16)*8
        const int k = k0 + (t >> 4), rb = r0 + (t & 15) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (k < Kend && rb + i < R) ? P[(size_t)k * ld + rb + i] : 0.f;
    }
}
template <bool TRANS>
__device__ __forceinline__ void gemm_stash(float (*S)[GM_BM + GM_PAD], int t, const float (&v)[8]) {
    if (!TRANS) {
        const int r = t >> 1, kb = (t & 1) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) S[kb + i][r] = v[i];
    } else {
        const int k = t >> 4, rb = (t & 15) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) S[k][rb + i] = v[i];
    }
}

template <bool TA, bool TB, bool ATOMIC>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const float *__restrict__ A, int lda,
                                                   const float *__restrict__ B, int ldb, float *__restrict__ C, int ldc,
                                                   int k_per_split) {
    __shared__ __align__(16) float As[2][GM_BK][GM_BM + GM_PAD];
    __shared__ __align__(16) float Bs[2][GM_BK][GM_BN + GM_PAD];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    if (kbeg >= kend) return;
    const int ty = t >> 4, tx = t & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ra[8], rb[8];
    // A tile indexed (m,k): "row-major [m][k]" is the non-transposed fetch; TA means A is stored [k][m]
    gemm_fetch<TA>(A, lda, m0, kbeg, M, kend, t, ra);
    // B tile indexed (n,k): B(k,n)=B[k*ldb+n] is the [k][n] (transposed-fetch) storage; TB means stored [n][k]
    gemm_fetch<!TB>(B, ldb, n0, kbeg, N, kend, t, rb);
    gemm_stash<TA>(As[0], t, ra);
    gemm_stash<!TB>(Bs[0], t, rb);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
        const bool more = k0 + GM_BK < kend;
        if (more) {
            gemm_fetch<TA>(A, lda, m0, k0 + GM_BK, M, kend, t, ra);
            gemm_fetch<!TB>(B, ldb, n0, k0 + GM_BK, N, kend, t, rb);
        }
#pragma unroll
        for (int kk = 0; kk < GM_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            gemm_stash<TA>(As[buf ^ 1], t, ra);
            gemm_stash<!TB>(Bs[buf ^ 1], t, rb);
        }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n >= N) continue;
            if (ATOMIC) atomicAdd(&C[(size_t)m * ldc + n], acc[i][j]);
            else C[(size_t)m * ldc + n] = acc[i][j];
        }
    }
}

// ---- trunk activation (mlp.py:13: [nn.Tanh(), nn.ReLU()][use_ReLU]) -----------------------------------------------
// ACT_IDENT: no activation — the LayerNorm behind the GRU of a recurrent policy (rnn.py:22,79) reuses the block kernels
enum { ACT_RELU = 0, ACT_TANH = 1, ACT_IDENT = 2 };
__device__ __forceinline__ float act_fwd(float z, int act) {
    return act == ACT_RELU ? fmaxf(z, 0.f) : (act == ACT_TANH ? tanhf(z) : z);
}
// derivative expressed through the saved OUTPUT a = act(z): relu' = [a > 0], tanh' = 1 - a^2
__device__ __forceinline__ float act_bwd(float da, float a, int act) {
    return act == ACT_RELU ? (a > 0.f ? da : 0.f) : (act == ACT_TANH ? da * (1.f - a * a) : da);
}

// fp16 hi/lo split of two fp32 values, packed in memory order (same arithmetic as tc::split_f16_pair)
__device__ __forceinline__ void tc_split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Power of two that brings a tensor whose |values| are bounded by the float with these bits into [2^13, 2^14) — the scale a
// PRE-SPLIT gradient tensor is stored with.  Must stay identical to tc::f16_scale_from_absmax (the consumers undo it from the same bits).
__device__ __forceinline__ float split_scale_up(uint32_t bits) {
    const uint32_t e = (bits >> 23) & 0xffu;
    if (e == 0) return 1.f;
    uint32_t be = 267u - e;
    if (be > 254u) be = 254u;
    return __uint_as_float(be << 23);
}

// max |x| over a flat array as float bits (atomicMax on the bits of non-negative floats orders them); *out zeroed first
__global__ void absmax_flat_kernel(const float *__restrict__ x, size_t n, uint32_t *__restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// ---- row-wise kernels: one warp per row ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Input feature LayerNorm WITHOUT its affine part: xhat = (x - mean) * rstd over F features (two-pass mean /
// biased variance, eps 1e-5).  The affine (gamma0, beta0) is folded into the first Linear (fold_ln0_kernel):
//   LN(x) W1^T + b1 = xhat (W1 * gamma0)^T + (b1 + W1 beta0)
// which makes xhat the only [rows, F] operand of the layer (forward GEMM and weight-gradient GEMM) and removes the
// dX GEMM of layer 1 from the backward pass altogether (ln0_finalize_kernel).
// The output rows have leading dimension ldo >= F; the pad columns are zero-filled (the tensor-core GEMMs read K in
// multiples of 32).
// ridx (optional, minibatch path): output row r is computed from source row ridx[r] / rdiv (agent-row indices of a
// permutation; rdiv = N maps them to centralised rows for the critic).
// normalize == 0 (use_feature_normalization = false, mlp.py:52-53): plain copy into the padded layout.
__global__ void ln_noaffine_fwd_kernel(const float *__restrict__ x, float *__restrict__ xhat, int rows, int F, int ldo,
                                       const long long *__restrict__ ridx, int rdiv, int normalize) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const size_t sr = ridx ? (size_t)(ridx[r] / rdiv) : (size_t)r;
        const float *xr = x + sr * F;
        float mean = 0.f, rstd = 1.f;
        if (normalize) {
            float s = 0.f;
            for (int c = lane; c < F; c += 32) s += xr[c];
            mean = warp_sum_f(s) / (float)F;
            float q = 0.f;
            for (int c = lane; c < F; c += 32) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
            rstd = rsqrtf(warp_sum_f(q) / (float)F + LN_EPS);
        }
        float *yr = xhat + (size_t)r * ldo;
        for (int c = lane; c < ldo; c += 32) yr[c] = (c < F) ? (normalize ? (xr[c] - mean) * rstd : xr[c]) : 0.f;
    }
}

// Same, for rows that are 16-byte aligned and short enough to live in registers (F % 4 == 0, F <= 128 * NV): one
// global read pass with 16-byte loads, NV float4 per lane (the critic's centralised input: F = N*D = 2704 -> NV = 22).
template <int NV>
__global__ void ln_noaffine_fwd_vec_kernel(const float *__restrict__ x, float *__restrict__ xhat, int rows, int F, int ldo,
                                           const long long *__restrict__ ridx, int rdiv, int normalize) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int F4 = F >> 2, L4 = ldo >> 2;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const size_t sr = ridx ? (size_t)(ridx[r] / rdiv) : (size_t)r;
        const float4 *xr = reinterpret_cast<const float4 *>(x + sr * F);
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            v[i] = (c < F4) ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        float mean = 0.f, rstd = 1.f;
        if (normalize) {
            mean = warp_sum_f(s) / (float)F;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (lane + 32 * i < F4) {
                    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                    q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
                }
            }
            rstd = rsqrtf(warp_sum_f(q) / (float)F + LN_EPS);
        }
        float4 *yr = reinterpret_cast<float4 *>(xhat + (size_t)r * ldo);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < F4) yr[c] = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
            else if (c < L4) yr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int c = lane + 32 * NV; c < L4; c += 32) yr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Same, for rows that are only 8-byte aligned (F even but not a multiple of 4 — the actor's D = 338): one global read
// pass with 8-byte loads, NV float2 per lane (F <= 64 * NV); the padded output rows are 16-byte aligned.
template <int NV>
__global__ void ln_noaffine_fwd_vec2_kernel(const float *__restrict__ x, float *__restrict__ xhat, int rows, int F, int ldo,
                                            const long long *__restrict__ ridx, int rdiv, int normalize) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int F2 = F >> 1, L2 = ldo >> 1;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const size_t sr = ridx ? (size_t)(ridx[r] / rdiv) : (size_t)r;
        const float2 *xr = reinterpret_cast<const float2 *>(x + sr * F);
        float2 v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            v[i] = (c < F2) ? __ldg(xr + c) : make_float2(0.f, 0.f);
            s += v[i].x + v[i].y;
        }
        float mean = 0.f, rstd = 1.f;
        if (normalize) {
            mean = warp_sum_f(s) / (float)F;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (lane + 32 * i < F2) {
                    const float a = v[i].x - mean, b = v[i].y - mean;
                    q = fmaf(a, a, q); q = fmaf(b, b, q);
                }
            }
            rstd = rsqrtf(warp_sum_f(q) / (float)F + LN_EPS);
        }
        float2 *yr = reinterpret_cast<float2 *>(xhat + (size_t)r * ldo);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < F2) yr[c] = make_float2((v[i].x - mean) * rstd, (v[i].y - mean) * rstd);
            else if (c < L2) yr[c] = make_float2(0.f, 0.f);
        }
        for (int c = lane + 32 * NV; c < L2; c += 32) yr[c] = make_float2(0.f, 0.f);
    }
}

// W1g[h,c] = W1[h,c] * gamma0[c];  b1g[h] = b1[h] + sum_c W1[h,c] * beta0[c].  One warp per output unit h.
__global__ void fold_ln0_kernel(const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ g0,
                                const float *__restrict__ be0, float *__restrict__ W1g, float *__restrict__ b1g, int H,
                                int F) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= H) return;
    float s = 0.f;
    for (int c = lane; c < F; c += 32) {
        const float w = W1[(size_t)h * F + c];
        W1g[(size_t)h * F + c] = g0 ? w * g0[c] : w;      // g0 == nullptr: no feature_norm, fc1 is used as is
        if (g0) s = fmaf(w, be0[c], s);
    }
    s = warp_sum_f(s);
    if (lane == 0) b1g[h] = b1[h] + s;
}

// Backward of the folded input LayerNorm, once per epoch after all chunks: G = sum_r dz1^T xhat sits in the fc1
// weight-gradient slot and db1 in the bias slot.  Per input column c:
//   dgamma0[c] = sum_h W1[h,c] G[h,c];  dbeta0[c] = sum_h W1[h,c] db1[h];
//   dW1[h,c] = sum_r dz1[r,h] (xhat[r,c] gamma0[c] + beta0[c]) = G[h,c] gamma0[c] + db1[h] beta0[c]   (in place)
__global__ void ln0_finalize_kernel(const float *__restrict__ W1, const float *__restrict__ g0,
                                    const float *__restrict__ be0, float *__restrict__ G,
                                    const float *__restrict__ db1, float *__restrict__ dg0, float *__restrict__ dbe0,
                                    int H, int F) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    const float g = g0[c], b = be0[c];
    // float64 sums: the H terms of a column cancel by up to 5-6 orders of magnitude (the exact dgamma0 / dbeta0 of a column
    // can sit near Adam's eps while the terms are O(0.1)); a float32 chain leaves ~1e-6 absolute, which Adam turns into a
    // 1e-4 parameter difference for such a column.  Once per optimiser step, F threads x H terms: free.
    double ag = 0.0, ab = 0.0;
    for (int h = 0; h < H; ++h) {
        const float w = W1[(size_t)h * F + c];
        const float gv = G[(size_t)h * F + c];
        const float d1 = db1[h];
        ag = fma((double)w, (double)gv, ag);
        ab = fma((double)w, (double)d1, ab);
        G[(size_t)h * F + c] = fmaf(gv, g, d1 * b);
    }
    dg0[c] = (float)ag;
    dbe0[c] = (float)ab;
}

// a = act(z + bias); h = LayerNorm(a) * gamma + beta.  H <= 256 (8 columns per lane).  a_out optional.
__global__ void bias_relu_ln_fwd_kernel(const float *__restrict__ z, const float *__restrict__ bias,
                                        const float *__restrict__ gamma, const float *__restrict__ beta,
                                        float *__restrict__ a_out, float *__restrict__ h_out, float *__restrict__ mean_out,
                                        float *__restrict__ rstd_out, int rows, int H, int act) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        float a[8];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            a[j] = (c < H) ? act_fwd(z[(size_t)r * H + c] + bias[c], act) : 0.f;
            s += a[j];
        }
        const float mean = warp_sum_f(s) / (float)H;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = (lane + 32 * j < H) ? a[j] - mean : 0.f;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum_f(q) / (float)H + LN_EPS);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < H) {
                if (a_out) a_out[(size_t)r * H + c] = a[j];
                h_out[(size_t)r * H + c] = (a[j] - mean) * rstd * gamma[c] + beta[c];
            }
        }
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
    }
}

// Block-level combine of per-warp partial column sums before they go to global memory: every warp of a 256-thread
// block deposits its NV vectors of 8 values per lane (column = lane + 32 j) in shared memory, then thread t adds up
// column t over the warps and issues ONE atomic per vector — 8x fewer atomics than per-warp, which lets these
// streaming kernels run 4 CTAs per SM (enough loads in flight for HBM) without flooding the L2 atomic units.
// Column owned by element j of a lane's 8-vector.  Default map: lane + 32 j (any H <= 256, scalar accesses).  VEC map (H == 256
// only): two runs of 4 consecutive columns, 4 lane + {0..3} and 128 + 4 lane + {0..3}, so that a row moves with two 16-byte
// accesses per lane instead of eight 4-byte ones (the streaming kernels below are issue-bound, not bandwidth-bound).
template <bool VEC>
__device__ __forceinline__ int col_of(int lane, int j) {
    return VEC ? ((j < 4) ? 4 * lane + j : 128 + 4 * lane + (j - 4)) : lane + 32 * j;
}

template <int NV, bool VEC = false>
__device__ __forceinline__ void block_combine_atomic(float (&acc)[NV][8], float *const (&dst)[NV], int H, float *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[(warp * NV + v) * 256 + col_of<VEC>(lane, j)] = acc[v][j];
    __syncthreads();
    const int c = threadIdx.x;
    if (c < H) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float t = 0.f;
            for (int w = 0; w < nw; ++w) t += sm[(w * NV + v) * 256 + c];
            atomicAdd(&dst[v][c], t);
        }
    }
}

// Backward of h = LN(a)*gamma+beta, a = relu(z+bias):  given dh, a, mean, rstd -> dz (in place over dh allowed),
// and accumulates dgamma, dbeta, dbias (one set of float atomics per block).  H <= 256, blockDim = 256,
// dynamic shared memory = 8 * 3 * 256 floats.
// (fallback for H % 4 != 0; the pipelined variant relu_ln_bwd_pipe_kernel below is the one the learner uses)
// launch bounds: 4 CTAs per SM (64 registers, 32 bytes of spill) measured 1.8 % faster end to end than 3 (80 registers); the
// head-fused variant below loses with 3 CTAs (192 bytes of spill) and stays at 2
__global__ void __launch_bounds__(256, 4) relu_ln_bwd_kernel(const float *__restrict__ dh, const float *__restrict__ a, const float *__restrict__ mean,
                                   const float *__restrict__ rstd, const float *__restrict__ gamma, float *__restrict__ dz,
                                   float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dbias, int rows,
                                   int H, int act) {
    extern __shared__ float dyn_sm[];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float g[8], acc[3][8];   // acc: dgamma, dbeta, dbias
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        g[j] = (c < H) ? gamma[c] : 0.f;
        acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
    }
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const float m = mean[r], rs = rstd[r];
        float xh[8], dxh[8], av[8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            const bool ok = c < H;
            av[j] = ok ? a[(size_t)r * H + c] : 0.f;
            const float d = ok ? dh[(size_t)r * H + c] : 0.f;
            xh[j] = ok ? (av[j] - m) * rs : 0.f;
            acc[0][j] = fmaf(d, xh[j], acc[0][j]);
            acc[1][j] += d;
            dxh[j] = d * g[j];
            s1 += dxh[j];
            s2 = fmaf(dxh[j], xh[j], s2);
        }
        const float c1 = warp_sum_f(s1) / (float)H, c2 = warp_sum_f(s2) / (float)H;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < H) {
                const float da = rs * (dxh[j] - c1 - xh[j] * c2);
                const float v = act_bwd(da, av[j], act);
                dz[(size_t)r * H + c] = v;
                acc[2][j] += v;
            }
        }
    }
    float *const dst[3] = {dgamma, dbeta, dbias};
    block_combine_atomic<3>(acc, dst, H, dyn_sm);
}

// Head backward fused with the backward of the last trunk block.  The gradient w.r.t. the trunk output is rank-OUT,
//   dh2[r,c] = sum_o dout[r,o] * Wh[o,c],
// so it is formed on the fly instead of being written and read back; h2 (needed for dWh) is rebuilt from the saved
// post-ReLU activation: h2 = LN(a2) * gamma + beta.  Per row: reads a2 (and mean / rstd), writes dz2; accumulates
// dgamma, dbeta, dbias (of the Linear) and dWh, dbh with one set of float atomics per block.  H <= 256, blockDim = 256,
// dynamic shared memory = 8 * (3 + OUT) * 256 floats.
template <int OUT>
__global__ void __launch_bounds__(256, 2) head_relu_ln_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ Wh,
                                        const float *__restrict__ a, const float *__restrict__ mean,
                                        const float *__restrict__ rstd, const float *__restrict__ gamma,
                                        const float *__restrict__ beta, float *__restrict__ dz, float *__restrict__ dgamma,
                                        float *__restrict__ dbeta, float *__restrict__ dbias, float *__restrict__ dWh,
                                        float *__restrict__ dbh, int rows, int H, int act) {
    extern __shared__ float dyn_sm[];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float g[8], be[8], w[OUT][8], acc[3 + OUT][8], acc_bh[OUT];   // acc: dgamma, dbeta, dbias, dWh[0..OUT)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        g[j] = (c < H) ? gamma[c] : 0.f;
        be[j] = (c < H) ? beta[c] : 0.f;
        acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) { w[o][j] = (c < H) ? Wh[o * H + c] : 0.f; acc[3 + o][j] = 0.f; }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc_bh[o] = 0.f;
    // software prefetch of the next row (see relu_ln_bwd_kernel): this kernel runs 2 CTAs per SM, so a warp's single
    // 1 KB row in flight left HBM at a third of its bandwidth
    const int stride = gridDim.x * wpb;
    int r = blockIdx.x * wpb + (threadIdx.x >> 5);
    float an[8], dn[OUT], mn = 0.f, rn = 0.f;
    auto fetch = [&](int row) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            an[j] = (c < H) ? __ldg(a + (size_t)row * H + c) : 0.f;
        }
#pragma unroll
        for (int o = 0; o < OUT; ++o) dn[o] = __ldg(dout + (size_t)row * OUT + o);
        mn = __ldg(mean + row); rn = __ldg(rstd + row);
    };
    if (r < rows) fetch(r);
    for (; r < rows; r += stride) {
        const float m = mn, rs = rn;
        float d[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) { d[o] = dn[o]; acc_bh[o] += d[o]; }
        float xh[8], dxh[8], av[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = an[j];
        if (r + stride < rows) fetch(r + stride);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            const bool ok = c < H;
            xh[j] = ok ? (av[j] - m) * rs : 0.f;
            const float h2 = fmaf(xh[j], g[j], be[j]);
            float dh = 0.f;
#pragma unroll
            for (int o = 0; o < OUT; ++o) { dh = fmaf(d[o], w[o][j], dh); acc[3 + o][j] = fmaf(d[o], h2, acc[3 + o][j]); }
            acc[0][j] = fmaf(dh, xh[j], acc[0][j]);
            acc[1][j] += dh;
            dxh[j] = dh * g[j];
            s1 += dxh[j];
            s2 = fmaf(dxh[j], xh[j], s2);
        }
        const float c1 = warp_sum_f(s1) / (float)H, c2 = warp_sum_f(s2) / (float)H;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < H) {
                const float da = rs * (dxh[j] - c1 - xh[j] * c2);
                const float v = act_bwd(da, av[j], act);
                dz[(size_t)r * H + c] = v;
                acc[2][j] += v;
            }
        }
    }
    float *dst[3 + OUT];
    dst[0] = dgamma; dst[1] = dbeta; dst[2] = dbias;
#pragma unroll
    for (int o = 0; o < OUT; ++o) dst[3 + o] = dWh + o * H;
    block_combine_atomic<3 + OUT>(acc, dst, H, dyn_sm);
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < OUT; ++o) atomicAdd(&dbh[o], acc_bh[o]);
    }
}

// ---- bulk-async row pipeline for the one-warp-per-row streaming kernels ---------------------------------------------------
// These kernels are latency-bound, not bandwidth-bound: a warp has one 1-2 KB row in flight, and register prefetch of
// the next row costs occupancy (measured: relu_ln_bwd 220 -> 249 us with 80 registers / 3 CTAs per SM).  Here each warp
// owns a ring of RP_SLOTS row slots in shared memory that the TMA engine fills (cp.async.bulk global -> shared, completion
// counted on an mbarrier), RP_SLOTS rows ahead of the row being reduced: 32 warps x 3 slots x 1-2 KB = 96-192 KB of loads
// in flight per SM at no register cost.  Requires rows that are 16-byte multiples (H % 4 == 0) and 16-byte aligned.
constexpr int RP_SLOTS = 3;
__device__ __forceinline__ uint32_t rp_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rp_bar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void rp_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rp_bulk_g2s(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void rp_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

// relu_ln_bwd_kernel with the row pipeline.  Dynamic shared memory = max(8 * RP_SLOTS * 2 * H, 8 * 3 * 256) floats: the ring
// is dead when the block-level combine of the column sums starts and is reused for it.
// absmax_out (optional): atomicMax of the float bits of max |dz| (scale of the fp16-split dX GEMM that reads dz next).
__global__ void __launch_bounds__(256, 3) relu_ln_bwd_pipe_kernel(const float *dh, const float *__restrict__ a,
                                                                  const float *__restrict__ mean, const float *__restrict__ rstd,
                                                                  const float *__restrict__ gamma, float *dz,
                                                                  float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                                  float *__restrict__ dbias, int rows, int H, int act,
                                                                  uint32_t *__restrict__ absmax_out) {
    extern __shared__ __align__(128) float dyn_sm[];
    __shared__ __align__(8) uint64_t bars[8 * RP_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    const uint32_t row_bytes = (uint32_t)H * 4;
    float *ring = dyn_sm + (size_t)warp * RP_SLOTS * 2 * H;
    const uint32_t ring_u32 = rp_smem_u32(ring), bar_u32 = rp_smem_u32(bars + warp * RP_SLOTS);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RP_SLOTS; ++s) rp_bar_init(bar_u32 + 8 * s);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float g[8], acc[3][8];   // acc: dgamma, dbeta, dbias
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        g[j] = (c < H) ? gamma[c] : 0.f;
        acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
    }
    const int r0 = blockIdx.x * wpb + warp;
    const int n_my = r0 < rows ? (rows - r0 + stride - 1) / stride : 0;
    auto issue = [&](int slot, int row) {      // lane 0
        const uint32_t bar = bar_u32 + 8 * slot, dst = ring_u32 + (uint32_t)slot * 2 * row_bytes;
        rp_expect_tx(bar, 2 * row_bytes);
        rp_bulk_g2s(dst, a + (size_t)row * H, row_bytes, bar);
        rp_bulk_g2s(dst + row_bytes, dh + (size_t)row * H, row_bytes, bar);
    };
    if (lane == 0)
        for (int s = 0; s < RP_SLOTS && s < n_my; ++s) issue(s, r0 + s * stride);
    float mn = 0.f, rn = 0.f, amax = 0.f;
    if (n_my > 0) { mn = __ldg(mean + r0); rn = __ldg(rstd + r0); }
    int slot = 0;
    uint32_t parity = 0;
    for (int k = 0; k < n_my; ++k) {
        const int r = r0 + k * stride;
        const float m = mn, rs = rn;
        if (k + 1 < n_my) { mn = __ldg(mean + r + stride); rn = __ldg(rstd + r + stride); }
        rp_wait(bar_u32 + 8 * slot, parity);
        const float *sa = ring + (size_t)slot * 2 * H, *sd = sa + H;
        float xh[8], dxh[8], av[8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            const bool ok = c < H;
            av[j] = ok ? sa[c] : 0.f;
            const float d = ok ? sd[c] : 0.f;
            xh[j] = ok ? (av[j] - m) * rs : 0.f;
            acc[0][j] = fmaf(d, xh[j], acc[0][j]);
            acc[1][j] += d;
            dxh[j] = d * g[j];
            s1 += dxh[j];
            s2 = fmaf(dxh[j], xh[j], s2);
        }
        __syncwarp();                               // every lane has read the slot
        if (lane == 0 && k + RP_SLOTS < n_my) {
            fence_proxy_async_smem();               // generic-proxy reads before the async-proxy refill of the slot
            issue(slot, r + RP_SLOTS * stride);
        }
        const float c1 = warp_sum_f(s1) / (float)H, c2 = warp_sum_f(s2) / (float)H;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < H) {
                const float da = rs * (dxh[j] - c1 - xh[j] * c2);
                const float v = act_bwd(da, av[j], act);
                dz[(size_t)r * H + c] = v;
                acc[2][j] += v;
                amax = fmaxf(amax, fabsf(v));
            }
        }
        if (++slot == RP_SLOTS) { slot = 0; parity ^= 1; }
    }
    if (absmax_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL_MASK, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(absmax_out, __float_as_uint(amax));
    }
    __syncthreads();                                // all rings idle: the combine scratch aliases them
    float *const dst[3] = {dgamma, dbeta, dbias};
    block_combine_atomic<3>(acc, dst, H, dyn_sm);
}

// "xhat mode" backward of an inner block (dcc_mappo.cu, MappoHandle::xhat): the block's LayerNorm output was stored WITHOUT affine
// and pre-split, xhat = hi + lo (two fp16 rows), and the gradient arriving here is already the one w.r.t. xhat (the next block ran
// on gamma-folded weights, so dX = dz W' = (dz W) * gamma).  Per row: LayerNorm backward through the normalisation only, then the
// ReLU mask rebuilt from xhat: a > 0  <=>  xhat > xhat(a = 0) = (0 - mean) rstd, compared after the same fp16 hi/lo rounding the
// stored values went through (monotone, so the only rows that can differ from the exact mask are activations within 2^-22 of the
// row's |xhat(0)| above zero).  Accumulates only the Linear's bias gradient: dgamma / dbeta of this block's LayerNorm follow once
// per optimiser step from the next block's G = dz^T xhat (ln0_finalize_kernel).  Reads 2 KB per row (dxh fp32 + hi + lo), the same
// bytes as relu_ln_bwd_pipe_kernel; one column accumulator instead of three.
// Dynamic shared memory = max(8 * RP_SLOTS * 2 * H, 8 * 1 * 256) floats.  H % 8 == 0.
// SPLIT (VEC only): dz leaves PRE-SPLIT — fp16 hi at dz16, lo at dz16 + lo_off, row pitch H halves — multiplied by the power of two
// that follows from an upper BOUND of |dz| known before a single row is processed: da / rstd = (I - 11^T/H - xhat xhat^T/H) dxh is the
// image of dxh under a matrix with eigenvalues in [0, 1] (|xhat|^2 / H = var / (var + eps) <= 1), so every |da_c| <= rstd |dxh|_2 <=
// max rstd * sqrt(H) * max |dxh| (x 1.25 for rounding), and |dz| <= |da|; both maxima were
// left behind by the kernels that produced rstd / dxh (sc_rstd, sc_dxh: float bits).  Block 0 publishes the bound's bits in
// *sc_bnd for the GEMMs that read dz (they undo the same power of two).  A loose bound costs nothing: see DESIGN.md §5.7.
template <bool VEC, bool SPLIT = false>
__global__ void __launch_bounds__(256, 4) relu_lnx_bwd_pipe_kernel(const float *dxh_in, const __half *__restrict__ xh_hi,
                                                                   const __half *__restrict__ xh_lo, const float *__restrict__ mean,
                                                                   const float *__restrict__ rstd, float *dz,
                                                                   float *__restrict__ dbias, int rows, int H,
                                                                   uint32_t *__restrict__ absmax_out,
                                                                   const uint32_t *__restrict__ sc_dxh = nullptr,
                                                                   const uint32_t *__restrict__ sc_rstd = nullptr,
                                                                   uint32_t *__restrict__ sc_bnd = nullptr, size_t lo_off = 0) {
    static_assert(!SPLIT || VEC, "the pre-split output uses the 16-byte column map");
    float up = 1.f;
    if constexpr (SPLIT) {
        const float bound = __uint_as_float(__ldg(sc_dxh)) * __uint_as_float(__ldg(sc_rstd)) * sqrtf((float)H) * 1.25f;
        const uint32_t bits = __float_as_uint(bound);
        up = split_scale_up(bits);
        if (blockIdx.x == 0 && threadIdx.x == 0) *sc_bnd = bits;
    }
    extern __shared__ __align__(128) float dyn_sm[];
    __shared__ __align__(8) uint64_t bars[8 * RP_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    const uint32_t row_bytes = (uint32_t)H * 4, half_bytes = (uint32_t)H * 2;
    float *ring = dyn_sm + (size_t)warp * RP_SLOTS * 2 * H;
    const uint32_t ring_u32 = rp_smem_u32(ring), bar_u32 = rp_smem_u32(bars + warp * RP_SLOTS);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RP_SLOTS; ++s) rp_bar_init(bar_u32 + 8 * s);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float accb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) accb[j] = 0.f;
    const int r0 = blockIdx.x * wpb + warp;
    const int n_my = r0 < rows ? (rows - r0 + stride - 1) / stride : 0;
    auto issue = [&](int slot, int row) {      // lane 0: slot = [dxh row | hi row | lo row]
        const uint32_t bar = bar_u32 + 8 * slot, dst = ring_u32 + (uint32_t)slot * 2 * row_bytes;
        rp_expect_tx(bar, 2 * row_bytes);
        rp_bulk_g2s(dst, dxh_in + (size_t)row * H, row_bytes, bar);
        rp_bulk_g2s(dst + row_bytes, xh_hi + (size_t)row * H, half_bytes, bar);
        rp_bulk_g2s(dst + row_bytes + half_bytes, xh_lo + (size_t)row * H, half_bytes, bar);
    };
    if (lane == 0)
        for (int s = 0; s < RP_SLOTS && s < n_my; ++s) issue(s, r0 + s * stride);
    float mn = 0.f, rn = 0.f, amax = 0.f;
    if (n_my > 0) { mn = __ldg(mean + r0); rn = __ldg(rstd + r0); }
    int slot = 0;
    uint32_t parity = 0;
    for (int k = 0; k < n_my; ++k) {
        const int r = r0 + k * stride;
        const float m = mn, rs = rn;
        if (k + 1 < n_my) { mn = __ldg(mean + r + stride); rn = __ldg(rstd + r + stride); }
        // xhat of a zero activation, through the same hi / lo rounding as the stored values
        const float t0 = (0.f - m) * rs;
        const float th = __half2float(__float2half_rn(t0));
        const float thr = th + __half2float(__float2half_rn(t0 - th));
        rp_wait(bar_u32 + 8 * slot, parity);
        const float *sd = ring + (size_t)slot * 2 * H;
        const __half *shi = reinterpret_cast<const __half *>(sd + H), *slo = shi + H;
        float xh[8], dxh[8];
        float s1 = 0.f, s2 = 0.f;
        if constexpr (VEC) {      // H == 256: 16-byte gradient loads, 8-byte loads of 4 halves
            const float4 d0 = *reinterpret_cast<const float4 *>(sd + 4 * lane), d1 = *reinterpret_cast<const float4 *>(sd + 128 + 4 * lane);
            dxh[0] = d0.x; dxh[1] = d0.y; dxh[2] = d0.z; dxh[3] = d0.w; dxh[4] = d1.x; dxh[5] = d1.y; dxh[6] = d1.z; dxh[7] = d1.w;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint2 hv = *reinterpret_cast<const uint2 *>(shi + 128 * q + 4 * lane), lv = *reinterpret_cast<const uint2 *>(slo + 128 * q + 4 * lane);
                const float2 h01 = __half22float2(*reinterpret_cast<const __half2 *>(&hv.x)), h23 = __half22float2(*reinterpret_cast<const __half2 *>(&hv.y));
                const float2 l01 = __half22float2(*reinterpret_cast<const __half2 *>(&lv.x)), l23 = __half22float2(*reinterpret_cast<const __half2 *>(&lv.y));
                xh[4 * q + 0] = h01.x + l01.x; xh[4 * q + 1] = h01.y + l01.y; xh[4 * q + 2] = h23.x + l23.x; xh[4 * q + 3] = h23.y + l23.y;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1 += dxh[j]; s2 = fmaf(dxh[j], xh[j], s2); }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                const bool ok = c < H;
                dxh[j] = ok ? sd[c] : 0.f;
                xh[j] = ok ? __half2float(shi[c]) + __half2float(slo[c]) : 0.f;
                s1 += dxh[j];
                s2 = fmaf(dxh[j], xh[j], s2);
            }
        }
        __syncwarp();                               // every lane has read the slot
        if (lane == 0 && k + RP_SLOTS < n_my) {
            fence_proxy_async_smem();
            issue(slot, r + RP_SLOTS * stride);
        }
        const float c1 = warp_sum_f(s1) / (float)H, c2 = warp_sum_f(s2) / (float)H;
        float vout[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col_of<VEC>(lane, j);
            vout[j] = 0.f;
            if (VEC || c < H) {
                const float da = rs * (dxh[j] - c1 - xh[j] * c2);
                const float v = xh[j] > thr ? da : 0.f;
                vout[j] = v;
                if constexpr (!VEC) dz[(size_t)r * H + c] = v;
                accb[j] += v;
                amax = fmaxf(amax, fabsf(v));
            }
        }
        if constexpr (SPLIT) {
            __half *hrow = reinterpret_cast<__half *>(dz) + (size_t)r * H;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint2 hi, lo;
                tc_split_pair(vout[4 * q] * up, vout[4 * q + 1] * up, hi.x, lo.x);
                tc_split_pair(vout[4 * q + 2] * up, vout[4 * q + 3] * up, hi.y, lo.y);
                *reinterpret_cast<uint2 *>(hrow + 128 * q + 4 * lane) = hi;
                *reinterpret_cast<uint2 *>(hrow + lo_off + 128 * q + 4 * lane) = lo;
            }
        } else if constexpr (VEC) {
            float *drow = dz + (size_t)r * H;
            *reinterpret_cast<float4 *>(drow + 4 * lane) = make_float4(vout[0], vout[1], vout[2], vout[3]);
            *reinterpret_cast<float4 *>(drow + 128 + 4 * lane) = make_float4(vout[4], vout[5], vout[6], vout[7]);
        }
        if (++slot == RP_SLOTS) { slot = 0; parity ^= 1; }
    }
    if (absmax_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL_MASK, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(absmax_out, __float_as_uint(amax));
    }
    __syncthreads();                                // all rings idle: the combine scratch aliases them
    float acc1[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc1[0][j] = accb[j];
    float *const dst[1] = {dbias};
    block_combine_atomic<1, VEC>(acc1, dst, H, dyn_sm);
}

// head_relu_ln_bwd_kernel with the row pipeline (one array: the saved activation a).  Dynamic shared memory =
// max(8 * RP_SLOTS * H, 8 * (3 + OUT) * 256) floats.
// Because dh2 = dout Wh has rank OUT, three of the per-row column accumulators are not needed: with
//   P[o][c] = sum_r dout[r,o] xhat[r,c]   and   D[o] = sum_r dout[r,o]
// the gradients are  dWh[o,c] = gamma[c] P[o][c] + beta[c] D[o],  dgamma[c] = sum_o Wh[o,c] P[o][c],  dbeta[c] = sum_o Wh[o,c] D[o]
// (formed once per warp after the row loop), so the loop carries OUT + 1 column accumulators (P, dbias) instead of OUT + 3
// and fewer instructions per row.  (Squeezed into 80 registers for 3 CTAs per SM it ran SLOWER, 158 -> 188 us: kept at 2.)
// SLOTS = rows in flight per warp.  6 instead of 3 was measured (96 KB instead of 48 KB of loads in flight per SM): no change,
// 179.0 vs 178.4 ms per update at 8192 envs — the kernel is issue-bound (ncu: 72 % issue utilisation with 3.9 warps per scheduler,
// 368 warp instructions per row), not latency-bound.
// SPLIT (VEC only): dz leaves pre-split and pre-scaled like relu_lnx_bwd_pipe_kernel<true, true>; here the gradient w.r.t. xhat is
// dxh[c] = sum_o dout[o] Wh[o,c] gamma[c], so |dxh| <= max |dout| * S with S = sum_o max_c |Wh[o,c] gamma[c]| (formed from the
// registers every warp holds anyway), and |dz| <= max rstd * sqrt(H) * max |dout| * S (x 1.25).
template <int OUT, int SLOTS = RP_SLOTS, bool VEC = false, bool SPLIT = false>
__global__ void __launch_bounds__(256, 2) head_relu_ln_bwd_pipe_kernel(const float *__restrict__ dout, const float *__restrict__ Wh,
                                        const float *__restrict__ a, const float *__restrict__ mean,
                                        const float *__restrict__ rstd, const float *__restrict__ gamma,
                                        const float *__restrict__ beta, float *__restrict__ dz, float *__restrict__ dgamma,
                                        float *__restrict__ dbeta, float *__restrict__ dbias, float *__restrict__ dWh,
                                        float *__restrict__ dbh, int rows, int H, int act, uint32_t *__restrict__ absmax_out,
                                        const uint32_t *__restrict__ sc_dout = nullptr, const uint32_t *__restrict__ sc_rstd = nullptr,
                                        uint32_t *__restrict__ sc_bnd = nullptr, size_t lo_off = 0) {
    static_assert(!SPLIT || VEC, "the pre-split output uses the 16-byte column map");
    extern __shared__ __align__(128) float dyn_sm[];
    __shared__ __align__(8) uint64_t bars[8 * SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    const uint32_t row_bytes = (uint32_t)H * 4;
    float *ring = dyn_sm + (size_t)warp * SLOTS * H;
    const uint32_t ring_u32 = rp_smem_u32(ring), bar_u32 = rp_smem_u32(bars + warp * SLOTS);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) rp_bar_init(bar_u32 + 8 * s);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float g[8], wg[OUT][8], P[OUT][8], accb[8], D[OUT];   // wg = Wh * gamma: dxhat = sum_o dout[o] wg[o]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = col_of<VEC>(lane, j);
        g[j] = (c < H) ? gamma[c] : 0.f;
        accb[j] = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) { wg[o][j] = (c < H) ? Wh[o * H + c] * g[j] : 0.f; P[o][j] = 0.f; }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) D[o] = 0.f;
    float up = 1.f;
    if constexpr (SPLIT) {
        float S = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
            float m = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(wg[o][j]));
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, sh));
            S += m;
        }
        const float bound = __uint_as_float(__ldg(sc_dout)) * S * __uint_as_float(__ldg(sc_rstd)) * sqrtf((float)H) * 1.25f;
        const uint32_t bits = __float_as_uint(bound);
        up = split_scale_up(bits);
        if (blockIdx.x == 0 && threadIdx.x == 0) *sc_bnd = bits;
    }
    const int r0 = blockIdx.x * wpb + warp;
    const int n_my = r0 < rows ? (rows - r0 + stride - 1) / stride : 0;
    auto issue = [&](int slot, int row) {      // lane 0
        const uint32_t bar = bar_u32 + 8 * slot;
        rp_expect_tx(bar, row_bytes);
        rp_bulk_g2s(ring_u32 + (uint32_t)slot * row_bytes, a + (size_t)row * H, row_bytes, bar);
    };
    if (lane == 0)
        for (int s = 0; s < SLOTS && s < n_my; ++s) issue(s, r0 + s * stride);
    float mn = 0.f, rn = 0.f, dn[OUT], amax = 0.f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) dn[o] = 0.f;
    if (n_my > 0) {
        mn = __ldg(mean + r0); rn = __ldg(rstd + r0);
#pragma unroll
        for (int o = 0; o < OUT; ++o) dn[o] = __ldg(dout + (size_t)r0 * OUT + o);
    }
    int slot = 0;
    uint32_t parity = 0;
    for (int k = 0; k < n_my; ++k) {
        const int r = r0 + k * stride;
        const float m = mn, rs = rn;
        float d[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) { d[o] = dn[o]; D[o] += d[o]; }
        if (k + 1 < n_my) {
            mn = __ldg(mean + r + stride); rn = __ldg(rstd + r + stride);
#pragma unroll
            for (int o = 0; o < OUT; ++o) dn[o] = __ldg(dout + (size_t)(r + stride) * OUT + o);
        }
        rp_wait(bar_u32 + 8 * slot, parity);
        const float *sa = ring + (size_t)slot * H;
        float xh[8], dxh[8], av[8];
        float s1 = 0.f, s2 = 0.f;
        if constexpr (VEC) {
            const float4 v0 = *reinterpret_cast<const float4 *>(sa + 4 * lane), v1 = *reinterpret_cast<const float4 *>(sa + 128 + 4 * lane);
            av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w; av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col_of<VEC>(lane, j);
            const bool ok = VEC || c < H;
            if constexpr (!VEC) av[j] = ok ? sa[c] : 0.f;
            xh[j] = ok ? (av[j] - m) * rs : 0.f;
            float dx = 0.f;
#pragma unroll
            for (int o = 0; o < OUT; ++o) { dx = fmaf(d[o], wg[o][j], dx); P[o][j] = fmaf(d[o], xh[j], P[o][j]); }
            dxh[j] = dx;
            s1 += dx;
            s2 = fmaf(dx, xh[j], s2);
        }
        __syncwarp();
        if (lane == 0 && k + SLOTS < n_my) {
            fence_proxy_async_smem();
            issue(slot, r + SLOTS * stride);
        }
        const float c1 = warp_sum_f(s1) / (float)H, c2 = warp_sum_f(s2) / (float)H;
        float vout[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col_of<VEC>(lane, j);
            vout[j] = 0.f;
            if (VEC || c < H) {
                const float da = rs * (dxh[j] - c1 - xh[j] * c2);
                const float v = act_bwd(da, av[j], act);
                vout[j] = v;
                if constexpr (!VEC) dz[(size_t)r * H + c] = v;
                accb[j] += v;
                amax = fmaxf(amax, fabsf(v));
            }
        }
        if constexpr (SPLIT) {
            __half *hrow = reinterpret_cast<__half *>(dz) + (size_t)r * H;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint2 hi, lo;
                tc_split_pair(vout[4 * q] * up, vout[4 * q + 1] * up, hi.x, lo.x);
                tc_split_pair(vout[4 * q + 2] * up, vout[4 * q + 3] * up, hi.y, lo.y);
                *reinterpret_cast<uint2 *>(hrow + 128 * q + 4 * lane) = hi;
                *reinterpret_cast<uint2 *>(hrow + lo_off + 128 * q + 4 * lane) = lo;
            }
        } else if constexpr (VEC) {
            float *drow = dz + (size_t)r * H;
            *reinterpret_cast<float4 *>(drow + 4 * lane) = make_float4(vout[0], vout[1], vout[2], vout[3]);
            *reinterpret_cast<float4 *>(drow + 128 + 4 * lane) = make_float4(vout[4], vout[5], vout[6], vout[7]);
        }
        if (++slot == SLOTS) { slot = 0; parity ^= 1; }
    }
    if (absmax_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL_MASK, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(absmax_out, __float_as_uint(amax));
    }
    // this warp's contributions to dgamma, dbeta, dbias, dWh[o] from P, D (see the header)
    float acc[3 + OUT][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = col_of<VEC>(lane, j);
        const float bj = (c < H) ? beta[c] : 0.f;
        float dg = 0.f, db = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
            const float w = (c < H) ? Wh[o * H + c] : 0.f;
            dg = fmaf(w, P[o][j], dg);
            db = fmaf(w, D[o], db);
            acc[3 + o][j] = fmaf(g[j], P[o][j], bj * D[o]);
        }
        acc[0][j] = dg; acc[1][j] = db; acc[2][j] = accb[j];
    }
    __syncthreads();
    float *dst[3 + OUT];
    dst[0] = dgamma; dst[1] = dbeta; dst[2] = dbias;
#pragma unroll
    for (int o = 0; o < OUT; ++o) dst[3 + o] = dWh + o * H;
    block_combine_atomic<3 + OUT, VEC>(acc, dst, H, dyn_sm);
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < OUT; ++o) atomicAdd(&dbh[o], D[o]);
    }
}

// ---- counter-based RNG (Philox4x32-10) + Box-Muller: action sampling a = mu + sigma * eps ---------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float2 normal_pair(uint64_t seed, uint64_t offset, uint64_t row) {
    // counter = (global agent row, call offset): the sample stream does not depend on chunking or launch geometry
    uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u1 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
    const float u2 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincospif(2.f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// Actor head: mu = h . Wm^T + bm (act_dim = 2), sigma = exp(logstd).
//   mode 0 (rollout, R_Actor.forward): sample (or mode when deterministic) the action, write action + logp
//   mode 1 (update, evaluate_actions): logp of the given action, also writes mu
__global__ void actor_head_kernel(const float *__restrict__ h, const float *__restrict__ Wm, const float *__restrict__ bm,
                                  const float *__restrict__ logstd, float *__restrict__ actions, float *__restrict__ mu_out,
                                  float *__restrict__ logp_out, int rows, int H, int mode, int deterministic, uint64_t seed,
                                  uint64_t offset, uint64_t row_base, const long long *__restrict__ ridx = nullptr) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float ls0 = logstd[0], ls1 = logstd[1];
    // two rows per warp iteration: twice the loads in flight per warp (the kernel is latency-, not bandwidth-bound)
    for (int r0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 2; r0 < rows; r0 += gridDim.x * wpb * 2) {
        const bool two = r0 + 1 < rows;
        float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        for (int c = lane; c < H; c += 32) {
            const float w0 = Wm[c], w1 = Wm[H + c];
            const float v0 = h[(size_t)r0 * H + c];
            const float v1 = two ? h[(size_t)(r0 + 1) * H + c] : 0.f;
            s[0][0] = fmaf(v0, w0, s[0][0]); s[0][1] = fmaf(v0, w1, s[0][1]);
            s[1][0] = fmaf(v1, w0, s[1][0]); s[1][1] = fmaf(v1, w1, s[1][1]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            s[k][0] = warp_sum_f(s[k][0]) + bm[0];
            s[k][1] = warp_sum_f(s[k][1]) + bm[1];
        }
        if (lane < 2 && (lane == 0 || two)) {      // lane k finishes row r0 + k
            const int r = r0 + lane;
            const float s0 = lane ? s[1][0] : s[0][0], s1 = lane ? s[1][1] : s[0][1];
            const float sd0 = expf(ls0), sd1 = expf(ls1);
            float a0, a1;
            if (mode == 0) {
                if (deterministic) { a0 = s0; a1 = s1; }
                else {
                    const float2 e = normal_pair(seed, offset, row_base + (uint64_t)r);
                    a0 = fmaf(sd0, e.x, s0); a1 = fmaf(sd1, e.y, s1);
                }
                actions[(size_t)r * 2 + 0] = a0; actions[(size_t)r * 2 + 1] = a1;
            } else {
                const size_t ar = ridx ? (size_t)ridx[r] : (size_t)r;   // minibatch path: the given actions are gathered
                a0 = actions[ar * 2 + 0]; a1 = actions[ar * 2 + 1];
            }
            if (mu_out) { mu_out[(size_t)r * 2 + 0] = s0; mu_out[(size_t)r * 2 + 1] = s1; }
            const float d0 = a0 - s0, d1 = a1 - s1;
            // Normal.log_prob summed over the 2 action dims (distributions.py:33-35)
            const float lp = -(d0 * d0) / (2.f * sd0 * sd0) - ls0 - 0.5f * LOG_2PI - (d1 * d1) / (2.f * sd1 * sd1) - ls1 -
                             0.5f * LOG_2PI;
            if (logp_out) logp_out[r] = lp;
        }
    }
}

// Output head folded through the last block's LayerNorm affine, for the fused GEMM epilogue (TcfParams::head_fold):
//   fold[o * 256 + c] = gamma[c] * Wh[o, c];  fold[out * 256 + o] = sum_c gamma[c] Wh[o, c];
//   fold[out * 256 + out + o] = sum_c beta[c] Wh[o, c] + bh[o]            (one block of 256 threads, H <= 256, sums in float64)
__global__ void head_fold_kernel(const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ Wh,
                                 const float *__restrict__ bh, int H, int out, float *__restrict__ fold) {
    __shared__ double red[2][8];
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
    for (int o = 0; o < out; ++o) {
        double gw = 0.0, bw = 0.0;
        if (c < H) {
            const float w = Wh[o * H + c];
            const float g = gamma[c] * w;
            fold[o * 256 + c] = g;
            gw = (double)g;
            bw = (double)beta[c] * (double)w;
        } else if (c < 256) fold[o * 256 + c] = 0.f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) { gw += __shfl_xor_sync(FULL_MASK, gw, s); bw += __shfl_xor_sync(FULL_MASK, bw, s); }
        if (lane == 0) { red[0][warp] = gw; red[1][warp] = bw; }
        __syncthreads();
        if (c == 0) {
            double a = 0.0, b = 0.0;
            for (int w8 = 0; w8 < 8; ++w8) { a += red[0][w8]; b += red[1][w8]; }
            fold[out * 256 + o] = (float)a;
            fold[out * 256 + out + o] = (float)(b + (double)bh[o]);
        }
        __syncthreads();
    }
}

// Second half of the actor head when mu was produced by the fused GEMM epilogue (tcgen05 backend): sampling / log-prob
// from mu [rows, 2].  One thread per row; same modes, Philox keying and log-prob formula as actor_head_kernel.
// Normal.log_prob summed over the 2 action dims (distributions.py:33-35); ONE expression for every kernel that forms it
__device__ __forceinline__ float gauss_logp2(float a0, float a1, float s0, float s1, float ls0, float ls1, float sd0, float sd1) {
    const float d0 = a0 - s0, d1 = a1 - s1;
    return -(d0 * d0) / (2.f * sd0 * sd0) - ls0 - 0.5f * LOG_2PI - (d1 * d1) / (2.f * sd1 * sd1) - ls1 - 0.5f * LOG_2PI;
}

__global__ void gauss_finish_kernel(const float *__restrict__ mu, const float *__restrict__ logstd, float *__restrict__ actions,
                                    float *__restrict__ mu_out, float *__restrict__ logp_out, int rows, int mode, int deterministic,
                                    uint64_t seed, uint64_t offset, uint64_t row_base, const long long *__restrict__ ridx = nullptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float ls0 = logstd[0], ls1 = logstd[1];
    const float2 m = *reinterpret_cast<const float2 *>(mu + (size_t)r * 2);
    const float s0 = m.x, s1 = m.y;
    const float sd0 = expf(ls0), sd1 = expf(ls1);
    float a0, a1;
    if (mode == 0) {
        if (deterministic) { a0 = s0; a1 = s1; }
        else {
            const float2 e = normal_pair(seed, offset, row_base + (uint64_t)r);
            a0 = fmaf(sd0, e.x, s0); a1 = fmaf(sd1, e.y, s1);
        }
        actions[(size_t)r * 2 + 0] = a0; actions[(size_t)r * 2 + 1] = a1;
    } else {
        const size_t ar = ridx ? (size_t)ridx[r] : (size_t)r;
        a0 = actions[ar * 2 + 0]; a1 = actions[ar * 2 + 1];
    }
    if (mu_out && mu_out != mu) { mu_out[(size_t)r * 2 + 0] = s0; mu_out[(size_t)r * 2 + 1] = s1; }
    const float lp = gauss_logp2(a0, a1, s0, s1, ls0, ls1, sd0, sd1);
    if (logp_out) logp_out[r] = lp;
}

// Critic head: v = h . wv + bv
__global__ void critic_head_kernel(const float *__restrict__ h, const float *__restrict__ wv, const float *__restrict__ bv,
                                   float *__restrict__ v_out, int rows, int H) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s = fmaf(h[(size_t)r * H + c], wv[c], s);
        s = warp_sum_f(s);
        if (lane == 0) v_out[r] = s + bv[0];
    }
}

// ---- ValueNorm helpers (utils/valuenorm.py:32-36): state = {running_mean, running_mean_sq, debiasing_term} ----
// vn == nullptr: no value normaliser (use_valuenorm = false, mappo.py:100-101) -> identity.
__device__ __forceinline__ void vn_mean_std(const float *vn, float &mean, float &stdv) {
    if (!vn) { mean = 0.f; stdv = 1.f; return; }
    const float c = fmaxf(vn[2], 1e-5f);
    mean = vn[0] / c;
    const float var = fmaxf(vn[1] / c - mean * mean, 1e-2f);
    stdv = sqrtf(var);
}

// Learner.insert bookkeeping (learner.py:254-276): per-env reward and mask = 1 - done
__global__ void rollout_insert_kernel(const float *__restrict__ rew_in, const uint8_t *__restrict__ done_in, int E, int N,
                                      float *__restrict__ rew_out, float *__restrict__ mask_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    rew_out[e] = rew_in[(size_t)e * N];
    mask_out[e] = done_in[(size_t)e * N] ? 0.f : 1.f;
}

// GAE over the rollout (shared_buffer.py:199-208): one thread per env, backwards in time, arrays [T(+1), E]
// (coalesced over envs); masks cut the recurrence at episode ends (the "segments").
// use_gae == 0: plain discounted returns, bootstrapped from the RAW (not denormalised) next value, exactly as
// shared_buffer.py:209-212 does: returns[T] = next_value; returns[t] = returns[t+1] * gamma * masks[t+1] + rewards[t].
__global__ void gae_kernel(const float *__restrict__ rew, const float *__restrict__ val, const float *__restrict__ masks,
                           const float *__restrict__ vn, float *__restrict__ ret, int T, int E, float gamma, float lam,
                           int use_gae) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (!use_gae) {
        float r = val[(size_t)T * E + e];
        ret[(size_t)T * E + e] = r;
        for (int t = T - 1; t >= 0; --t) {
            r = __fadd_rn(__fmul_rn(__fmul_rn(r, gamma), masks[(size_t)(t + 1) * E + e]), rew[(size_t)t * E + e]);
            ret[(size_t)t * E + e] = r;
        }
        return;
    }
    float mean, sd;
    vn_mean_std(vn, mean, sd);
    float gae = 0.f;
    float vnext = val[(size_t)T * E + e] * sd + mean;
    for (int t = T - 1; t >= 0; --t) {
        const float m = masks[(size_t)(t + 1) * E + e];
        const float v = val[(size_t)t * E + e] * sd + mean;
        const float delta = rew[(size_t)t * E + e] + gamma * vnext * m - v;
        gae = delta + gamma * lam * m * gae;
        ret[(size_t)t * E + e] = gae + v;
        vnext = v;
    }
}

// sums[0] += sum(x), sums[1] += sum(x^2) in float64 (advantage statistics, ValueNorm batch statistics)
// ridx (optional): element i is x[ridx[i] / rdiv] (the returns of a minibatch's agent rows).
__global__ void sum_sumsq_kernel(const float *__restrict__ x, const float *__restrict__ sub, const float *__restrict__ vn,
                                 double *__restrict__ sums, size_t n, const long long *__restrict__ ridx = nullptr,
                                 int rdiv = 1) {
    // value = x[i] - (sub ? denormalize(sub[i]) : 0)
    float mean = 0.f, sd = 1.f;
    if (sub) vn_mean_std(vn, mean, sd);
    double s = 0.0, q = 0.0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = ridx ? (size_t)(ridx[k] / rdiv) : k;
        const float v = sub ? x[i] - (sub[i] * sd + mean) : x[i];
        s += v; q += (double)v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL_MASK, s, o); q += __shfl_xor_sync(FULL_MASK, q, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], s); atomicAdd(&sums[1], q); }
}

// ValueNorm.update with a precomputed batch mean / mean-square (valuenorm.py:38-55); sums = {sum, sumsq}, n rows
// w = float(beta), omw = float(1.0 - beta) with the subtraction done in double on the host, as torch evaluates
// `batch_mean * (1.0 - weight)` (1.0f - 0.99999f would be off by 0.14 %).
__global__ void vn_update_kernel(float *vn, const double *sums, double n, float w, float omw) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const float bm = (float)(sums[0] / n), bs = (float)(sums[1] / n);
        vn[0] = __fadd_rn(__fmul_rn(vn[0], w), __fmul_rn(bm, omw));
        vn[1] = __fadd_rn(__fmul_rn(vn[1], w), __fmul_rn(bs, omw));
        vn[2] = __fadd_rn(__fmul_rn(vn[2], w), omw);
    }
}

struct PpoLossParams {
    float clip, huber_delta, value_coef, inv_rows;  // inv_rows = 1 / (agent rows of the whole (mini)batch, all ranks)
    int n_agents;
    int use_huber, use_clipped;                     // mappo.yaml use_huber_loss / use_clipped_value_loss
};

// Fused PPO loss forward + backward for one chunk (algos/mappo.py:133-169 with cal_value_loss :103-131), split into
// the policy part (needs the actor outputs) and the value part (needs the critic output) so that actor and critic
// can share one activation scratch.  One thread per env-step row r; the N agent rows of r share adv / returns.
// adv_stats = {sum, sumsq} of the raw advantages over all n_adv env-step rows (population std + 1e-5, mappo.py:195-198).
__device__ __forceinline__ float normalized_adv(const float *ret, const float *v_old, const float *vn_gae,
                                                const double *adv_stats, double n_adv, int r) {
    float gm, gs;
    vn_mean_std(vn_gae, gm, gs);
    const double am = adv_stats[0] / n_adv;
    const double avar = fmax(adv_stats[1] / n_adv - am * am, 0.0);
    return (float)(((double)(ret[r] - (v_old[r] * gs + gm)) - am) / (sqrt(avar) + 1e-5));
}

// per-agent-row policy term (mappo.py:150-163): returns dlogp; accumulates the loss / ratio sums
__device__ __forceinline__ float ppo_policy_row(float logp_new, float logp_old, float adv, const PpoLossParams &P, float &pl,
                                                float &rs) {
    const float ratio = expf(logp_new - logp_old);
    const float s1 = ratio * adv;
    const float s2 = fminf(fmaxf(ratio, 1.f - P.clip), 1.f + P.clip) * adv;
    pl += -2.f * fminf(s1, s2);  // two equal log-prob columns (shared_buffer.py:61-62): the loss is 2x
    rs += ratio;
    const bool inrange = (ratio >= 1.f - P.clip) && (ratio <= 1.f + P.clip);
    const float dmin = inrange ? adv : ((s1 < s2) ? adv : 0.f);   // torch.min / clamp sub-gradients
    return -2.f * P.inv_rows * dmin * ratio;
}

__device__ __forceinline__ void ppo_policy_reduce(float pl, float rs, float dls0, float dls1, double *stats, float *dlogstd) {
    double a = pl, c = rs;
    float e0 = dls0, e1 = dls1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL_MASK, a, o); c += __shfl_xor_sync(FULL_MASK, c, o);
        e0 += __shfl_xor_sync(FULL_MASK, e0, o); e1 += __shfl_xor_sync(FULL_MASK, e1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], a); atomicAdd(&stats[2], c);
        atomicAdd(&dlogstd[0], e0); atomicAdd(&dlogstd[1], e1);
    }
}

// policy part: writes dmu [rows*N,2]; accumulates dlogstd[2] and stats {0: policy_loss_sum, 2: ratio_sum} over agent rows.
// logp_new == nullptr: the new log-prob is formed here from mu and the given action (what gauss_finish_kernel would have written:
// the same expression, gauss_logp2) — the fused-head update path then needs no separate log-prob pass.  dout_max (optional):
// atomicMax of the float bits of max |dmu| (first factor of the a-priori bound of the pre-split gradients).
__global__ void ppo_policy_loss_kernel(const float *__restrict__ mu, const float *__restrict__ logp_new,
                                       const float *__restrict__ actions, const float *__restrict__ logstd,
                                       const float *__restrict__ logp_old, const float *__restrict__ ret,
                                       const float *__restrict__ v_old, const float *__restrict__ vn_gae,
                                       const double *__restrict__ adv_stats, double n_adv, float *__restrict__ dmu,
                                       float *__restrict__ dlogstd, double *__restrict__ stats, int rows, PpoLossParams P,
                                       uint32_t *__restrict__ dout_max = nullptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float pl = 0.f, rs = 0.f, dls0 = 0.f, dls1 = 0.f, dm = 0.f;
    if (r < rows) {
        const int N = P.n_agents;
        const float adv = normalized_adv(ret, v_old, vn_gae, adv_stats, n_adv, r);
        const float ls0 = logstd[0], ls1 = logstd[1];
        const float sd0 = expf(ls0), sd1 = expf(ls1);
        const float iv0 = 1.f / (sd0 * sd0), iv1 = 1.f / (sd1 * sd1);
        for (int n = 0; n < N; ++n) {
            const size_t i = (size_t)r * N + n;
            const float a0 = actions[i * 2 + 0], a1 = actions[i * 2 + 1], m0 = mu[i * 2 + 0], m1 = mu[i * 2 + 1];
            const float lpn = logp_new ? logp_new[i] : gauss_logp2(a0, a1, m0, m1, ls0, ls1, sd0, sd1);
            const float dlogp = ppo_policy_row(lpn, logp_old[i], adv, P, pl, rs);
            const float d0 = a0 - m0, d1 = a1 - m1;
            const float g0 = dlogp * d0 * iv0, g1 = dlogp * d1 * iv1;
            dmu[i * 2 + 0] = g0;
            dmu[i * 2 + 1] = g1;
            dm = fmaxf(dm, fmaxf(fabsf(g0), fabsf(g1)));
            dls0 = fmaf(dlogp, d0 * d0 * iv0 - 1.f, dls0);
            dls1 = fmaf(dlogp, d1 * d1 * iv1 - 1.f, dls1);
        }
    }
    ppo_policy_reduce(pl, rs, dls0, dls1, stats, dlogstd);
    if (dout_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmaxf(dm, __shfl_xor_sync(FULL_MASK, dm, o));
        if ((threadIdx.x & 31) == 0 && dm > 0.f) atomicMax(dout_max, __float_as_uint(dm));
    }
}

// minibatch variant (num_mini_batch > 1, shared_buffer.py:219-279): one thread per agent row k of the minibatch;
// ridx[k] = agent-row index into the rollout (actions, logp_old), ridx[k] / N = env-step row (returns, values);
// mu / logp_new / dmu are chunk-local (row k).  ret / v_old / actions / logp_old point at the rollout's row 0.
__global__ void ppo_policy_loss_mb_kernel(const float *__restrict__ mu, const float *__restrict__ logp_new,
                                          const float *__restrict__ actions, const float *__restrict__ logstd,
                                          const float *__restrict__ logp_old, const float *__restrict__ ret,
                                          const float *__restrict__ v_old, const float *__restrict__ vn_gae,
                                          const double *__restrict__ adv_stats, double n_adv,
                                          const long long *__restrict__ ridx, float *__restrict__ dmu,
                                          float *__restrict__ dlogstd, double *__restrict__ stats, int rows, PpoLossParams P) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float pl = 0.f, rs = 0.f, dls0 = 0.f, dls1 = 0.f;
    if (k < rows) {
        const size_t a = (size_t)ridx[k];
        const float adv = normalized_adv(ret, v_old, vn_gae, adv_stats, n_adv, (int)(a / P.n_agents));
        const float sd0 = expf(logstd[0]), sd1 = expf(logstd[1]);
        const float iv0 = 1.f / (sd0 * sd0), iv1 = 1.f / (sd1 * sd1);
        const float dlogp = ppo_policy_row(logp_new[k], logp_old[a], adv, P, pl, rs);
        const float d0 = actions[a * 2 + 0] - mu[(size_t)k * 2 + 0], d1 = actions[a * 2 + 1] - mu[(size_t)k * 2 + 1];
        dmu[(size_t)k * 2 + 0] = dlogp * d0 * iv0;
        dmu[(size_t)k * 2 + 1] = dlogp * d1 * iv1;
        dls0 = dlogp * (d0 * d0 * iv0 - 1.f);
        dls1 = dlogp * (d1 * d1 * iv1 - 1.f);
    }
    ppo_policy_reduce(pl, rs, dls0, dls1, stats, dlogstd);
}

// value part: writes dv [rows]; accumulates stats[1] = value_loss_sum over agent rows.  vn_now = ValueNorm state
// AFTER this epoch's update (cal_value_loss updates before normalising, mappo.py:107-109); nullptr = no normaliser.
// Whole-rollout path (ridx == nullptr): row r is an env-step row standing for its N identical agent rows (weight N).
// Minibatch path: row k is ONE agent row; ret / v_old are read at env-step row ridx[k] / N, weight 1.
__global__ void ppo_value_loss_kernel(const float *__restrict__ ret, const float *__restrict__ v_old,
                                      const float *__restrict__ v_new, const float *__restrict__ vn_now,
                                      float *__restrict__ dv, double *__restrict__ stats, int rows, PpoLossParams P,
                                      const long long *__restrict__ ridx = nullptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float vl = 0.f;
    if (r < rows) {
        const size_t er = ridx ? (size_t)(ridx[r] / P.n_agents) : (size_t)r;
        const float wgt = ridx ? 1.f : (float)P.n_agents;
        float nm, ns;
        vn_mean_std(vn_now, nm, ns);
        const float nret = (ret[er] - nm) / ns;
        const float v = v_new[r], vo = v_old[er];
        const float e = nret - v;
        const float ec = nret - (vo + fminf(fmaxf(v - vo, -P.clip), P.clip));
        const float d = P.huber_delta;
        float h, hc, dh, dhc;
        if (P.use_huber) {
            // one-sided Huber, as the reference computes it (utils/util.py:36-39): zero for e < -d
            h = (fabsf(e) <= d ? e * e * 0.5f : 0.f) + (e > d ? d * (fabsf(e) - 0.5f * d) : 0.f);
            hc = (fabsf(ec) <= d ? ec * ec * 0.5f : 0.f) + (ec > d ? d * (fabsf(ec) - 0.5f * d) : 0.f);
            dh = (fabsf(e) <= d ? e : 0.f) + (e > d ? d : 0.f);
            dhc = (fabsf(ec) <= d ? ec : 0.f) + (ec > d ? d : 0.f);
        } else {   // mse_loss = e^2 / 2 (utils/util.py:42-43)
            h = e * e * 0.5f; hc = ec * ec * 0.5f; dh = e; dhc = ec;
        }
        float w1 = 1.f;                                                   // use_clipped_value_loss = false: original only
        if (P.use_clipped) w1 = h > hc ? 1.f : (h < hc ? 0.f : 0.5f);     // torch.max splits ties evenly
        const float inv = (fabsf(v - vo) <= P.clip) ? 1.f : 0.f;          // clamp passes gradient inside the range
        vl = wgt * (P.use_clipped ? fmaxf(h, hc) : h);
        // whole-rollout path: N identical agent rows per env step, gradient = N * per-row term / B
        dv[r] = (w1 * (-dh) + (1.f - w1) * (-dhc) * inv) * P.inv_rows * wgt * P.value_coef;
    }
    double b = vl;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(FULL_MASK, b, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&stats[1], b);
}

// sum of squares of a flat gradient buffer -> out (float64)
__global__ void sumsq_kernel(const float *__restrict__ g, size_t n, double *__restrict__ out) {
    double q = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        q += (double)g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL_MASK, q, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, q);
}

// clip_grad_norm_(max_norm) + Adam step (torch.optim.Adam with optional L2 weight decay, no amsgrad), fused over the flat buffer.
// sumsq = squared global grad norm of this net (after the cross-GPU all-reduce, if any).
__global__ void clip_adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                 float *__restrict__ v, size_t n, const double *__restrict__ sumsq, float max_norm, float lr,
                                 float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale, int do_clip,
                                 float weight_decay) {
    const float total = (float)sqrt(*sumsq) * grad_scale;
    // use_max_grad_norm = false (mappo.py:179-181): the norm is only reported, the gradients are applied unscaled
    const float coef = (do_clip ? fminf(max_norm / (total + 1e-6f), 1.0f) : 1.0f) * grad_scale;
    const float step = lr / bc1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);   // torch.optim.Adam: grad = grad + weight_decay * param
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

}  // namespace dcc

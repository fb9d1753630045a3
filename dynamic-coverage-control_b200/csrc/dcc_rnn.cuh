// dcc_rnn.cuh — recurrent policies (mappo.yaml: use_recurrent_policy / use_naive_recurrent_policy).
//
// The reference puts an RNNLayer between the MLP trunk and the output head of both nets (algos/r_actor_critic.py:36-37,
// 55-57,102-103,118-120): torch.nn.GRU with recurrent_N layers followed by a LayerNorm (algos/algo_utils/rnn.py:8-22), the
// hidden state multiplied by the step's mask before every step (rnn.py:26-27,66-67).  torch's GRU cell, gate order (r, z, n)
// along the 3H rows of weight_ih / weight_hh:
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))     h' = (1 - z) * n + z * h
// Sequences are processed time-major: row t * S + s = step t of sequence s.  The two matrix products per step are plain
// fp32 GEMMs (gemm_kernel: x W_ih^T for ALL steps of a pass at once, h W_hh^T per step — the recurrence); the kernels here
// are the element-wise halves: masking / gathering the previous state, the gates forward, the gates backward (BPTT), and
// column sums for the bias gradients.  float32 throughout, fmaf-free gate algebra in the reference's operation order.
// The GEMMs run on the tcgen05 kernels when the handle's backend is 2 (hidden 256: fp16 hi/lo split for the forward products —
// trunk outputs and GRU states are bounded —, 3xTF32 for everything that multiplies a gradient) and on the fp32 FFMA kernel
// otherwise.  Not on the benchmarked path (BASELINE configs are MLP policies): written for parity first.
#pragma once
#include "dcc_ops.cuh"

namespace dcc {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// per-row mask of a pass: mask_rows[k] = masks[ridx[k] / rdiv] (ridx NULL: row k itself, k / rdiv)
__global__ void rnn_mask_rows_kernel(const float *__restrict__ masks, const long long *__restrict__ ridx, int rdiv,
                                     float *__restrict__ mask_rows, int rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < rows) mask_rows[k] = masks[(size_t)((ridx ? ridx[k] : (long long)k) / rdiv)];
}

// Hp[s, :] = prev[s, :] * mask[s]   (rnn.py:26-27,66-67).  prev = the previous step's output rows [S, H] (gather == NULL), or the
// stored hidden states [*, RN, H] read at row gather[s] / gdiv (s itself when `identity`), layer `layer`.
__global__ void gru_prev_kernel(const float *__restrict__ prev, const long long *__restrict__ gather, int gdiv, int identity,
                                int RN, int layer, const float *__restrict__ mask, float *__restrict__ Hp, int S, int H) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * H) return;
    const int s = (int)(i / H), c = (int)(i % H);
    float v;
    if (gather || identity) {
        const size_t src = gather ? (size_t)(gather[s] / gdiv) : (size_t)(s / gdiv);
        v = prev[(src * RN + layer) * H + c];
    } else
        v = prev[i];
    Hp[i] = v * mask[s];
}

// gates forward of one step: GI = x W_ih^T, GH = hp W_hh^T (both WITHOUT bias, [S, 3H]); writes the gate activations
// gates[s] = [r | z | n | W_hn hp + b_hn] ([S, 4H], kept for the backward pass) and the new state Hout [S, H]
__global__ void gru_gate_fwd_kernel(const float *__restrict__ GI, const float *__restrict__ GH, const float *__restrict__ b_ih,
                                    const float *__restrict__ b_hh, const float *__restrict__ Hp, float *__restrict__ gates,
                                    float *__restrict__ Hout, int S, int H) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * H) return;
    const size_t s = i / H;
    const int c = (int)(i % H);
    const float *gi = GI + s * 3 * H, *gh = GH + s * 3 * H;
    const float r = sigmoidf_((gi[c] + b_ih[c]) + (gh[c] + b_hh[c]));
    const float z = sigmoidf_((gi[H + c] + b_ih[H + c]) + (gh[H + c] + b_hh[H + c]));
    const float ghn = gh[2 * H + c] + b_hh[2 * H + c];
    const float n = tanhf((gi[2 * H + c] + b_ih[2 * H + c]) + r * ghn);
    const float hp = Hp[i];
    float *g = gates + s * 4 * H;
    g[c] = r; g[H + c] = z; g[2 * H + c] = n; g[3 * H + c] = ghn;
    Hout[i] = (1.f - z) * n + z * hp;
}

// gates backward of one step (BPTT).  dh = dHout[s] + dh_next[s] * mask_next[s] (dh_next / mask_next NULL at the last step:
// the gradient flowing back from step t+1 into h_t passes through h_prev(t+1) = h_t * mask_{t+1}).  Writes the pre-activation
// gradients dGI = [dr', dz', dn'] and dGH = [dr', dz', dn' * r] (both [S, 3H], in place over GI / GH) and the direct part of the
// gradient w.r.t. the masked previous state, dhp = dh * z (the GEMM dGH W_hh is accumulated onto it afterwards).
// dh_next2 (optional): a second addend of the gradient from step t+1 (the dGH W_hh product when it is formed by a GEMM that
// writes rather than accumulates: the tensor-core path).
__global__ void gru_gate_bwd_kernel(const float *__restrict__ dHout, const float *__restrict__ dh_next,
                                    const float *__restrict__ mask_next, const float *__restrict__ gates,
                                    const float *__restrict__ Hp, float *__restrict__ dGI, float *__restrict__ dGH,
                                    float *__restrict__ dhp, int S, int H, const float *__restrict__ dh_next2 = nullptr) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * H) return;
    const size_t s = i / H;
    const int c = (int)(i % H);
    float dh = dHout[i];
    if (dh_next) dh += (dh_next2 ? dh_next[i] + dh_next2[i] : dh_next[i]) * mask_next[s];
    const float *g = gates + s * 4 * H;
    const float r = g[c], z = g[H + c], n = g[2 * H + c], ghn = g[3 * H + c];
    const float dn = dh * (1.f - z);
    const float dz = dh * (Hp[i] - n);
    const float dnp = dn * (1.f - n * n);
    const float drp = dnp * ghn * r * (1.f - r);
    const float dzp = dz * z * (1.f - z);
    float *di = dGI + s * 3 * H, *dg = dGH + s * 3 * H;
    di[c] = drp; di[H + c] = dzp; di[2 * H + c] = dnp;
    dg[c] = drp; dg[H + c] = dzp; dg[2 * H + c] = dnp * r;
    dhp[i] = dh * z;
}

// out[c] += sum_r A[r, c]   (bias gradients); grid.x covers the columns, grid.y splits the rows
__global__ void colsum_atomic_kernel(const float *__restrict__ A, size_t rows, int cols, float *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const size_t per = (rows + gridDim.y - 1) / gridDim.y;
    const size_t r0 = per * blockIdx.y, r1 = r0 + per < rows ? r0 + per : rows;
    float s = 0.f;
    for (size_t r = r0; r < r1; ++r) s += A[r * cols + c];
    if (r1 > r0) atomicAdd(&out[c], s);
}

// dst[(drow(s) * RN + layer) * H + c] = src[s, c]: the new hidden state of a rollout step goes to its slot (layer `layer`)
__global__ void gru_store_state_kernel(const float *__restrict__ src, float *__restrict__ dst, int RN, int layer, int S, int H,
                                       size_t row0) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)S * H) return;
    const size_t s = i / H;
    dst[((row0 + s) * RN + layer) * H + (i % H)] = src[i];
}

}  // namespace dcc

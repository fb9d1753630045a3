// dcc_mappo.cu — C ABI of the MAPPO learner path: policy forward (get_actions / get_values / evaluate_actions),
// GAE returns, and the PPO update split into "epoch gradients" and "clip + Adam apply" so that the host can put one
// NCCL all-reduce of the flat gradient buffers between the two.
//
// Parameters, gradients and Adam moments are CALLER-owned flat float32 buffers (torch tensors) in the reference's
// state_dict order without the never-used fc_h block:
//   [feature_norm.weight, feature_norm.bias, fc1.0.weight (H x in), fc1.0.bias, fc1.2.weight, fc1.2.bias,
//    fc2.0.0.weight (H x H), fc2.0.0.bias, fc2.0.2.weight, fc2.0.2.bias, head.weight (out x H), head.bias, (logstd)]
// The handle owns only activation scratch sized for `chunk_rows` env-step rows; larger batches are streamed in
// chunks with gradient accumulation, which equals the reference's single giant minibatch (num_mini_batch = 1).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>

#include "dcc_compact.cuh"
#include "dcc_ops.cuh"
#include "dcc_rnn.cuh"
#include "dcc_tc.cuh"

namespace dcc {

constexpr int MAX_BLOCKS = 4;   // hidden blocks [Linear, act, LayerNorm]: fc1 + layer_N clones of fc_h (layer_N <= 3)
constexpr int MAX_RNN = 4;      // GRU layers of a recurrent policy (recurrent_N, rnn.py:13)

struct NetLayout {
    int in, H, out;
    int inp;   // `in` rounded up to a multiple of 32: leading dimension of the xhat scratch (zero-padded)
    int nblk;  // 1 + layer_N hidden blocks; block 0 = fc1 (in -> H), blocks 1.. = fc2[i] (H -> H)  (mlp.py:16-29)
    bool has_ln0;   // use_feature_normalization: feature_norm.weight / .bias present
    size_t ln0_g, ln0_b, W[MAX_BLOCKS], b[MAX_BLOCKS], lg[MAX_BLOCKS], lb[MAX_BLOCKS], Wh, bh, logstd, total;
    // recurrent policies: rn GRU layers (weight_ih [3H, H], weight_hh [3H, H], bias_ih [3H], bias_hh [3H] each) and the
    // RNNLayer's LayerNorm, between the trunk and the head — the reference's state_dict order (base.*, rnn.*, act.* / v_out.*)
    int rn;
    size_t Wih[MAX_RNN], Whh[MAX_RNN], bih[MAX_RNN], bhh[MAX_RNN], rnn_g, rnn_b;
    void init(int in_, int H_, int out_, bool has_logstd, bool has_ln0_, int layer_N, int recurrent_N = 0) {
        in = in_; H = H_; out = out_; inp = (in_ + 31) / 32 * 32; has_ln0 = has_ln0_; nblk = 1 + layer_N; rn = recurrent_N;
        size_t o = 0;
        ln0_g = o; if (has_ln0) o += in;
        ln0_b = o; if (has_ln0) o += in;
        for (int k = 0; k < nblk; ++k) {
            W[k] = o; o += (size_t)H * (k == 0 ? in : H);
            b[k] = o; o += H; lg[k] = o; o += H; lb[k] = o; o += H;
        }
        for (int l = 0; l < rn; ++l) {
            Wih[l] = o; o += (size_t)3 * H * H; Whh[l] = o; o += (size_t)3 * H * H; bih[l] = o; o += 3 * H; bhh[l] = o; o += 3 * H;
        }
        rnn_g = o; if (rn) o += H;
        rnn_b = o; if (rn) o += H;
        Wh = o; o += (size_t)out * H; bh = o; o += out;
        logstd = o; if (has_logstd) o += out;
        total = o;
    }
};

struct MappoHandle {
    uint32_t magic;
    dcc_mappo_cfg cfg;
    int device, sm_count;
    int backend;     // 1 = SIMT fp32, 2 = tcgen05 3xTF32
    bool f16_fwd;    // backend 2: forward GEMMs on LayerNorm outputs use the fp16 hi/lo split kernel (DCC_TC_F16=0 disables)
    bool f16_wgrad;  // backend 2, experimental: fp16-split weight-gradient GEMMs (DCC_TC_WGRAD_F16=1 enables; measured no faster)
    bool f16_dx;     // backend 2: the backward dX = dZ W GEMMs run the fp16-split kernel with a per-tensor power-of-two scale of
                     // dZ (its max |dZ| is produced for free by the LayerNorm-backward kernel that writes dZ); DCC_TC_DX_F16=0 disables
    bool ln_pipe;    // LayerNorm-backward kernels with the bulk-async row pipeline (H % 4 == 0); DCC_LN_PIPE=0 disables
    bool split16;    // backend 2: activations that only GEMMs consume (inner trunk outputs h_k, compact features) are stored pre-split
                     // as fp16 hi/lo by their producers and fetched by TMA in the consumers (tc::TcfParams::a_split); DCC_TC_SPLIT=0 disables
    uint32_t *dz_absmax;   // device scalar: bits of max |dZ| for the fp16-split weight-gradient kernel
    NetLayout la, lc;
    int chunk_rows;  // env-step rows per chunk
    // scratch
    float *x0;       // [chunk*N, D] == [chunk, N*D]: xhat = input LayerNorm without affine
    float *a[MAX_BLOCKS], *hh[MAX_BLOCKS];          // per hidden block: post-activation a_k and h_k = LN(a_k), [chunk*N, H]
    float *mean[MAX_BLOCKS], *rstd[MAX_BLOCKS];      // per hidden block: LayerNorm row statistics, [chunk*N]
    float *dA, *dB;                                  // [chunk*N, H] gradient ping-pong
    float *w1g_a, *b1g_a, *w1g_c, *b1g_c;   // fc1 weights with the input LayerNorm affine folded in (per call)
    // tcgen05 backend: hi/lo-split, 128B-swizzled shared-memory images of the weights (per call), [actor, critic]
    float *img_w1[2], *img_w[2][MAX_BLOCKS], *img_wt[2][MAX_BLOCKS];   // blocks >= 1: forward / transposed (dX) images
    float *mu, *logp, *dmu, *vnew, *dv;   // [chunk*N,2], [chunk*N], [chunk*N,2], [chunk], [chunk]
    double *dsums;   // small float64 scratch: [4] actor grad sumsq, [5] critic grad sumsq
    float *vn_gae;   // ValueNorm state snapshot taken at train_begin (3 floats)
    // compact-state path (dcc_mappo_set_env_layout): layer 1 evaluated from the env's compact state (dcc_compact.cuh)
    bool compact;
    CompactDims cd;
    double *d_poi;           // [M, 2] PoI table
    float *fc;               // [chunk, cd.ldc] critic features (the actor's [chunk*N, cd.lda] live in x0)
    float *head_fold[2];     // output head folded through the last LayerNorm affine, for the fused epilogue (head_fold_kernel)
    // "xhat mode" (backend 2, pre-split operands, ReLU trunk): an inner block k < last stores ONLY xhat_k = (a_k - mean) rstd, pre-split
    // (no a_k, no affine), and block k+1 runs on weights with the LayerNorm affine folded in, W_{k+1} * gamma_k and
    // b_{k+1} + W_{k+1} beta_k — the trick of the input LayerNorm (§5.1) applied to every inner LayerNorm.  Halves the bytes the
    // fused forward epilogue has to push out (its bound); the backward pass rebuilds the ReLU mask from xhat (relu_lnx_bwd_pipe_kernel)
    // and gets dgamma_k / dbeta_k from G_{k+1} = dz_{k+1}^T xhat_k once per optimiser step (ln0_finalize_kernel).  DCC_TC_XHAT=0 disables.
    bool xhat;
    // pre-split gradients (xhat mode): dz_k leaves the LayerNorm-backward kernels as fp16 hi/lo, scaled by a power of two derived from
    // an a-priori bound of |dz_k|, so the weight-gradient and dX GEMMs are fed by TMA alone.  `sc` = device scalars
    // (float bits): [SC_DOUT] max |dout| of the pass, [SC_RSTD + k] max rstd of block k, [SC_DXH + k] max |d xhat_k|, [SC_BND] bound of
    // the dz tensor currently alive.  DCC_TC_DZSPLIT=0 disables.
    bool dzsplit;
    bool dout_ready;     // the loss kernel of this pass already left max |dout| in sc[SC_DOUT] (else trunk_backward_dzs reduces it)
    uint32_t *sc;
    float *wf[2][MAX_BLOCKS], *bf[2][MAX_BLOCKS];   // folded weights / biases of blocks >= 1 [actor, critic]
    float *ones, *zeros;                            // [H]: unit LayerNorm affine handed to the epilogue of inner blocks
    float *wt[2];            // folded fc1 weights [H, ld] (actor, critic)
    float *gt[2];            // running dz1^T f of the epoch [H, ld]
    // recurrent policies (dcc_rnn.cuh): per GRU layer the masked previous states, x W_ih^T / h W_hh^T (overwritten by their
    // gradients in the backward pass), gate activations and outputs of every step of a pass; [rows, *] with rows <= chunk * N
    float *r_Hp[MAX_RNN], *r_GI[MAX_RNN], *r_GH[MAX_RNN], *r_gates[MAX_RNN], *r_Hout[MAX_RNN];
    float *r_dH[2], *r_Y, *r_mean, *r_rstd, *r_mask, *r_dhp[2], *r_dfeat, *r_zero, *r_dummy;
    // tcgen05 backend: weight images of the GRU matrices (per net and layer, rebuilt once per ABI call): the three 256-row gate
    // slices of weight_ih / weight_hh for the forward products (fp16 split) and the two matrices transposed for dGI W_ih / dGH W_hh
    // (3xTF32, K = 768); r_dT: the per-step dGH W_hh products (the tensor-core GEMM writes, it does not accumulate)
    float *r_img_ih[2][MAX_RNN][3], *r_img_hh[2][MAX_RNN][3], *r_img_ihT[2][MAX_RNN], *r_img_hhT[2][MAX_RNN], *r_dT[2];
    int64_t launches;
};
enum { SC_DOUT = 0, SC_RSTD = 1, SC_DXH = 1 + MAX_BLOCKS, SC_BND = 1 + 2 * MAX_BLOCKS, SC_COUNT = 2 + 2 * MAX_BLOCKS };
constexpr uint32_t MAPPO_MAGIC = 0xDCCA0002u;

// ValueNorm state as of train_begin (advantages use it), or nullptr without a value normaliser
static inline const float *vn_snapshot(const MappoHandle *h) { return h->cfg.use_valuenorm ? h->vn_gae : nullptr; }

static inline int act_of(const MappoHandle *h) { return h->cfg.use_relu ? ACT_RELU : ACT_TANH; }

// tcgen05 backend: the output heads are fused into the last block's GEMM epilogue (TcfParams::head_*); DCC_TC_HEAD=0 keeps
// the separate head kernels (tuning / A-B knob)
static inline bool fused_head(const MappoHandle *h) {
    static const bool on = !(getenv("DCC_TC_HEAD") && atoi(getenv("DCC_TC_HEAD")) == 0);
    return on && h->backend == 2;
}

static MappoHandle *as_mappo(void *h) {
    MappoHandle *m = static_cast<MappoHandle *>(h);
    return (m && m->magic == MAPPO_MAGIC) ? m : nullptr;
}

// the tcgen05 3xTF32 kernels cover the shipped trunk width only
static inline bool tc_supported(const dcc_mappo_cfg *c) { return c->hidden == tc::TC_N; }

static inline int grid_for_rows(const MappoHandle *h, long rows, int warps_per_block) {
    long b = (rows + warps_per_block - 1) / warps_per_block;
    const long cap = (long)h->sm_count * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// kernels that end with one set of atomics per block into a small parameter-gradient vector: 4 CTAs per SM
static inline int grid_for_reduce(const MappoHandle *h, long rows, int warps_per_block, int ctas_per_sm = 4) {
    long b = (rows + warps_per_block - 1) / warps_per_block;
    const long cap = (long)h->sm_count * ctas_per_sm;      // grid-stride kernels: exactly one wave of resident CTAs
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// C (+)= op(A) op(B), see gemm_kernel.  accumulate != 0 adds into C (float atomics); split-K is used when the
// output grid alone cannot fill the GPU (the weight-gradient GEMMs: tiny M x N, huge K).
static int launch_gemm(MappoHandle *h, bool ta, bool tb, int M, int N, int K, const float *A, int lda, const float *B,
                       int ldb, float *C, int ldc, bool accumulate, cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return DCC_OK;
    const int gx = (N + GM_BN - 1) / GM_BN, gy = (M + GM_BM - 1) / GM_BM;
    int splits = 1;
    const long tiles = (long)gx * gy;
    if (tiles < 2L * h->sm_count && K >= 8 * GM_BK) {
        splits = (int)((4L * h->sm_count + tiles - 1) / tiles);
        const int max_splits = (K + 4 * GM_BK - 1) / (4 * GM_BK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int kps = (K + splits - 1) / splits;
    kps = (kps + GM_BK - 1) / GM_BK * GM_BK;
    splits = (K + kps - 1) / kps;
    const bool atomic = accumulate || splits > 1;
    if (atomic && !accumulate) DCC_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, s));
    dim3 grid(gx, gy, splits), block(256);
#define DCC_GEMM_CASE(TA_, TB_)                                                                                   \
    if (ta == TA_ && tb == TB_) {                                                                                 \
        if (atomic) gemm_kernel<TA_, TB_, true><<<grid, block, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, kps);    \
        else gemm_kernel<TA_, TB_, false><<<grid, block, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, kps);          \
    }
    DCC_GEMM_CASE(false, false)
    DCC_GEMM_CASE(false, true)
    DCC_GEMM_CASE(true, false)
    DCC_GEMM_CASE(true, true)
#undef DCC_GEMM_CASE
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

// ---- tcgen05 backend launchers ---------------------------------------------------------------------------------
// power-of-two scale of the fp16-split weight images (keeps the lo halves normal; |W| < 255 stays finite)
constexpr float TC_F16_WSCALE = 256.f;

static int tc_prep_weights(MappoHandle *h, const float *W, int ldw, bool transposed, int K, float *img, cudaStream_t s,
                           bool f16 = false) {
    const int bk = f16 ? tc::TC_BK16 : tc::TC_BK;
    const int KT = (K + bk - 1) / bk;
    const int n = KT * tc::TC_N * 8;
    if (f16) tc::tc_prep_weights_f16_kernel<<<(n + 255) / 256, 256, 0, s>>>(W, ldw, transposed ? 1 : 0, K, KT, TC_F16_WSCALE, img);
    else tc::tc_prep_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(W, ldw, transposed ? 1 : 0, K, KT, img);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

// per-device opt-in of the row-pipeline kernels to their dynamic shared memory (ring of RP_SLOTS rows per warp; with the
// static mbarriers it crosses the 48 KB default at H = 256)
static int pipe_set_kernel_attributes() {
    const int bytes = 64 * 1024;
    DCC_CUDA_TRY(cudaFuncSetAttribute(relu_ln_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute(relu_lnx_bwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute(relu_lnx_bwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute((relu_lnx_bwd_pipe_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute((head_relu_ln_bwd_pipe_kernel<1, RP_SLOTS, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute((head_relu_ln_bwd_pipe_kernel<2, RP_SLOTS, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute((head_relu_ln_bwd_pipe_kernel<1, RP_SLOTS, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute((head_relu_ln_bwd_pipe_kernel<2, RP_SLOTS, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute(head_relu_ln_bwd_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    DCC_CUDA_TRY(cudaFuncSetAttribute(head_relu_ln_bwd_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return DCC_OK;
}

// per-device opt-in of the tcgen05 kernels to their dynamic shared-memory footprint (current device)
static int tc_set_kernel_attributes() {
    DCC_CUDA_TRY(cudaFuncSetAttribute(tc::tc_gemm_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TCF_SMEM_BYTES));
    DCC_CUDA_TRY(cudaFuncSetAttribute(tc::tc_gemm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TCF_SMEM_BYTES));
    DCC_CUDA_TRY(cudaFuncSetAttribute(tc::tc_gemm_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TCF_SMEM_BYTES));
    DCC_CUDA_TRY(cudaFuncSetAttribute(tc::tc_gemm_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TCF_SMEM_BYTES));
    DCC_CUDA_TRY(cudaFuncSetAttribute(tc::tc_gemm_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TCF_SMEM_BYTES));
    return DCC_OK;
}

// C[M,256] = A[M,K] * B^T with B given as a weight image (K-major on both sides).  With `bias` the epilogue is the
// fused bias + ReLU + LayerNorm of an MLP block: a -> a_out (optional), LN(a) * gamma + beta -> h_out.
// An activation stored PRE-SPLIT as two fp16 matrices (hi = fp16(x), lo = fp16(x - hi)), row pitch `ld` halves: what the GEMM
// kernels fetch with TMA instead of splitting fp32 values in their producer warps (tc::TcfParams::a_split).
struct Split16 {
    const void *hi, *lo;
    int ld;
};

// a16 != nullptr: the A operand is pre-split (A / lda ignored; fp16-split kernel only).  h16 != nullptr (fused epilogue): h leaves
// pre-split into h16 instead of fp32 h_out.
static int tc_gemm_fwd(MappoHandle *h, int M, int K, const float *A, int lda, const float *img, float *C, int ldc,
                       cudaStream_t s, const float *bias = nullptr, const float *gamma = nullptr,
                       const float *beta = nullptr, float *h_out = nullptr, float *mean = nullptr, float *rstd = nullptr,
                       bool f16 = false, const uint32_t *a_absmax_bits = nullptr, const float *head_fold = nullptr,
                       int head_out = 0, float *head_dst = nullptr, const Split16 *a16 = nullptr, const Split16 *h16 = nullptr,
                       uint32_t *rstd_max_out = nullptr, uint32_t *c_absmax_out = nullptr) {
    if (M <= 0) return DCC_OK;
    if ((ldc & 3) || (K & 3)) return DCC_ERR_INVALID_ARG;
    if (!a16 && ((lda & 3) || ((uintptr_t)A & 15))) return DCC_ERR_INVALID_ARG;
    // a16 with a_absmax_bits: the pre-split operand was stored ALREADY multiplied by the power of two of those bits (pre-split dz)
    if (a16 && !f16) return DCC_ERR_INVALID_ARG;
    tc::TcfParams p;
    memset(&p, 0, sizeof p);
    const int bk = f16 ? tc::TC_BK16 : tc::TC_BK;    // `img` must come from tc_prep_weights with the same f16 flag
    p.A = A; p.Bimg = img; p.C = C; p.M = M; p.K = K; p.KT = (K + bk - 1) / bk; p.lda = lda; p.ldc = ldc;
    p.out_scale = f16 ? 1.f / TC_F16_WSCALE : 1.f;
    p.a_absmax_bits = f16 ? a_absmax_bits : nullptr;
    p.epi = bias ? tc::TCF_EPI_BIAS_RELU_LN : tc::TCF_EPI_STORE;
    p.bias = bias; p.gamma = gamma; p.beta = beta; p.H = h_out; p.mean = mean; p.rstd = rstd;
    p.act = act_of(h);
    if (bias && head_dst && head_out > 0) { p.head_out = head_out; p.head_fold = head_fold; p.head_dst = head_dst; }
    if (a16) {
        p.a_split = 1;
        if (!tc::tc_make_map_2d_f16(&p.tmAhi, a16->hi, K, M, a16->ld, tc::TC_BK16, tc::TC_BM, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc::tc_make_map_2d_f16(&p.tmAlo, a16->lo, K, M, a16->ld, tc::TC_BK16, tc::TC_BM, CU_TENSOR_MAP_SWIZZLE_128B))
            return DCC_ERR_UNSUPPORTED;
    }
    p.unit_affine = (h->xhat && gamma == h->ones) ? 1 : 0;      // xhat mode, inner block
    p.rstd_max_out = bias ? rstd_max_out : nullptr;
    p.c_absmax_out = bias ? nullptr : c_absmax_out;
    if (h16 && bias) {
        p.h_split = 1;
        p.H = reinterpret_cast<float *>(const_cast<void *>(h16->hi));     // non-NULL: "store h"
        if (!tc::tc_make_map_2d_f16(&p.tmHhi, h16->hi, tc::TC_N, M, h16->ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::tc_make_map_2d_f16(&p.tmHlo, h16->lo, tc::TC_N, M, h16->ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))
            return DCC_ERR_UNSUPPORTED;
    }
    const int row_tiles = (M + tc::TC_BM - 1) / tc::TC_BM;
    // split-K only to fill the GPU when there are few row tiles and a long K (raw-store epilogue only)
    p.splits = 1;
    if (!bias && row_tiles < 2 * h->sm_count && p.KT >= 16) {
        p.splits = (3 * h->sm_count + row_tiles - 1) / row_tiles;
        if (p.splits > p.KT / 8) p.splits = p.KT / 8;
        if (p.splits < 1) p.splits = 1;
    }
    p.kt_per_split = (p.KT + p.splits - 1) / p.splits;
    p.splits = (p.KT + p.kt_per_split - 1) / p.kt_per_split;
    if (p.splits > 1) DCC_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)tc::TC_N * 4, M, s));
    // epilogue stores through the TMA engine (see TcfParams::use_tma) unless split-K needs atomics; DCC_TC_TMA=0 keeps
    // the st.global epilogue (tuning knob)
    static const bool tma_env = !(getenv("DCC_TC_TMA") && atoi(getenv("DCC_TC_TMA")) == 0);
    if (tma_env && p.splits == 1) {
        bool ok = true;
        if (C) ok = ok && tc::tc_make_store_map(&p.tmC, C, M, ldc);
        if (bias && h_out && !p.h_split) ok = ok && tc::tc_make_store_map(&p.tmH, h_out, M, ldc);
        p.use_tma = ok ? 1 : 0;
    }
    if (p.h_split && !p.use_tma) return DCC_ERR_UNSUPPORTED;      // the pre-split output exists only on the TMA store path
    // L2 prefetch of the activation tiles two stages (64 KB per SM) ahead of the fp16-split kernel's loads; measured:
    // 2 stages > 4 > none > 8, and no gain for the MMA-bound 3xTF32 kernel (DCC_TC_PF=<stages> overrides, 0 = off)
    static const int pf_env = getenv("DCC_TC_PF") ? atoi(getenv("DCC_TC_PF")) : -1;
    p.pf_dist = a16 ? 0 : (pf_env >= 0 ? pf_env : (f16 ? 2 : 0));
    if (p.pf_dist > 0 && !tc::tc_make_prefetch_map(&p.tmA, A, M, K, lda, bk)) p.pf_dist = 0;
    const int work = row_tiles * p.splits;
    const int grid = work < h->sm_count ? work : h->sm_count;
    if (f16 && a_absmax_bits) tc::tc_gemm_fwd_kernel<true, true><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    else if (f16) tc::tc_gemm_fwd_kernel<true><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    else tc::tc_gemm_fwd_kernel<false><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

// G[256, Nout] += dZ[R,256]^T X[R, Nout]  (tcgen05, MN-major operands; G must already hold the running sum).
// f16 = the experimental fp16-split variant (X must be a LayerNorm output; one absmax pass over dZ supplies its scale).
// x16 != nullptr: X is pre-split (X / ldx ignored; forces the fp16-split kernel).  absmax_ready: h->dz_absmax already holds
// max |dZ| (left by the kernel that wrote dZ), so the separate pass over dZ is skipped.
// z16 != nullptr (with x16): dZ is pre-split and pre-scaled as well (dZ / ldz ignored); bnd_bits = the device scalar whose power of
// two it was stored with.
static int tc_gemm_wgrad(MappoHandle *h, int R, int Nout, const float *dZ, int ldz, const float *X, int ldx, float *G,
                         int ldg, cudaStream_t s, bool f16 = false, const Split16 *x16 = nullptr, bool absmax_ready = false,
                         const Split16 *z16 = nullptr, const uint32_t *bnd_bits = nullptr) {
    if (R <= 0 || Nout <= 0) return DCC_OK;
    if (z16 && (!x16 || !bnd_bits)) return DCC_ERR_INVALID_ARG;
    if (!z16 && ((ldz & 3) || ((uintptr_t)dZ & 15))) return DCC_ERR_INVALID_ARG;
    if (!x16 && ((ldx & 3) || ((uintptr_t)X & 15))) return DCC_ERR_INVALID_ARG;
    if (x16) f16 = true;
    if (f16 && !h->dz_absmax) { if (x16) return DCC_ERR_UNSUPPORTED; f16 = false; }
    tc::TcwParams p;
    memset(&p, 0, sizeof p);
    p.dZ = dZ; p.X = X; p.G = G; p.R = R; p.Nout = Nout; p.ldz = ldz; p.ldx = ldx; p.ldg = ldg;
    p.n_tiles = (Nout + tc::TC_N - 1) / tc::TC_N;
    p.tile_n = ((Nout + p.n_tiles - 1) / p.n_tiles + 31) / 32 * 32;      // balanced column tiles, whole 32-feature groups
    const int out_tiles = 2 * p.n_tiles;
    const int bkw = f16 ? 64 : tc::TC_BK;                                // batch rows per stage
    // split the batch rows so that the work items fill two waves of the persistent grid without spilling into a
    // third (floor, not ceil), each split at least 8 stages long
    int ks = (2 * h->sm_count) / out_tiles;
    const int max_ks = (R + 8 * bkw - 1) / (8 * bkw);
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    p.rows_per_split = ((R + ks - 1) / ks + bkw - 1) / bkw * bkw;
    p.ksplits = (R + p.rows_per_split - 1) / p.rows_per_split;
    const int work = out_tiles * p.ksplits;
    const int grid = work < h->sm_count ? work : h->sm_count;
    if (x16) {
        p.x_split = 1;
        if (!tc::tc_make_map_2d_f16(&p.tmXhi, x16->hi, Nout, R, x16->ld, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc::tc_make_map_2d_f16(&p.tmXlo, x16->lo, Nout, R, x16->ld, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B))
            return DCC_ERR_UNSUPPORTED;
    }
    if (z16) {
        p.dz_split = 1;
        if (!tc::tc_make_map_2d_f16(&p.tmZhi, z16->hi, tc::TC_N, R, z16->ld, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc::tc_make_map_2d_f16(&p.tmZlo, z16->lo, tc::TC_N, R, z16->ld, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B))
            return DCC_ERR_UNSUPPORTED;
        p.dz_absmax_bits = bnd_bits;
        tc::tc_gemm_wgrad_kernel<true><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    } else if (f16) {
        if (!absmax_ready) {
            DCC_CUDA_TRY(cudaMemsetAsync(h->dz_absmax, 0, sizeof(uint32_t), s));
            tc::tc_absmax_bits_kernel<<<h->sm_count * 8, 256, 0, s>>>(dZ, (long)R, tc::TC_N, ldz, h->dz_absmax);
            h->launches++;
        }
        p.dz_absmax_bits = h->dz_absmax;
        tc::tc_gemm_wgrad_kernel<true><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    } else {
        tc::tc_gemm_wgrad_kernel<false><<<grid, tc::TCF_THREADS, tc::TCF_SMEM_BYTES, s>>>(p);
    }
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

// Which forward GEMMs run the fp16-split kernel: those whose activation operand is a LayerNorm output (block 0 only
// when the input LayerNorm is on: |xhat| <= sqrt(D); blocks >= 1 always read LN(a) * gamma + beta), i.e. values far
// inside fp16's range.  Raw observations and the backward dX GEMMs (unbounded dynamic range) stay on 3xTF32.
static inline bool fwd_f16(const MappoHandle *h, const NetLayout &L, int k) { return h->f16_fwd && (k > 0 || L.has_ln0); }
// same condition on the X operand of the weight-gradient GEMM of block k (experimental, DCC_TC_WGRAD_F16=1)
static inline bool wgrad_f16(const MappoHandle *h, const NetLayout &L, int k) { return h->f16_wgrad && (k > 0 || L.has_ln0); }

// output head folded through the last block's LayerNorm affine (fused epilogue, see TcfParams::head_fold); once per ABI call
static int fold_head(MappoHandle *h, const NetLayout &L, const float *P, int net, cudaStream_t s) {
    if (!fused_head(h)) return DCC_OK;
    const int last = L.nblk - 1;
    head_fold_kernel<<<1, 256, 0, s>>>(P + L.lg[last], P + L.lb[last], P + L.Wh, P + L.bh, L.H, L.out, h->head_fold[net]);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// weight images of the blocks >= 1 (forward, and transposed for dX).  xhat mode: block k runs on the previous block's UN-affined
// LayerNorm output, so its weights carry that LayerNorm's affine: W' = W_k * gamma_{k-1}, b' = b_k + W_k beta_{k-1} (fold_ln0_kernel)
static int prep_inner_images(MappoHandle *h, const NetLayout &L, const float *P, int net, bool for_backward, cudaStream_t s) {
    int rc;
    for (int k = 1; k < L.nblk; ++k) {
        const float *Wk = P + L.W[k];
        if (h->xhat) {
            fold_ln0_kernel<<<(L.H + 7) / 8, 256, 0, s>>>(P + L.W[k], P + L.b[k], P + L.lg[k - 1], P + L.lb[k - 1], h->wf[net][k],
                                                          h->bf[net][k], L.H, L.H);
            h->launches++;
            DCC_CUDA_TRY(cudaGetLastError());
            Wk = h->wf[net][k];
        }
        if ((rc = tc_prep_weights(h, Wk, L.H, false, L.H, h->img_w[net][k], s, fwd_f16(h, L, k)))) return rc;
        if (for_backward && (rc = tc_prep_weights(h, Wk, L.H, true, L.H, h->img_wt[net][k], s, h->f16_dx))) return rc;
    }
    return DCC_OK;
}

// fold the input LayerNorm affine into fc1 (done once per ABI call: the parameters change after every apply)
static int fold_ln0(MappoHandle *h, const NetLayout &L, const float *P, int net, bool for_backward, cudaStream_t s) {
    float *w1g = net ? h->w1g_c : h->w1g_a, *b1g = net ? h->b1g_c : h->b1g_a;
    fold_ln0_kernel<<<(L.H + 7) / 8, 256, 0, s>>>(P + L.W[0], P + L.b[0], L.has_ln0 ? P + L.ln0_g : nullptr,
                                                  L.has_ln0 ? P + L.ln0_b : nullptr, w1g, b1g, L.H, L.in);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    if (h->backend == 2) {
        int rc;
        if ((rc = fold_head(h, L, P, net, s))) return rc;
        if ((rc = tc_prep_weights(h, w1g, L.in, false, L.in, h->img_w1[net], s, fwd_f16(h, L, 0)))) return rc;
        if ((rc = prep_inner_images(h, L, P, net, for_backward, s))) return rc;
    }
    return DCC_OK;
}

// pre-split storage (h->split16): hi halves at the start of the fp32-sized buffer, lo halves behind them (capacity-based offset)
static inline Split16 hh_split(const MappoHandle *h, int k) {
    const size_t cap = (size_t)h->chunk_rows * h->cfg.n_agents * h->cfg.hidden;
    return Split16{h->hh[k], reinterpret_cast<const __half *>(h->hh[k]) + cap, h->cfg.hidden};
}
// Critic rows one pass may hold on the compact path.  The activation scratch is sized for the actor's chunk * N agent rows,
// the critic runs ONE row per env step: it is therefore fed N chunks' worth of env steps at a time (a "super-chunk"), so that its
// GEMMs see the same 16 waves of row tiles as the actor's instead of 2 (per tile they ran 2x slower: pipeline fill, weight
// fetch and launch latency of a 2-wave grid).  The feature matrix `fc`, `vnew` and `dv` are sized to match.
static inline size_t critic_cap_rows(const MappoHandle *h) { return (size_t)h->chunk_rows * h->cfg.n_agents; }

// a gradient buffer (dA / dB, fp32-sized) holding a PRE-SPLIT dz: hi halves first, lo halves behind them (row pitch H halves)
static inline Split16 dz_split16(const MappoHandle *h, const float *buf) {
    const size_t cap = (size_t)h->chunk_rows * h->cfg.n_agents * h->cfg.hidden;
    return Split16{buf, reinterpret_cast<const __half *>(buf) + cap, h->cfg.hidden};
}
static inline Split16 feat_split(const MappoHandle *h, int net) {
    if (net == 0) {
        const size_t cap = (size_t)h->chunk_rows * h->cfg.n_agents * h->cd.lda;
        return Split16{h->x0, reinterpret_cast<const __half *>(h->x0) + cap, h->cd.lda};
    }
    const size_t cap = critic_cap_rows(h) * h->cd.ldc;
    return Split16{h->fc, reinterpret_cast<const __half *>(h->fc) + cap, h->cd.ldc};
}

// compact path: fold the input LayerNorm affine AND the observation structure into fc1 (Wt = (W1 * gamma0) A)
static int fold_compact(MappoHandle *h, const NetLayout &L, const float *P, int net, bool for_backward, cudaStream_t s) {
    float *b1g = net ? h->b1g_c : h->b1g_a;
    const int ldk = net ? h->cd.ldc : h->cd.lda, nb = net ? h->cd.N : 1;
    fold_compact_kernel<<<(L.H + 7) / 8, 256, 0, s>>>(P + L.W[0], P + L.b[0], L.has_ln0 ? P + L.ln0_g : nullptr,
                                                      L.has_ln0 ? P + L.ln0_b : nullptr, h->d_poi, h->wt[net], b1g, L.H, h->cd, nb, ldk);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    if (h->backend == 2) {
        int rc;
        if ((rc = fold_head(h, L, P, net, s))) return rc;
        if ((rc = tc_prep_weights(h, h->wt[net], ldk, false, ldk, h->img_w1[net], s, h->f16_fwd))) return rc;
        if ((rc = prep_inner_images(h, L, P, net, for_backward, s))) return rc;
    }
    return DCC_OK;
}

// compact path: features of `rows` env-step rows -> x0 (actor, [rows*N, lda]) and fc (critic, [rows, ldc])
// ridx != nullptr (minibatch path): output row k comes from agent row ridx[k] (see compact_features_kernel)
// fc_row0: the critic rows land at row fc_row0 of `fc` (super-chunks: N actor chunks fill one critic pass, critic_cap_rows)
static int compact_features(MappoHandle *h, const double *pv, const uint8_t *en, int rows, bool want_actor, bool want_critic,
                            cudaStream_t s, const long long *ridx = nullptr, size_t fc_row0 = 0) {
    const int wpb = 8;
    const size_t smem = compact_features_smem(h->cd, wpb);
    const size_t lo_a = h->split16 ? (size_t)h->chunk_rows * h->cfg.n_agents * h->cd.lda : 0;     // halves; 0 = fp32 rows
    const size_t lo_c = h->split16 ? critic_cap_rows(h) * h->cd.ldc : 0;
    if (fc_row0 + (size_t)rows > critic_cap_rows(h)) return DCC_ERR_INVALID_ARG;
    // row offset into fc: rows are ldc halves apart when pre-split (hi block first), ldc floats otherwise
    float *fc_dst = h->split16 ? reinterpret_cast<float *>(reinterpret_cast<__half *>(h->fc) + fc_row0 * h->cd.ldc)
                               : h->fc + fc_row0 * h->cd.ldc;
    compact_features_kernel<<<grid_for_rows(h, rows, wpb), wpb * 32, smem, s>>>(pv, en, want_actor ? h->x0 : nullptr,
                                                                              want_critic ? fc_dst : nullptr, rows, h->cd,
                                                                              h->cfg.use_feature_normalization ? 1 : 0, lo_a, lo_c, ridx);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// trunk forward on `rows` rows of width L.in: x -> h2.  save = keep a1/a2/stats for the backward pass.
// feat != nullptr (compact path): block 0 reads the precomputed feature rows [rows, ldf] against the folded weights
// h->wt[net] instead of xhat / w1g (x is ignored; no input-LayerNorm kernel).
static int trunk_forward(MappoHandle *h, const NetLayout &L, const float *P, int net, const float *x, int rows, bool save,
                         cudaStream_t s, const long long *ridx = nullptr, int rdiv = 1, const float *feat = nullptr, int ldf = 0,
                         float *head_dst = nullptr) {
    // head_dst (tcgen05 backend only): the output head (action mean [rows, 2] / value [rows]) is evaluated inside the last
    // block's GEMM epilogue and h_last is NOT stored (fused_head()).
    const int H = L.H;
    const int wpb = 8;
    const float *w1g = feat ? h->wt[net] : (net ? h->w1g_c : h->w1g_a), *b1g = net ? h->b1g_c : h->b1g_a;
    const bool dzs = h->dzsplit && save;      // training pass: the forward epilogues leave max rstd per block for the dz bounds
    if (dzs) DCC_CUDA_TRY(cudaMemsetAsync(h->sc, 0, SC_COUNT * sizeof(uint32_t), s));
    // input LayerNorm (without affine): register-resident single-pass variant when the rows are 16-byte aligned
    if (feat) {
        // features carry the normalisation already
    } else if ((L.in & 3) == 0 && ((uintptr_t)x & 15) == 0 && L.in <= 128 * 8)
        ln_noaffine_fwd_vec_kernel<8><<<grid_for_rows(h, rows, wpb), wpb * 32, 0, s>>>(x, h->x0, rows, L.in, L.inp, ridx, rdiv, L.has_ln0 ? 1 : 0);
    else if ((L.in & 3) == 0 && ((uintptr_t)x & 15) == 0 && L.in <= 128 * 24)
        ln_noaffine_fwd_vec_kernel<24><<<grid_for_rows(h, rows, wpb), wpb * 32, 0, s>>>(x, h->x0, rows, L.in, L.inp, ridx, rdiv, L.has_ln0 ? 1 : 0);
    else if ((L.in & 1) == 0 && ((uintptr_t)x & 7) == 0 && L.in <= 64 * 6)
        ln_noaffine_fwd_vec2_kernel<6><<<grid_for_rows(h, rows, wpb), wpb * 32, 0, s>>>(x, h->x0, rows, L.in, L.inp, ridx, rdiv, L.has_ln0 ? 1 : 0);
    else
        ln_noaffine_fwd_kernel<<<grid_for_rows(h, rows, wpb), wpb * 32, 0, s>>>(x, h->x0, rows, L.in, L.inp, ridx, rdiv, L.has_ln0 ? 1 : 0);
    if (!feat) h->launches++;
    int rc;
    // hidden blocks: block 0 reads xhat (K = padded input width, folded fc1), block k >= 1 reads h_{k-1}
    for (int k = 0; k < L.nblk; ++k) {
        const float *in = k == 0 ? (feat ? feat : h->x0) : h->hh[k - 1];
        const int K = k == 0 ? (feat ? ldf : L.in) : H, ldin = k == 0 ? (feat ? ldf : L.inp) : H;
        const float *Wk = k == 0 ? w1g : P + L.W[k], *bk = k == 0 ? b1g : P + L.b[k];
        if (h->backend == 2) {
            // tcgen05 GEMM with the block's bias + activation + LayerNorm fused into its epilogue (one kernel per block)
            const bool last_blk = k == L.nblk - 1;
            const bool fuse = head_dst && last_blk;
            // pre-split operands: the compact features and every inner h_k exist only as fp16 hi/lo (see Split16)
            const bool in_split = h->split16 && (k > 0 || feat != nullptr), out_split = h->split16 && !last_blk;
            const Split16 a16 = in_split ? (k == 0 ? feat_split(h, net) : hh_split(h, k - 1)) : Split16{nullptr, nullptr, 0};
            const Split16 h16 = out_split ? hh_split(h, k) : Split16{nullptr, nullptr, 0};
            // xhat mode: an inner block leaves only xhat_k (pre-split, unit affine; no a_k) and runs on the folded bias
            const bool xo = h->xhat && out_split, xi = h->xhat && k > 0;
            rc = tc_gemm_fwd(h, rows, ldin, in, ldin, k == 0 ? h->img_w1[net] : h->img_w[net][k], (save && !xo) ? h->a[k] : nullptr, H, s,
                             xi ? h->bf[net][k] : bk, xo ? h->ones : P + L.lg[k], xo ? h->zeros : P + L.lb[k],
                             (fuse || out_split) ? nullptr : h->hh[k], save ? h->mean[k] : nullptr,
                             save ? h->rstd[k] : nullptr,
                             (feat && k == 0) ? h->f16_fwd : fwd_f16(h, L, k),   // compact features are bounded: fp16-split eligible
                             nullptr, fuse ? h->head_fold[net] : nullptr, fuse ? L.out : 0, fuse ? head_dst : nullptr,
                             in_split ? &a16 : nullptr, out_split ? &h16 : nullptr, dzs ? h->sc + SC_RSTD + k : nullptr);
            if (rc) return rc;
            continue;
        }
        rc = launch_gemm(h, false, true, rows, H, K, in, ldin, Wk, K, h->a[k], H, false, s);
        if (rc) return rc;
        bias_relu_ln_fwd_kernel<<<grid_for_rows(h, rows, wpb), wpb * 32, 0, s>>>(h->a[k], bk, P + L.lg[k], P + L.lb[k],
                                                                                save ? h->a[k] : nullptr, h->hh[k], h->mean[k],
                                                                                h->rstd[k], rows, H, act_of(h));
        h->launches++;
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// trunk backward: dA holds dL/dh2 on entry; x0 still holds this chunk's xhat.  Accumulates into grads G; the fc1
// weight slot receives G1 = dz1^T xhat (turned into dW1 / dgamma0 / dbeta0 by ln0_finalize once per epoch).
// `dout` = gradient w.r.t. the head output ([rows, out]); on return all parameter gradients of the net have been
// accumulated into G (the fc1 slot holds G1, see ln0_finalize).
// trunk backward with PRE-SPLIT gradients (MappoHandle::dzsplit; xhat mode, head path): every dz_k is written by its LayerNorm-backward
// kernel as fp16 hi/lo into dA, scaled by the power of two of an a-priori bound (published in sc[SC_BND]); the weight-gradient and dX
// GEMMs fetch it with TMA (no producer warps), the dX GEMM leaves max |d xhat| for the next bound.  dB holds the fp32 dX outputs.
static int trunk_backward_dzs(MappoHandle *h, const NetLayout &L, const float *P, int net, float *G, const float *dout, int rows,
                              cudaStream_t s, const float *feat, int ldf) {
    const int H = L.H, wpb = 8, last = L.nblk - 1;
    const size_t lo_off = (size_t)h->chunk_rows * h->cfg.n_agents * H;
    int rc;
    const size_t nd = (size_t)rows * L.out;
    if (!h->dout_ready)
        absmax_flat_kernel<<<(unsigned)std::min<size_t>((nd + 255) / 256, (size_t)h->sm_count * 4), 256, 0, s>>>(dout, nd, h->sc + SC_DOUT);
    h->dout_ready = false;
    const int gr_head = grid_for_reduce(h, rows, wpb, 2);
    const size_t ring1 = (size_t)wpb * RP_SLOTS * H;
    const size_t sm2 = std::max(ring1, (size_t)wpb * 5 * 256) * sizeof(float), sm1 = std::max(ring1, (size_t)wpb * 4 * 256) * sizeof(float);
    if (L.out == 2)
        head_relu_ln_bwd_pipe_kernel<2, RP_SLOTS, true, true><<<gr_head, wpb * 32, sm2, s>>>(
            dout, P + L.Wh, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], P + L.lb[last], h->dA, G + L.lg[last], G + L.lb[last],
            G + L.b[last], G + L.Wh, G + L.bh, rows, H, act_of(h), nullptr, h->sc + SC_DOUT, h->sc + SC_RSTD + last, h->sc + SC_BND, lo_off);
    else
        head_relu_ln_bwd_pipe_kernel<1, RP_SLOTS, true, true><<<gr_head, wpb * 32, sm1, s>>>(
            dout, P + L.Wh, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], P + L.lb[last], h->dA, G + L.lg[last], G + L.lb[last],
            G + L.b[last], G + L.Wh, G + L.bh, rows, H, act_of(h), nullptr, h->sc + SC_DOUT, h->sc + SC_RSTD + last, h->sc + SC_BND, lo_off);
    h->launches += 2;
    DCC_CUDA_TRY(cudaGetLastError());
    const Split16 z16 = dz_split16(h, h->dA);
    float *dx = h->dB;
    static const int lnx_ctas = getenv("DCC_LNX_CTAS") ? atoi(getenv("DCC_LNX_CTAS")) : 4;
    const int gx = grid_for_reduce(h, rows, wpb, lnx_ctas);
    const size_t smx = std::max((size_t)wpb * RP_SLOTS * 2 * H, (size_t)wpb * 256) * sizeof(float);
    for (int k = last; k >= 1; --k) {
        const Split16 x16 = hh_split(h, k - 1);
        if ((rc = tc_gemm_wgrad(h, rows, H, nullptr, H, nullptr, H, G + L.W[k], H, s, true, &x16, true, &z16, h->sc + SC_BND))) return rc;
        // d xhat_{k-1} = dz_k (W_k * gamma_{k-1}); its maximum goes to sc[SC_DXH + k - 1]
        if ((rc = tc_gemm_fwd(h, rows, H, nullptr, H, h->img_wt[net][k], dx, H, s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true,
                              h->sc + SC_BND, nullptr, 0, nullptr, &z16, nullptr, nullptr, h->sc + SC_DXH + (k - 1))))
            return rc;
        const bool split_out = (k - 1 >= 1) || feat != nullptr;     // block 0 of the materialised path keeps an fp32 dz (its X operand is fp32)
        if (split_out)
            relu_lnx_bwd_pipe_kernel<true, true><<<gx, wpb * 32, smx, s>>>(
                dx, static_cast<const __half *>(x16.hi), static_cast<const __half *>(x16.lo), h->mean[k - 1], h->rstd[k - 1], h->dA,
                G + L.b[k - 1], rows, H, nullptr, h->sc + SC_DXH + (k - 1), h->sc + SC_RSTD + (k - 1), h->sc + SC_BND, lo_off);
        else
            relu_lnx_bwd_pipe_kernel<true, false><<<gx, wpb * 32, smx, s>>>(
                dx, static_cast<const __half *>(x16.hi), static_cast<const __half *>(x16.lo), h->mean[k - 1], h->rstd[k - 1], dx,
                G + L.b[k - 1], rows, H, nullptr);
        h->launches++;
        DCC_CUDA_TRY(cudaGetLastError());
    }
    if (feat) {     // compact path: Gt += dz_0^T f, both operands pre-split
        const Split16 x16 = feat_split(h, net);
        rc = tc_gemm_wgrad(h, rows, ldf, nullptr, H, nullptr, ldf, h->gt[net], ldf, s, true, &x16, true, &z16, h->sc + SC_BND);
    } else
        rc = tc_gemm_wgrad(h, rows, L.in, dx, H, h->x0, L.inp, G + L.W[0], L.in, s, wgrad_f16(h, L, 0));   // G1 += dz_0^T xhat (fp32 dz_0 in dx)
    if (rc) return rc;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// dh_last != nullptr (recurrent policies): the gradient w.r.t. the trunk OUTPUT [rows, H] is given (it comes out of the GRU's
// backward pass, rnn_backward) and the head's gradients have been accumulated already; `dout` is ignored.
static int trunk_backward(MappoHandle *h, const NetLayout &L, const float *P, int net, float *G, const float *dout, int rows,
                          cudaStream_t s, const float *feat = nullptr, int ldf = 0, const float *dh_last = nullptr) {
    const int H = L.H;
    const int wpb = 8;
    const int gr = grid_for_reduce(h, rows, wpb, h->ln_pipe ? 3 : 4);    // relu_ln_bwd_pipe: 3 CTAs per SM
    const int gr_head = grid_for_reduce(h, rows, wpb, h->ln_pipe ? 2 : 4);   // head_relu_ln_bwd_pipe: 2 CTAs per SM
    const int last = L.nblk - 1;
    // head backward + activation/LayerNorm backward of the last block in one pass: dA := dz_last
    // (f16_dx: the kernel that writes a dz also leaves max |dz| in h->dz_absmax for the fp16-split dX GEMM that reads it)
    const bool dx16 = h->f16_dx && last >= 1;
    const bool sp = h->split16 && h->backend == 2;      // weight-gradient X operands (h_{k-1}, compact features) are pre-split
    if (h->dzsplit && !dh_last && sp && h->xhat && dx16 && H == tc::TC_N) return trunk_backward_dzs(h, L, P, net, G, dout, rows, s, feat, ldf);
    uint32_t *amax = (dx16 || sp) ? h->dz_absmax : nullptr;
    if (amax) DCC_CUDA_TRY(cudaMemsetAsync(amax, 0, sizeof(uint32_t), s));
    if (dh_last) {
        if (h->ln_pipe) {
            const size_t ring = (size_t)wpb * RP_SLOTS * 2 * H;
            relu_ln_bwd_pipe_kernel<<<gr, wpb * 32, std::max(ring, (size_t)wpb * 3 * 256) * sizeof(float), s>>>(
                dh_last, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], h->dA, G + L.lg[last], G + L.lb[last], G + L.b[last],
                rows, H, act_of(h), amax);
        } else
            relu_ln_bwd_kernel<<<gr, wpb * 32, wpb * 3 * 256 * sizeof(float), s>>>(dh_last, h->a[last], h->mean[last], h->rstd[last],
                                                                                 P + L.lg[last], h->dA, G + L.lg[last], G + L.lb[last],
                                                                                 G + L.b[last], rows, H, act_of(h));
        DCC_CUDA_TRY(cudaGetLastError());
    } else if (h->ln_pipe) {
        // H == 256: the column map with 16-byte accesses (col_of<true>); DCC_LN_VEC=0 keeps the scalar map (A/B knob)
        static const bool vec_env = !(getenv("DCC_LN_VEC") && atoi(getenv("DCC_LN_VEC")) == 0);
        const bool vec = vec_env && H == 256;
        const size_t ring = (size_t)wpb * RP_SLOTS * H;
#define DCC_HEAD_ARGS dout, P + L.Wh, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], P + L.lb[last], h->dA, G + L.lg[last], \
                      G + L.lb[last], G + L.b[last], G + L.Wh, G + L.bh, rows, H, act_of(h), amax
        const size_t sm2 = std::max(ring, (size_t)wpb * 5 * 256) * sizeof(float), sm1 = std::max(ring, (size_t)wpb * 4 * 256) * sizeof(float);
        if (L.out == 2 && vec) head_relu_ln_bwd_pipe_kernel<2, RP_SLOTS, true><<<gr_head, wpb * 32, sm2, s>>>(DCC_HEAD_ARGS);
        else if (L.out == 2) head_relu_ln_bwd_pipe_kernel<2><<<gr_head, wpb * 32, sm2, s>>>(DCC_HEAD_ARGS);
        else if (vec) head_relu_ln_bwd_pipe_kernel<1, RP_SLOTS, true><<<gr_head, wpb * 32, sm1, s>>>(DCC_HEAD_ARGS);
        else head_relu_ln_bwd_pipe_kernel<1><<<gr_head, wpb * 32, sm1, s>>>(DCC_HEAD_ARGS);
#undef DCC_HEAD_ARGS
        DCC_CUDA_TRY(cudaGetLastError());
    } else if (L.out == 2)
        head_relu_ln_bwd_kernel<2><<<gr_head, wpb * 32, wpb * 5 * 256 * sizeof(float), s>>>(dout, P + L.Wh, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], P + L.lb[last],
                                                          h->dA, G + L.lg[last], G + L.lb[last], G + L.b[last], G + L.Wh, G + L.bh, rows, H, act_of(h));
    else
        head_relu_ln_bwd_kernel<1><<<gr_head, wpb * 32, wpb * 4 * 256 * sizeof(float), s>>>(dout, P + L.Wh, h->a[last], h->mean[last], h->rstd[last], P + L.lg[last], P + L.lb[last],
                                                          h->dA, G + L.lg[last], G + L.lb[last], G + L.b[last], G + L.Wh, G + L.bh, rows, H, act_of(h));
    h->launches++;
    int rc;
    float *dz = h->dA, *dx = h->dB;
    for (int k = last; k >= 1; --k) {
        if (sp) {
            const Split16 x16 = hh_split(h, k - 1);
            rc = tc_gemm_wgrad(h, rows, H, dz, H, nullptr, H, G + L.W[k], H, s, true, &x16, h->ln_pipe);   // dW_k += dz_k^T h_{k-1}
        } else
            rc = h->backend == 2 ? tc_gemm_wgrad(h, rows, H, dz, H, h->hh[k - 1], H, G + L.W[k], H, s, wgrad_f16(h, L, k))
                                 : launch_gemm(h, true, false, H, H, rows, dz, H, h->hh[k - 1], H, G + L.W[k], H, true, s);
        if (rc) return rc;
        rc = h->backend == 2 ? tc_gemm_fwd(h, rows, H, dz, H, h->img_wt[net][k], dx, H, s, nullptr, nullptr, nullptr, nullptr, nullptr,
                                           nullptr, dx16, dx16 ? h->dz_absmax : nullptr)   // dh_{k-1} = dz_k W_k
                             : launch_gemm(h, false, false, rows, H, H, dz, H, P + L.W[k], H, dx, H, false, s);
        if (rc) return rc;
        if (h->xhat && sp) {
            // xhat mode: dx holds the gradient w.r.t. xhat_{k-1} (folded weights); the mask comes from the stored xhat itself
            uint32_t *am = ((dx16 && k - 1 >= 1) || (k - 1 >= 1 || feat)) ? h->dz_absmax : nullptr;
            if (am) DCC_CUDA_TRY(cudaMemsetAsync(am, 0, sizeof(uint32_t), s));
            const Split16 xs = hh_split(h, k - 1);
            const size_t ring = (size_t)wpb * RP_SLOTS * 2 * H;
            static const bool vec_env = !(getenv("DCC_LN_VEC") && atoi(getenv("DCC_LN_VEC")) == 0);
            static const int lnx_ctas = getenv("DCC_LNX_CTAS") ? atoi(getenv("DCC_LNX_CTAS")) : 4;
            const int gx = grid_for_reduce(h, rows, wpb, lnx_ctas);
            const size_t smx = std::max(ring, (size_t)wpb * 256) * sizeof(float);
            if (vec_env && H == 256)
                relu_lnx_bwd_pipe_kernel<true><<<gx, wpb * 32, smx, s>>>(dx, static_cast<const __half *>(xs.hi), static_cast<const __half *>(xs.lo),
                                                                       h->mean[k - 1], h->rstd[k - 1], dx, G + L.b[k - 1], rows, H, am);
            else
                relu_lnx_bwd_pipe_kernel<false><<<gx, wpb * 32, smx, s>>>(dx, static_cast<const __half *>(xs.hi), static_cast<const __half *>(xs.lo),
                                                                        h->mean[k - 1], h->rstd[k - 1], dx, G + L.b[k - 1], rows, H, am);   // dx := dz_{k-1}
            DCC_CUDA_TRY(cudaGetLastError());
        } else if (h->ln_pipe) {
            // max |dz_{k-1}|: needed by the next dX GEMM (k-1 >= 1) and by the fp16-split weight-gradient GEMM of block k-1
            uint32_t *am = ((dx16 && k - 1 >= 1) || (sp && (k - 1 >= 1 || feat))) ? h->dz_absmax : nullptr;
            if (am) DCC_CUDA_TRY(cudaMemsetAsync(am, 0, sizeof(uint32_t), s));
            const size_t ring = (size_t)wpb * RP_SLOTS * 2 * H;
            relu_ln_bwd_pipe_kernel<<<gr, wpb * 32, std::max(ring, (size_t)wpb * 3 * 256) * sizeof(float), s>>>(
                dx, h->a[k - 1], h->mean[k - 1], h->rstd[k - 1], P + L.lg[k - 1], dx, G + L.lg[k - 1], G + L.lb[k - 1], G + L.b[k - 1],
                rows, H, act_of(h), am);   // dx := dz_{k-1}
            DCC_CUDA_TRY(cudaGetLastError());
        } else
            relu_ln_bwd_kernel<<<gr, wpb * 32, wpb * 3 * 256 * sizeof(float), s>>>(dx, h->a[k - 1], h->mean[k - 1], h->rstd[k - 1], P + L.lg[k - 1], dx,
                                                       G + L.lg[k - 1], G + L.lb[k - 1], G + L.b[k - 1], rows, H, act_of(h));   // dx := dz_{k-1}
        h->launches++;
        float *t = dz; dz = dx; dx = t;
    }
    if (feat && sp) {   // compact path, pre-split features: Gt += dz_0^T f  (unfolded once per optimiser step, compact_finalize)
        const Split16 x16 = feat_split(h, net);
        rc = tc_gemm_wgrad(h, rows, ldf, dz, H, nullptr, ldf, h->gt[net], ldf, s, true, &x16, h->ln_pipe);
    } else if (feat)
        rc = h->backend == 2 ? tc_gemm_wgrad(h, rows, ldf, dz, H, feat, ldf, h->gt[net], ldf, s, h->f16_wgrad)
                             : launch_gemm(h, true, false, H, ldf, rows, dz, H, feat, ldf, h->gt[net], ldf, true, s);
    else
        rc = h->backend == 2 ? tc_gemm_wgrad(h, rows, L.in, dz, H, h->x0, L.inp, G + L.W[0], L.in, s, wgrad_f16(h, L, 0))   // G1 += dz_0^T xhat
                             : launch_gemm(h, true, false, H, L.in, rows, dz, H, h->x0, L.inp, G + L.W[0], L.in, true, s);
    if (rc) return rc;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// xhat mode, once per optimiser step: the weight slot of every block k >= 1 holds G_k = dz_k^T xhat_{k-1}; turn it into
// dW_k = G_k * gamma_{k-1} + db_k (x) beta_{k-1} and produce dgamma_{k-1} = sum_h W_k * G_k, dbeta_{k-1} = W_k^T db_k
static int inner_ln_finalize(MappoHandle *h, const NetLayout &L, const float *P, float *G, cudaStream_t s) {
    if (!h->xhat) return DCC_OK;
    for (int k = 1; k < L.nblk; ++k) {
        ln0_finalize_kernel<<<(L.H + 127) / 128, 128, 0, s>>>(P + L.W[k], P + L.lg[k - 1], P + L.lb[k - 1], G + L.W[k], G + L.b[k],
                                                             G + L.lg[k - 1], G + L.lb[k - 1], L.H, L.H);
        h->launches++;
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

static int ln0_finalize(MappoHandle *h, const NetLayout &L, const float *P, float *G, cudaStream_t s) {
    int rc = inner_ln_finalize(h, L, P, G, s);
    if (rc) return rc;
    if (!L.has_ln0) return DCC_OK;   // no feature_norm: the fc1 slot already holds dW1 = dz1^T x
    ln0_finalize_kernel<<<(L.in + 127) / 128, 128, 0, s>>>(P + L.W[0], P + L.ln0_g, P + L.ln0_b, G + L.W[0], G + L.b[0], G + L.ln0_g,
                                                          G + L.ln0_b, L.H, L.in);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// compact path, once per optimiser step: G (fc1 slot) = Gt A^T, then the usual input-LayerNorm finalisation
static int compact_finalize(MappoHandle *h, const NetLayout &L, const float *P, int net, float *G, cudaStream_t s) {
    const int ldk = net ? h->cd.ldc : h->cd.lda, nb = net ? h->cd.N : 1;
    const size_t n = (size_t)L.H * L.in;
    unfold_compact_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->gt[net], h->d_poi, G + L.W[0], L.H, h->cd, nb, ldk);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    return ln0_finalize(h, L, P, G, s);
}

// ---- recurrent policies: GRU x recurrent_N + LayerNorm between the trunk and the head (dcc_rnn.cuh) --------------------------
// where the hidden state in front of step 0 comes from: stored states [*, rn, H], read at row gather[s] / gdiv (sequence
// update: the rollout row of the sequence's first step) or at row s itself (one rollout step)
struct RnnSrc {
    const float *states;
    const long long *gather;
    int gdiv;
};

static inline dim3 ew_grid(size_t n) { return dim3((unsigned)((n + 255) / 256)); }

// X [steps * S, H] (time-major trunk outputs) -> h->r_Y = LayerNorm(GRU(X)) [steps * S, H]; everything the backward pass needs
// stays in the r_* scratch.  mask_rows [steps * S]: the step's mask per row (rnn.py:26-27,66-67: state *= mask before the step).
// weight images of the GRU matrices for the tensor-core GEMMs (backend 2), once per ABI call
static int rnn_prep_images(MappoHandle *h, const NetLayout &L, const float *P, int net, bool for_backward, cudaStream_t s) {
    if (h->backend != 2) return DCC_OK;
    int rc;
    for (int l = 0; l < L.rn; ++l) {
        for (int j = 0; j < 3; ++j) {
            if ((rc = tc_prep_weights(h, P + L.Wih[l] + (size_t)j * L.H * L.H, L.H, false, L.H, h->r_img_ih[net][l][j], s, true))) return rc;
            if ((rc = tc_prep_weights(h, P + L.Whh[l] + (size_t)j * L.H * L.H, L.H, false, L.H, h->r_img_hh[net][l][j], s, true))) return rc;
        }
        if (for_backward) {
            if ((rc = tc_prep_weights(h, P + L.Wih[l], L.H, true, 3 * L.H, h->r_img_ihT[net][l], s, false))) return rc;
            if ((rc = tc_prep_weights(h, P + L.Whh[l], L.H, true, 3 * L.H, h->r_img_hhT[net][l], s, false))) return rc;
        }
    }
    return DCC_OK;
}

// C[M, 3H] = A[M, H] W^T for a GRU matrix W [3H, H]: three 256-column slices on the fp16-split tensor-core kernel, or one FFMA GEMM
static int rnn_gemm_gates(MappoHandle *h, int M, const float *A, const float *W, float *const (&img)[3], float *C, cudaStream_t s) {
    const int H = h->cfg.hidden;
    if (h->backend != 2) return launch_gemm(h, false, true, M, 3 * H, H, A, H, W, H, C, 3 * H, false, s);
    int rc;
    for (int j = 0; j < 3; ++j)
        if ((rc = tc_gemm_fwd(h, M, H, A, H, img[j], C + j * H, 3 * H, s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true))) return rc;
    return DCC_OK;
}

static int rnn_forward(MappoHandle *h, const NetLayout &L, const float *P, int net, const float *X, int S, int steps, const RnnSrc &src,
                       const float *mask_rows, cudaStream_t s) {
    const int H = L.H, H3 = 3 * L.H;
    const size_t rows = (size_t)S * steps, SH = (size_t)S * H;
    int rc;
    for (int l = 0; l < L.rn; ++l) {
        const float *Xl = l == 0 ? X : h->r_Hout[l - 1];
        // x W_ih^T for all steps at once (bias added in the gate kernel)
        if ((rc = rnn_gemm_gates(h, (int)rows, Xl, P + L.Wih[l], h->r_img_ih[net][l], h->r_GI[l], s))) return rc;
        for (int t = 0; t < steps; ++t) {
            float *Hp = h->r_Hp[l] + t * SH;
            if (t == 0)
                gru_prev_kernel<<<ew_grid(SH), 256, 0, s>>>(src.states, src.gather, src.gdiv, src.gather ? 0 : 1, L.rn, l, mask_rows, Hp, S, H);
            else
                gru_prev_kernel<<<ew_grid(SH), 256, 0, s>>>(h->r_Hout[l] + (t - 1) * SH, nullptr, 1, 0, L.rn, l, mask_rows + (size_t)t * S,
                                                           Hp, S, H);
            h->launches++;
            float *GH = h->r_GH[l] + (size_t)t * S * H3;
            if ((rc = rnn_gemm_gates(h, S, Hp, P + L.Whh[l], h->r_img_hh[net][l], GH, s))) return rc;   // the recurrence
            gru_gate_fwd_kernel<<<ew_grid(SH), 256, 0, s>>>(h->r_GI[l] + (size_t)t * S * H3, GH, P + L.bih[l], P + L.bhh[l], Hp,
                                                           h->r_gates[l] + (size_t)t * S * 4 * H, h->r_Hout[l] + t * SH, S, H);
            h->launches++;
        }
    }
    bias_relu_ln_fwd_kernel<<<grid_for_rows(h, (long)rows, 8), 256, 0, s>>>(h->r_Hout[L.rn - 1], h->r_zero, P + L.rnn_g, P + L.rnn_b, nullptr,
                                                                           h->r_Y, h->r_mean, h->r_rstd, (int)rows, H, ACT_IDENT);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// Backward of head + RNNLayer: dout [steps * S, out] -> parameter gradients of the head, the LayerNorm and every GRU layer
// accumulated into G; the gradient w.r.t. the trunk output lands in h->r_dfeat.  BPTT runs inside the pass only: the stored
// state in front of step 0 is a constant, exactly like the reference's chunked / whole-episode generators.
static int rnn_backward(MappoHandle *h, const NetLayout &L, const float *P, int net, float *G, const float *X, const float *dout, int S,
                        int steps, const float *mask_rows, cudaStream_t s) {
    const bool tcg = h->backend == 2;
    const int H = L.H, H3 = 3 * L.H, wpb = 8;
    const size_t rows = (size_t)S * steps, SH = (size_t)S * H;
    const int top = L.rn - 1;
    const int gr_head = grid_for_reduce(h, (long)rows, wpb, 2);
    int cur = 0, rc;
    // head backward fused with the backward of the RNNLayer's LayerNorm ("activation" = identity on the GRU output)
    if (L.out == 2)
        head_relu_ln_bwd_kernel<2><<<gr_head, wpb * 32, wpb * 5 * 256 * sizeof(float), s>>>(
            dout, P + L.Wh, h->r_Hout[top], h->r_mean, h->r_rstd, P + L.rnn_g, P + L.rnn_b, h->r_dH[cur], G + L.rnn_g, G + L.rnn_b,
            h->r_dummy, G + L.Wh, G + L.bh, (int)rows, H, ACT_IDENT);
    else
        head_relu_ln_bwd_kernel<1><<<gr_head, wpb * 32, wpb * 4 * 256 * sizeof(float), s>>>(
            dout, P + L.Wh, h->r_Hout[top], h->r_mean, h->r_rstd, P + L.rnn_g, P + L.rnn_b, h->r_dH[cur], G + L.rnn_g, G + L.rnn_b,
            h->r_dummy, G + L.Wh, G + L.bh, (int)rows, H, ACT_IDENT);
    h->launches++;
    DCC_CUDA_TRY(cudaGetLastError());
    for (int l = top; l >= 0; --l) {
        const float *dHl = h->r_dH[cur];
        float *dGI = h->r_GI[l], *dGH = h->r_GH[l];     // gradients overwrite the pre-activations they belong to
        for (int t = steps - 1; t >= 0; --t) {
            const bool lastt = t == steps - 1;
            gru_gate_bwd_kernel<<<ew_grid(SH), 256, 0, s>>>(dHl + t * SH, lastt ? nullptr : h->r_dhp[(t + 1) & 1],
                                                           lastt ? nullptr : mask_rows + (size_t)(t + 1) * S,
                                                           h->r_gates[l] + (size_t)t * S * 4 * H, h->r_Hp[l] + t * SH,
                                                           dGI + (size_t)t * S * H3, dGH + (size_t)t * S * H3, h->r_dhp[t & 1], S, H,
                                                           (tcg && !lastt) ? h->r_dT[(t + 1) & 1] : nullptr);
            h->launches++;
            // gradient w.r.t. the masked previous state: dhp += dGH W_hh (not needed in front of step 0: that state is data);
            // the tensor-core GEMM writes its product to r_dT, which the next (earlier) step's gate kernel adds
            if (t > 0) {
                rc = tcg ? tc_gemm_fwd(h, S, H3, dGH + (size_t)t * S * H3, H3, h->r_img_hhT[net][l], h->r_dT[t & 1], H, s)
                         : launch_gemm(h, false, false, S, H, H3, dGH + (size_t)t * S * H3, H3, P + L.Whh[l], H, h->r_dhp[t & 1], H, true, s);
                if (rc) return rc;
            }
        }
        const float *Xl = l == 0 ? X : h->r_Hout[l - 1];
        if (tcg) {      // dW += dG^T X on the 3xTF32 weight-gradient kernel, one 256-row gate slice at a time
            for (int j = 0; j < 3; ++j) {
                if ((rc = tc_gemm_wgrad(h, (int)rows, H, dGH + j * H, H3, h->r_Hp[l], H, G + L.Whh[l] + (size_t)j * H * H, H, s))) return rc;
                if ((rc = tc_gemm_wgrad(h, (int)rows, H, dGI + j * H, H3, Xl, H, G + L.Wih[l] + (size_t)j * H * H, H, s))) return rc;
            }
        } else {
            if ((rc = launch_gemm(h, true, false, H3, H, (int)rows, dGH, H3, h->r_Hp[l], H, G + L.Whh[l], H, true, s))) return rc;
            if ((rc = launch_gemm(h, true, false, H3, H, (int)rows, dGI, H3, Xl, H, G + L.Wih[l], H, true, s))) return rc;
        }
        const dim3 cg((H3 + 127) / 128, (unsigned)std::min<size_t>((rows + 255) / 256, 256));
        colsum_atomic_kernel<<<cg, 128, 0, s>>>(dGI, rows, H3, G + L.bih[l]);
        colsum_atomic_kernel<<<cg, 128, 0, s>>>(dGH, rows, H3, G + L.bhh[l]);
        h->launches += 2;
        float *dX = l > 0 ? h->r_dH[cur ^ 1] : h->r_dfeat;
        rc = tcg ? tc_gemm_fwd(h, (int)rows, H3, dGI, H3, h->r_img_ihT[net][l], dX, H, s)
                 : launch_gemm(h, false, false, (int)rows, H, H3, dGI, H3, P + L.Wih[l], H, dX, H, false, s);
        if (rc) return rc;
        cur ^= 1;
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

// sequences per pass of the sequence update: whole sequences of `steps` rows within one chunk of the (critic's) scratch
static inline int rnn_pass_seqs(const MappoHandle *h, int steps) { return steps > 0 ? h->chunk_rows / steps : 0; }

__global__ void expand_values_kernel(const float *__restrict__ v, float *__restrict__ out, int rows, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * N) out[i] = v[i / N];
}
__global__ void add_scalar_kernel(float *p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += v;
}
__global__ void entropy_stat_kernel(const float *logstd, int A, double *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double e = 0.0;
        for (int d = 0; d < A; ++d) e += 0.5 + 0.5 * (double)LOG_2PI + (double)logstd[d];
        *out += e;
    }
}

}  // namespace dcc

using namespace dcc;

extern "C" {

int dcc_mappo_cfg_default(dcc_mappo_cfg *c) {
    if (!c) return DCC_ERR_INVALID_ARG;
    memset(c, 0, sizeof *c);
    c->n_agents = 4; c->obs_dim = 110; c->hidden = 256; c->act_dim = 2; c->chunk_rows = 0; c->gemm_backend = 0;
    c->clip_param = 0.2f; c->entropy_coef = 0.01f; c->value_loss_coef = 1.0f; c->huber_delta = 10.0f;
    c->max_grad_norm = 10.0f; c->gamma = 0.99f; c->gae_lambda = 0.95f; c->opti_eps = 1e-5f; c->vn_beta = 0.99999;
    c->adam_beta1 = 0.9f; c->adam_beta2 = 0.999f;
    c->use_huber_loss = 1; c->use_clipped_value_loss = 1; c->use_max_grad_norm = 1; c->use_valuenorm = 1; c->use_gae = 1;
    c->weight_decay = 0.f; c->use_feature_normalization = 1; c->use_relu = 1; c->layer_N = 1;
    return DCC_OK;
}

int dcc_mappo_create(const dcc_mappo_cfg *cfg, int device, void **handle) {
    if (!cfg || !handle) return DCC_ERR_INVALID_ARG;
    *handle = nullptr;
    if (cfg->n_agents < 1 || cfg->obs_dim < 1 || cfg->hidden < 1 || cfg->chunk_rows < 0) return DCC_ERR_INVALID_ARG;
    if (cfg->hidden > 256 || cfg->act_dim != 2 || cfg->gemm_backend < 0 || cfg->gemm_backend > 2 || cfg->layer_N < 1 ||
        cfg->layer_N > MAX_BLOCKS - 1)
        return DCC_ERR_UNSUPPORTED;  // MLP trunk with hidden <= 256, 1..3 fc2 blocks and the Box(2) action space of the env
    if (cfg->gemm_backend == 2 && !tc_supported(cfg)) return DCC_ERR_UNSUPPORTED;
    if (cfg->recurrent_N < 0 || cfg->recurrent_N > MAX_RNN) return DCC_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return DCC_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(device);
    cudaDeviceProp prop;
    DCC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return DCC_ERR_NO_DEVICE;
    if (cfg->gemm_backend == 2 || (cfg->gemm_backend == 0 && tc_supported(cfg))) {
        // the opt-in above 48 KB of dynamic shared memory is a per-device function attribute: set it for THIS handle's
        // device at creation (idempotent; no process-wide "done" flag, which would skip a second GPU)
        int rc = tc_set_kernel_attributes();
        if (rc) return rc;
    }
    {
        int rc = pipe_set_kernel_attributes();
        if (rc) return rc;
    }
    MappoHandle *h = new (std::nothrow) MappoHandle();
    if (!h) return DCC_ERR_ALLOC;
    memset(h, 0, sizeof *h);
    h->magic = MAPPO_MAGIC; h->cfg = *cfg; h->device = device; h->sm_count = prop.multiProcessorCount;
    h->backend = cfg->gemm_backend ? cfg->gemm_backend : (tc_supported(cfg) ? 2 : 1);
    h->f16_fwd = !(getenv("DCC_TC_F16") && atoi(getenv("DCC_TC_F16")) == 0);
    h->f16_wgrad = getenv("DCC_TC_WGRAD_F16") && atoi(getenv("DCC_TC_WGRAD_F16")) == 1;
    h->ln_pipe = (cfg->hidden % 4 == 0) && !(getenv("DCC_LN_PIPE") && atoi(getenv("DCC_LN_PIPE")) == 0);
    h->f16_dx = h->backend == 2 && h->ln_pipe && !(getenv("DCC_TC_DX_F16") && atoi(getenv("DCC_TC_DX_F16")) == 0);
    h->split16 = h->backend == 2 && h->f16_fwd && !(getenv("DCC_TC_SPLIT") && atoi(getenv("DCC_TC_SPLIT")) == 0) &&
                 !(getenv("DCC_TC_TMA") && atoi(getenv("DCC_TC_TMA")) == 0) && tc::tc_tensor_map_encoder() != nullptr;
    // xhat mode needs the pre-split operand path, the row-pipeline kernels and ReLU (the mask is rebuilt from xhat)
    h->xhat = h->split16 && h->ln_pipe && cfg->use_relu != 0 && !(getenv("DCC_TC_XHAT") && atoi(getenv("DCC_TC_XHAT")) == 0);
    h->dzsplit = h->xhat && h->f16_dx && !(getenv("DCC_TC_DZSPLIT") && atoi(getenv("DCC_TC_DZSPLIT")) == 0);
    const int N = cfg->n_agents, D = cfg->obs_dim, H = cfg->hidden;
    h->la.init(D, H, cfg->act_dim, true, cfg->use_feature_normalization != 0, cfg->layer_N, cfg->recurrent_N);
    h->lc.init(N * D, H, 1, false, cfg->use_feature_normalization != 0, cfg->layer_N, cfg->recurrent_N);
    // chunk: bound the activation scratch to ~7.5 GB unless the caller asks for a size.  Larger chunks mean fewer, longer launches:
    // 113 664 env steps per chunk (6 waves of critic tiles, 48 of actor tiles per kernel) instead of 37 888 took 4.9 % off the update at
    // 8/64/65 536 (ramp-up / drain of ~20 kernels per chunk); beyond that the gain saturates (profiles/r02end_chunk_size_sweep.txt)
    long chunk = cfg->chunk_rows;
    if (chunk <= 0) {
        const double per_row = (double)N * (D + (2.0 + 2.0 * (1 + cfg->layer_N)) * H + 16) * 4.0;
        chunk = (long)(7.5e9 / per_row);
        if (chunk > 131072) chunk = 131072;
        // recurrent policies keep ~13 H floats per row and GRU layer on top: bound the pass to 65 536 agent rows
        if (cfg->recurrent_N > 0 && chunk * N > 65536) chunk = 65536 / N;
        // whole waves of 128-row tiles on the persistent grid (one CTA per SM).  The critic runs one row per env step,
        // so its tile count is chunk / 128: with chunk a multiple of 128 * SMs BOTH nets fill every wave (a chunk
        // aligned for the actor only leaves the critic's forward-shaped GEMMs with a 25 %-full second wave).
        const long wave = (long)prop.multiProcessorCount * 128;
        if (chunk >= wave) chunk = chunk / wave * wave;
        else if (chunk * N >= wave) chunk = (chunk * N / wave) * wave / N;
        if (chunk < 64) chunk = 64;
    }
    h->chunk_rows = (int)chunk;
    const size_t RA = (size_t)chunk * N;
    cudaError_t ce = cudaSuccess;
    auto alloc = [&](float **p, size_t n) { if (ce == cudaSuccess) ce = cudaMalloc(p, n * sizeof(float)); };
    {   // xhat scratch: [chunk*N, pad32(D)] for the actor or [chunk, pad32(N*D)] for the critic, whichever is larger
        const size_t xa = RA * (size_t)h->la.inp, xc = (size_t)chunk * h->lc.inp;
        alloc(&h->x0, xa > xc ? xa : xc);
    }
    if (h->backend == 2) {
        const NetLayout *Ls[2] = {&h->la, &h->lc};
        for (int n = 0; n < 2; ++n) {
            const size_t kt1 = Ls[n]->inp / tc::TC_BK, kt2 = (H + tc::TC_BK - 1) / tc::TC_BK;
            alloc(&h->img_w1[n], kt1 * 2 * tc::TC_B_TILE_FLOATS);
            for (int k = 1; k < h->la.nblk; ++k) {
                alloc(&h->img_w[n][k], kt2 * 2 * tc::TC_B_TILE_FLOATS);
                alloc(&h->img_wt[n][k], kt2 * 2 * tc::TC_B_TILE_FLOATS);
            }
        }
    }
    for (int k = 0; k < h->la.nblk; ++k) {
        alloc(&h->a[k], RA * H); alloc(&h->hh[k], RA * H); alloc(&h->mean[k], RA); alloc(&h->rstd[k], RA);
    }
    alloc(&h->dA, RA * H); alloc(&h->dB, RA * H);
    alloc(&h->w1g_a, (size_t)H * D); alloc(&h->b1g_a, H); alloc(&h->w1g_c, (size_t)H * N * D); alloc(&h->b1g_c, H);
    alloc(&h->mu, RA * 2); alloc(&h->logp, RA); alloc(&h->dmu, RA * 2); alloc(&h->vnew, RA); alloc(&h->dv, RA);   // critic: critic_cap_rows
    alloc(&h->vn_gae, 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&h->dsums, 8 * sizeof(double));
    if (ce == cudaSuccess && h->backend == 2) ce = cudaMalloc(&h->dz_absmax, sizeof(uint32_t));
    for (int n = 0; n < 2 && h->backend == 2; ++n) alloc(&h->head_fold[n], 2 * 256 + 8);
    if (h->dzsplit && ce == cudaSuccess) {
        ce = cudaMalloc(&h->sc, SC_COUNT * sizeof(uint32_t));
        if (ce == cudaSuccess) ce = cudaMemset(h->sc, 0, SC_COUNT * sizeof(uint32_t));
    }
    if (h->xhat) {
        for (int n = 0; n < 2; ++n)
            for (int k = 1; k < h->la.nblk; ++k) { alloc(&h->wf[n][k], (size_t)H * H); alloc(&h->bf[n][k], H); }
        alloc(&h->ones, H); alloc(&h->zeros, H);
        if (ce == cudaSuccess) ce = cudaMemset(h->zeros, 0, H * sizeof(float));
        if (ce == cudaSuccess) {
            float one[256];
            for (int i = 0; i < 256; ++i) one[i] = 1.f;
            ce = cudaMemcpy(h->ones, one, H * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    if (cfg->recurrent_N > 0) {
        for (int l = 0; l < cfg->recurrent_N; ++l) {
            alloc(&h->r_Hp[l], RA * H); alloc(&h->r_GI[l], RA * 3 * H); alloc(&h->r_GH[l], RA * 3 * H);
            alloc(&h->r_gates[l], RA * 4 * H); alloc(&h->r_Hout[l], RA * H);
        }
        for (int i = 0; i < 2; ++i) { alloc(&h->r_dH[i], RA * H); alloc(&h->r_dhp[i], RA * H); }
        if (h->backend == 2) {
            const size_t kt16 = (H + tc::TC_BK16 - 1) / tc::TC_BK16, kt32 = (3 * H + tc::TC_BK - 1) / tc::TC_BK;
            for (int n = 0; n < 2; ++n)
                for (int l = 0; l < cfg->recurrent_N; ++l) {
                    for (int j = 0; j < 3; ++j) {
                        alloc(&h->r_img_ih[n][l][j], kt16 * 2 * tc::TC_B_TILE_FLOATS);
                        alloc(&h->r_img_hh[n][l][j], kt16 * 2 * tc::TC_B_TILE_FLOATS);
                    }
                    alloc(&h->r_img_ihT[n][l], kt32 * 2 * tc::TC_B_TILE_FLOATS);
                    alloc(&h->r_img_hhT[n][l], kt32 * 2 * tc::TC_B_TILE_FLOATS);
                }
            for (int i = 0; i < 2; ++i) alloc(&h->r_dT[i], RA * H);
        }
        alloc(&h->r_Y, RA * H); alloc(&h->r_mean, RA); alloc(&h->r_rstd, RA); alloc(&h->r_mask, RA); alloc(&h->r_dfeat, RA * H);
        alloc(&h->r_zero, 3 * H); alloc(&h->r_dummy, H);
        if (ce == cudaSuccess) ce = cudaMemset(h->r_zero, 0, 3 * H * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMemset(h->r_dummy, 0, H * sizeof(float));
    }
    if (ce != cudaSuccess) {
        set_last_cuda_error(ce, "cudaMalloc(mappo scratch)", __FILE__, __LINE__);
        dcc_mappo_destroy(h);
        return DCC_ERR_ALLOC;
    }
    *handle = h;
    return DCC_OK;
}

int dcc_mappo_destroy(void *handle) {
    MappoHandle *h = as_mappo(handle);
    if (!h) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    float *bufs[] = {h->x0, h->dA, h->dB, h->w1g_a, h->b1g_a, h->w1g_c, h->b1g_c, h->mu, h->logp, h->dmu, h->vnew, h->dv,
                     h->vn_gae, h->img_w1[0], h->img_w1[1]};
    for (float *b : bufs) cudaFree(b);
    for (int k = 0; k < MAX_BLOCKS; ++k) {
        float *blk[] = {h->a[k], h->hh[k], h->mean[k], h->rstd[k], h->img_w[0][k], h->img_w[1][k], h->img_wt[0][k], h->img_wt[1][k]};
        for (float *b : blk) cudaFree(b);
    }
    for (int n = 0; n < 2; ++n)
        for (int k = 0; k < MAX_BLOCKS; ++k) { cudaFree(h->wf[n][k]); cudaFree(h->bf[n][k]); }
    cudaFree(h->ones); cudaFree(h->zeros); cudaFree(h->sc);
    for (int l = 0; l < MAX_RNN; ++l) {
        float *rb[] = {h->r_Hp[l], h->r_GI[l], h->r_GH[l], h->r_gates[l], h->r_Hout[l]};
        for (float *b : rb) cudaFree(b);
    }
    for (int n = 0; n < 2; ++n)
        for (int l = 0; l < MAX_RNN; ++l) {
            for (int j = 0; j < 3; ++j) { cudaFree(h->r_img_ih[n][l][j]); cudaFree(h->r_img_hh[n][l][j]); }
            cudaFree(h->r_img_ihT[n][l]); cudaFree(h->r_img_hhT[n][l]);
        }
    float *rbufs[] = {h->r_dH[0], h->r_dH[1], h->r_dhp[0], h->r_dhp[1], h->r_Y, h->r_mean, h->r_rstd, h->r_mask, h->r_dfeat,
                      h->r_zero, h->r_dummy, h->r_dT[0], h->r_dT[1]};
    for (float *b : rbufs) cudaFree(b);
    cudaFree(h->dsums);
    cudaFree(h->dz_absmax);
    cudaFree(h->d_poi); cudaFree(h->fc);
    for (int n = 0; n < 2; ++n) { cudaFree(h->wt[n]); cudaFree(h->gt[n]); cudaFree(h->head_fold[n]); }
    h->magic = 0;
    delete h;
    return DCC_OK;
}

int64_t dcc_mappo_param_count(void *handle, int which) {
    MappoHandle *h = as_mappo(handle);
    if (!h || (which != 0 && which != 1)) return -1;
    return (int64_t)(which == 0 ? h->la.total : h->lc.total);
}

int dcc_mappo_chunk_rows(void *handle) {
    MappoHandle *h = as_mappo(handle);
    return h ? h->chunk_rows : -1;
}

int dcc_mappo_gemm_backend(void *handle) {
    MappoHandle *h = as_mappo(handle);
    return h ? h->backend : -1;
}

int64_t dcc_mappo_launch_count(void *handle) {
    MappoHandle *h = as_mappo(handle);
    return h ? h->launches : -1;
}

// shared body of get_actions / evaluate_actions.  mode 0 = sample, 1 = evaluate given actions.
// d_obs == nullptr: compact path, the rows come from the env state (pv [n_envs, N, 4] float64, en [n_envs, M] uint8).
static int policy_forward(MappoHandle *h, const float *actor, const float *critic, const float *d_obs, int n_envs, int mode,
                          uint64_t seed, uint64_t offset, int deterministic, float *d_actions, float *d_logp,
                          float *d_values, float *d_mu, cudaStream_t s, const double *pv = nullptr, const uint8_t *en = nullptr) {
    const int N = h->cfg.n_agents, D = h->cfg.obs_dim, H = h->cfg.hidden;
    const bool do_actor = actor && d_actions && (mode == 0 || d_logp);
    const bool do_critic = critic && d_values;
    const bool cmp = d_obs == nullptr;
    int rc;
    if (do_actor && (rc = cmp ? fold_compact(h, h->la, actor, 0, false, s) : fold_ln0(h, h->la, actor, 0, false, s))) return rc;
    if (do_critic && (rc = cmp ? fold_compact(h, h->lc, critic, 1, false, s) : fold_ln0(h, h->lc, critic, 1, false, s))) return rc;
    for (int e0 = 0; e0 < n_envs; e0 += h->chunk_rows) {
        const int ne = min(h->chunk_rows, n_envs - e0);
        const float *x = cmp ? nullptr : d_obs + (size_t)e0 * N * D;
        if (cmp && (rc = compact_features(h, pv + (size_t)e0 * N * 4, en + (size_t)e0 * h->cd.M, ne, do_actor, do_critic, s))) return rc;
        const float *fa = cmp ? h->x0 : nullptr, *fcr = cmp ? h->fc : nullptr;
        if (do_actor) {
            const int rows = ne * N;
            const bool fh = fused_head(h);
            if ((rc = trunk_forward(h, h->la, actor, 0, x, rows, false, s, nullptr, 1, fa, h->cd.lda, fh ? h->mu : nullptr))) return rc;
            if (fh)
                gauss_finish_kernel<<<(rows + 255) / 256, 256, 0, s>>>(
                    h->mu, actor + h->la.logstd, d_actions + (size_t)e0 * N * 2, d_mu ? d_mu + (size_t)e0 * N * 2 : nullptr,
                    d_logp ? d_logp + (size_t)e0 * N : nullptr, rows, mode, deterministic, seed, offset, (uint64_t)e0 * N);
            else
                actor_head_kernel<<<grid_for_rows(h, rows, 8), 256, 0, s>>>(
                    h->hh[h->la.nblk - 1], actor + h->la.Wh, actor + h->la.bh, actor + h->la.logstd, d_actions + (size_t)e0 * N * 2,
                    d_mu ? d_mu + (size_t)e0 * N * 2 : nullptr, d_logp ? d_logp + (size_t)e0 * N : nullptr, rows, H, mode,
                    deterministic, seed, offset, (uint64_t)e0 * N);
            h->launches++;
        }
        if (do_critic) {
            const bool fh = fused_head(h);
            if ((rc = trunk_forward(h, h->lc, critic, 1, x, ne, false, s, nullptr, 1, fcr, h->cd.ldc, fh ? d_values + e0 : nullptr))) return rc;
            if (!fh) {
                critic_head_kernel<<<grid_for_rows(h, ne, 8), 256, 0, s>>>(h->hh[h->la.nblk - 1], critic + h->lc.Wh, critic + h->lc.bh,
                                                                          d_values + e0, ne, H);
                h->launches++;
            }
        }
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_act(void *handle, const float *actor, const float *critic, const float *d_obs, int n_envs, uint64_t seed,
                  uint64_t offset, int deterministic, float *d_actions, float *d_logp, float *d_values,
                  dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_obs || n_envs < 1 || (!actor && !critic)) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N > 0) return DCC_ERR_UNSUPPORTED;      // recurrent handles: dcc_mappo_act_rnn
    if ((actor && !d_actions) || (critic && !d_values)) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    return policy_forward(h, actor, critic, d_obs, n_envs, 0, seed, offset, deterministic, d_actions, d_logp, d_values,
                          nullptr, static_cast<cudaStream_t>(stream));
}

int dcc_mappo_evaluate(void *handle, const float *actor, const float *critic, const float *d_obs, const float *d_actions,
                       int n_envs, float *d_logp, float *d_values, float *d_mu, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_obs || n_envs < 1 || (!actor && !critic)) return DCC_ERR_INVALID_ARG;
    if ((actor && (!d_actions || !d_logp)) || (critic && !d_values)) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N > 0) return DCC_ERR_UNSUPPORTED;      // recurrent handles: dcc_mappo_act_rnn (mode 1)
    DCC_DEVICE_GUARD(h->device);
    return policy_forward(h, actor, critic, d_obs, n_envs, 1, 0, 0, 0, const_cast<float *>(d_actions), d_logp, d_values,
                          d_mu, static_cast<cudaStream_t>(stream));
}

int dcc_mappo_set_env_layout(void *handle, int n_pois, const double *h_poi_xy, double m_energy) {
    MappoHandle *h = as_mappo(handle);
    if (!h || n_pois < 1 || !h_poi_xy || !(m_energy > 0)) return DCC_ERR_INVALID_ARG;
    const int N = h->cfg.n_agents;
    // the observation layout must be the env's (coverage.py:99-110) and the critic centralised over the N rows
    if (h->cfg.obs_dim != 4 + 2 * (N - 1) + 5 * n_pois || h->lc.in != N * h->cfg.obs_dim) return DCC_ERR_UNSUPPORTED;
    if (h->cfg.recurrent_N > 0) return DCC_ERR_UNSUPPORTED;      // recurrent policies run on materialised observation rows
    DCC_DEVICE_GUARD(h->device);
    CompactDims cd;
    cd.init(N, n_pois, m_energy);
    cd.set_poi(h_poi_xy);
    if ((size_t)cd.lda > (size_t)h->la.inp) return DCC_ERR_UNSUPPORTED;   // actor features live in the xhat scratch
    if (compact_features_smem(cd, 8) > 200 * 1024) return DCC_ERR_UNSUPPORTED;
    cudaFree(h->d_poi); cudaFree(h->fc);
    for (int n = 0; n < 2; ++n) { cudaFree(h->wt[n]); cudaFree(h->gt[n]); h->wt[n] = h->gt[n] = nullptr; }
    h->d_poi = nullptr; h->fc = nullptr; h->compact = false;
    const int H = h->cfg.hidden;
    cudaError_t ce = cudaMalloc(&h->d_poi, sizeof(double) * 2 * n_pois);
    if (ce == cudaSuccess) ce = cudaMalloc(&h->fc, critic_cap_rows(h) * cd.ldc * sizeof(float));
    const int lds[2] = {cd.lda, cd.ldc};
    for (int n = 0; n < 2 && ce == cudaSuccess; ++n) {
        ce = cudaMalloc(&h->wt[n], (size_t)H * lds[n] * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc(&h->gt[n], (size_t)H * lds[n] * sizeof(float));
    }
    if (ce != cudaSuccess) { set_last_cuda_error(ce, "cudaMalloc(compact scratch)", __FILE__, __LINE__); return DCC_ERR_ALLOC; }
    DCC_CUDA_TRY(cudaMemcpy(h->d_poi, h_poi_xy, sizeof(double) * 2 * n_pois, cudaMemcpyHostToDevice));
    const size_t smem = compact_features_smem(cd, 8);
    if (smem > 48 * 1024) {
        // the opt-in is a per-device attribute of the kernel: only ever RAISE it, so that a later handle with a smaller layout
        // cannot pull it under an earlier live handle's footprint
        static std::mutex mu;
        static int max_smem[64] = {0};
        std::lock_guard<std::mutex> lk(mu);
        int &mx = max_smem[h->device & 63];
        if ((int)smem > mx) {
            DCC_CUDA_TRY(cudaFuncSetAttribute(compact_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mx = (int)smem;
        }
    }
    h->cd = cd;
    h->compact = true;
    return DCC_OK;
}

int dcc_mappo_act_state(void *handle, const float *actor, const float *critic, const double *d_pos_vel, const uint8_t *d_energy,
                        int n_envs, uint64_t seed, uint64_t offset, int deterministic, float *d_actions, float *d_logp,
                        float *d_values, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_pos_vel || !d_energy || n_envs < 1 || (!actor && !critic)) return DCC_ERR_INVALID_ARG;
    if ((actor && !d_actions) || (critic && !d_values)) return DCC_ERR_INVALID_ARG;
    if (!h->compact) return DCC_ERR_UNSUPPORTED;
    DCC_DEVICE_GUARD(h->device);
    return policy_forward(h, actor, critic, nullptr, n_envs, 0, seed, offset, deterministic, d_actions, d_logp, d_values,
                          nullptr, static_cast<cudaStream_t>(stream), d_pos_vel, d_energy);
}

int dcc_mappo_evaluate_state(void *handle, const float *actor, const float *critic, const double *d_pos_vel,
                             const uint8_t *d_energy, const float *d_actions, int n_envs, float *d_logp, float *d_values,
                             float *d_mu, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_pos_vel || !d_energy || n_envs < 1 || (!actor && !critic)) return DCC_ERR_INVALID_ARG;
    if ((actor && (!d_actions || !d_logp)) || (critic && !d_values)) return DCC_ERR_INVALID_ARG;
    if (!h->compact) return DCC_ERR_UNSUPPORTED;
    DCC_DEVICE_GUARD(h->device);
    return policy_forward(h, actor, critic, nullptr, n_envs, 1, 0, 0, 0, const_cast<float *>(d_actions), d_logp, d_values,
                          d_mu, static_cast<cudaStream_t>(stream), d_pos_vel, d_energy);
}

int dcc_obs_from_state(void *handle, const double *d_pos_vel, const uint8_t *d_energy, int n_rows, float *d_obs,
                       dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_pos_vel || !d_energy || !d_obs || n_rows < 1) return DCC_ERR_INVALID_ARG;
    if (!h->compact) return DCC_ERR_UNSUPPORTED;
    DCC_DEVICE_GUARD(h->device);
    obs_from_state_kernel<<<grid_for_rows(h, n_rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_pos_vel, d_energy, h->d_poi,
                                                                                                  d_obs, n_rows, h->cd);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

int dcc_rollout_insert(const float *d_rew_in, const uint8_t *d_done_in, int n_envs, int n_agents, float *d_rew_out,
                       float *d_mask_out, dcc_stream_t stream) {
    if (!d_rew_in || !d_done_in || !d_rew_out || !d_mask_out || n_envs < 1 || n_agents < 1) return DCC_ERR_INVALID_ARG;
    rollout_insert_kernel<<<(n_envs + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_rew_in, d_done_in, n_envs,
                                                                                            n_agents, d_rew_out, d_mask_out);
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_gae(void *handle, const float *d_rewards, const float *d_values, const float *d_masks,
                  const float *d_vn_state, int T, int E, float *d_returns, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_rewards || !d_values || !d_masks || !d_returns || T < 1 || E < 1) return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && h->cfg.use_gae && !d_vn_state) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    gae_kernel<<<(E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_rewards, d_values, d_masks, h->cfg.use_valuenorm ? d_vn_state : nullptr, d_returns, T, E, h->cfg.gamma,
        h->cfg.gae_lambda, h->cfg.use_gae);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

int dcc_mappo_train_begin(void *handle, const float *d_returns, const float *d_values, const float *d_vn_state, int T, int E,
                          double *d_stats_out, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_returns || !d_values || !d_stats_out || T < 1 || E < 1) return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)T * E;
    DCC_CUDA_TRY(cudaMemsetAsync(d_stats_out, 0, 4 * sizeof(double), s));
    if (h->cfg.use_valuenorm)
        DCC_CUDA_TRY(cudaMemcpyAsync(h->vn_gae, d_vn_state, 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8);
    sum_sumsq_kernel<<<blocks, 256, 0, s>>>(d_returns, d_values, vn_snapshot(h), d_stats_out, n);
    sum_sumsq_kernel<<<blocks, 256, 0, s>>>(d_returns, nullptr, nullptr, d_stats_out + 2, n);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches += 2;
    return DCC_OK;
}

static PpoLossParams loss_params(const MappoHandle *h, double agent_rows_global) {
    PpoLossParams P;
    P.clip = h->cfg.clip_param; P.huber_delta = h->cfg.huber_delta; P.value_coef = h->cfg.value_loss_coef;
    P.inv_rows = (float)(1.0 / agent_rows_global); P.n_agents = h->cfg.n_agents;
    P.use_huber = h->cfg.use_huber_loss; P.use_clipped = h->cfg.use_clipped_value_loss;
    return P;
}

// shared prologue of an optimiser step: zero the gradients, ValueNorm.update with the (mini)batch return statistics
// (cal_value_loss, mappo.py:106-107), entropy statistic, fold the input LayerNorm into fc1 / refresh weight images
static int grads_prologue(MappoHandle *h, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                          float *d_vn_state, const double *ret_sums, double n_ret, double *d_epoch_stats, cudaStream_t s,
                          bool cmp = false) {
    DCC_CUDA_TRY(cudaMemsetAsync(grad_actor, 0, h->la.total * sizeof(float), s));
    DCC_CUDA_TRY(cudaMemsetAsync(grad_critic, 0, h->lc.total * sizeof(float), s));
    if (h->cfg.use_valuenorm) {
        vn_update_kernel<<<1, 32, 0, s>>>(d_vn_state, ret_sums, n_ret, (float)h->cfg.vn_beta, (float)(1.0 - h->cfg.vn_beta));
        h->launches++;
    }
    entropy_stat_kernel<<<1, 32, 0, s>>>(actor + h->la.logstd, h->cfg.act_dim, d_epoch_stats + 3);
    h->launches++;
    int rc;
    if (cmp) {
        DCC_CUDA_TRY(cudaMemsetAsync(h->gt[0], 0, (size_t)h->la.H * h->cd.lda * sizeof(float), s));
        DCC_CUDA_TRY(cudaMemsetAsync(h->gt[1], 0, (size_t)h->lc.H * h->cd.ldc * sizeof(float), s));
        if ((rc = fold_compact(h, h->la, actor, 0, true, s))) return rc;
        if ((rc = fold_compact(h, h->lc, critic, 1, true, s))) return rc;
        return DCC_OK;
    }
    if ((rc = fold_ln0(h, h->la, actor, 0, true, s))) return rc;
    if ((rc = fold_ln0(h, h->lc, critic, 1, true, s))) return rc;
    return DCC_OK;
}

static int epoch_grads_impl(MappoHandle *h, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                            const float *d_obs, const double *d_pv, const uint8_t *d_en, const float *d_actions,
                            const float *d_logp_old, const float *d_values, const float *d_returns, float *d_vn_state,
                            const double *d_stats4, double n_rows_global, int T, int E, double *d_epoch_stats, cudaStream_t s) {
    const int N = h->cfg.n_agents, D = h->cfg.obs_dim, H = h->cfg.hidden;
    const NetLayout &LA = h->la, &LC = h->lc;
    const bool cmp = d_obs == nullptr;
    // ValueNorm.update(return_batch) happens inside cal_value_loss, every epoch, before normalising (mappo.py:107)
    int rc;
    if ((rc = grads_prologue(h, actor, critic, grad_actor, grad_critic, d_vn_state, d_stats4 + 2, n_rows_global, d_epoch_stats, s, cmp)))
        return rc;
    const float *vn_now = h->cfg.use_valuenorm ? d_vn_state : nullptr;
    const PpoLossParams P = loss_params(h, n_rows_global * N);
    const long R = (long)T * E;
    const float *fa = cmp ? h->x0 : nullptr, *fcr = cmp ? h->fc : nullptr;
    // Compact path: the critic (one row per env step) is run once per SUPER-chunk of N actor chunks (critic_cap_rows): the
    // feature kernel of every actor chunk appends that chunk's critic rows to `fc`, and the critic's forward / loss / backward
    // then see chunk * N rows = as many row tiles as the actor's kernels.  DCC_CRITIC_SUPER=0 (A/B knob) and the
    // materialised path (whose xhat scratch holds one chunk of critic rows) keep one critic pass per actor chunk.
    static const bool super_env = !(getenv("DCC_CRITIC_SUPER") && atoi(getenv("DCC_CRITIC_SUPER")) == 0);
    const long SR = (cmp && super_env) ? (long)critic_cap_rows(h) : (long)h->chunk_rows;
    const bool fh = fused_head(h);
    for (long s0 = 0; s0 < R; s0 += SR) {
        const long ns = std::min<long>(SR, R - s0);
        for (long r0 = s0; r0 < s0 + ns; r0 += h->chunk_rows) {
            const int nr = (int)std::min<long>(h->chunk_rows, s0 + ns - r0);
            const float *x = cmp ? nullptr : d_obs + (size_t)r0 * N * D;
            // compact path: one pass over the chunk's 32 N + M bytes of state per row yields both nets' layer-1 operands
            if (cmp && (rc = compact_features(h, d_pv + (size_t)r0 * N * 4, d_en + (size_t)r0 * h->cd.M, nr, true, true, s, nullptr,
                                              (size_t)(r0 - s0))))
                return rc;
            // actor and critic share one activation scratch, so the nets are processed one after the other:
            // 1) actor: forward (activations saved) -> Gaussian head -> policy loss -> backward
            if ((rc = trunk_forward(h, LA, actor, 0, x, nr * N, true, s, nullptr, 1, fa, h->cd.lda, fh ? h->mu : nullptr))) return rc;
            // fused head: mu is there already; the loss kernel forms the new log-prob itself (and, for the pre-split gradients, leaves
            // max |dmu| behind) — no separate log-prob pass
            if (!fh) {
                actor_head_kernel<<<grid_for_rows(h, nr * N, 8), 256, 0, s>>>(
                    h->hh[h->la.nblk - 1], actor + LA.Wh, actor + LA.bh, actor + LA.logstd, const_cast<float *>(d_actions) + (size_t)r0 * N * 2,
                    h->mu, h->logp, nr * N, H, 1, 0, 0, 0, 0);
                h->launches++;
            }
            const bool dmax = h->dzsplit && h->sc;
            ppo_policy_loss_kernel<<<(nr + 127) / 128, 128, 0, s>>>(
                h->mu, fh ? nullptr : h->logp, d_actions + (size_t)r0 * N * 2, actor + LA.logstd, d_logp_old + (size_t)r0 * N, d_returns + r0,
                d_values + r0, vn_snapshot(h), d_stats4, n_rows_global, h->dmu, grad_actor + LA.logstd, d_epoch_stats, nr, P,
                dmax ? h->sc + SC_DOUT : nullptr);
            h->dout_ready = dmax;
            h->launches++;
            if ((rc = trunk_backward(h, LA, actor, 0, grad_actor, h->dmu, nr * N, s, fa, h->cd.lda))) return rc;
        }
        // 2) critic (one row per env step: the N agent rows of the reference are identical): forward -> value loss -> backward
        const int nc = (int)ns;
        const float *xc = cmp ? nullptr : d_obs + (size_t)s0 * N * D;
        if ((rc = trunk_forward(h, LC, critic, 1, xc, nc, true, s, nullptr, 1, fcr, h->cd.ldc, fh ? h->vnew : nullptr))) return rc;
        if (!fh) {
            critic_head_kernel<<<grid_for_rows(h, nc, 8), 256, 0, s>>>(h->hh[h->la.nblk - 1], critic + LC.Wh, critic + LC.bh, h->vnew, nc, H);
            h->launches++;
        }
        ppo_value_loss_kernel<<<(nc + 127) / 128, 128, 0, s>>>(d_returns + s0, d_values + s0, h->vnew, vn_now, h->dv,
                                                              d_epoch_stats, nc, P);
        h->launches++;
        if ((rc = trunk_backward(h, LC, critic, 1, grad_critic, h->dv, nc, s, fcr, h->cd.ldc))) return rc;
    }
    if (cmp) {
        if ((rc = compact_finalize(h, LA, actor, 0, grad_actor, s))) return rc;
        if ((rc = compact_finalize(h, LC, critic, 1, grad_critic, s))) return rc;
    } else {
        if ((rc = ln0_finalize(h, LA, actor, grad_actor, s))) return rc;
        if ((rc = ln0_finalize(h, LC, critic, grad_critic, s))) return rc;
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_epoch_grads(void *handle, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                          const float *d_obs, const float *d_actions, const float *d_logp_old, const float *d_values,
                          const float *d_returns, float *d_vn_state, const double *d_stats4, double n_rows_global, int T,
                          int E, double *d_epoch_stats, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !actor || !critic || !grad_actor || !grad_critic || !d_obs || !d_actions || !d_logp_old || !d_values ||
        !d_returns || !d_stats4 || !d_epoch_stats || T < 1 || E < 1 || !(n_rows_global >= 1.0))
        return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N > 0) return DCC_ERR_UNSUPPORTED;      // recurrent handles: dcc_mappo_seq_grads
    DCC_DEVICE_GUARD(h->device);
    return epoch_grads_impl(h, actor, critic, grad_actor, grad_critic, d_obs, nullptr, nullptr, d_actions, d_logp_old, d_values,
                            d_returns, d_vn_state, d_stats4, n_rows_global, T, E, d_epoch_stats, static_cast<cudaStream_t>(stream));
}

int dcc_mappo_epoch_grads_state(void *handle, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                                const double *d_pos_vel, const uint8_t *d_energy, const float *d_actions,
                                const float *d_logp_old, const float *d_values, const float *d_returns, float *d_vn_state,
                                const double *d_stats4, double n_rows_global, int T, int E, double *d_epoch_stats,
                                dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !actor || !critic || !grad_actor || !grad_critic || !d_pos_vel || !d_energy || !d_actions || !d_logp_old ||
        !d_values || !d_returns || !d_stats4 || !d_epoch_stats || T < 1 || E < 1 || !(n_rows_global >= 1.0))
        return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    if (!h->compact) return DCC_ERR_UNSUPPORTED;     // dcc_mappo_set_env_layout first
    DCC_DEVICE_GUARD(h->device);
    return epoch_grads_impl(h, actor, critic, grad_actor, grad_critic, nullptr, d_pos_vel, d_energy, d_actions, d_logp_old,
                            d_values, d_returns, d_vn_state, d_stats4, n_rows_global, T, E, d_epoch_stats,
                            static_cast<cudaStream_t>(stream));
}

int dcc_mappo_minibatch_stats(void *handle, const float *d_returns, const int64_t *d_row_index, int64_t n_index,
                              double *d_ret_sums_out, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_returns || !d_row_index || !d_ret_sums_out || n_index < 1) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DCC_CUDA_TRY(cudaMemsetAsync(d_ret_sums_out, 0, 2 * sizeof(double), s));
    const int blocks = (int)std::min<size_t>(((size_t)n_index + 255) / 256, (size_t)h->sm_count * 8);
    sum_sumsq_kernel<<<blocks, 256, 0, s>>>(d_returns, nullptr, nullptr, d_ret_sums_out, (size_t)n_index,
                                            reinterpret_cast<const long long *>(d_row_index), h->cfg.n_agents);
    DCC_CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DCC_OK;
}

static int minibatch_grads_impl(MappoHandle *h, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                                const float *d_obs, const double *d_pv, const uint8_t *d_en, const float *d_actions,
                                const float *d_logp_old, const float *d_values, const float *d_returns, float *d_vn_state,
                                const double *d_stats4, double n_rows_global, const int64_t *d_row_index, int64_t n_index,
                                const double *d_ret_sums, double n_index_global, double *d_epoch_stats, cudaStream_t s) {
    const int N = h->cfg.n_agents, H = h->cfg.hidden;
    const NetLayout &LA = h->la, &LC = h->lc;
    const long long *idx = reinterpret_cast<const long long *>(d_row_index);
    const bool cmp = d_obs == nullptr;      // compact rollout: rows come from the stored env state, gathered through the permutation
    int rc;
    if ((rc = grads_prologue(h, actor, critic, grad_actor, grad_critic, d_vn_state, d_ret_sums, n_index_global, d_epoch_stats, s, cmp)))
        return rc;
    const float *vn_now = h->cfg.use_valuenorm ? d_vn_state : nullptr;
    const PpoLossParams P = loss_params(h, n_index_global);
    const long RA = (long)h->chunk_rows * N;     // agent rows the actor scratch holds
    const float *fa = cmp ? h->x0 : nullptr, *fcr = cmp ? h->fc : nullptr;
    for (long k0 = 0; k0 < n_index; k0 += RA) {
        const int nk = (int)std::min<long>(RA, n_index - k0);
        // actor on the minibatch's agent rows, observation rows (or state features) gathered through the permutation
        const bool fh = fused_head(h);
        if (cmp && (rc = compact_features(h, d_pv, d_en, nk, true, false, s, idx + k0))) return rc;
        if ((rc = trunk_forward(h, LA, actor, 0, d_obs, nk, true, s, cmp ? nullptr : idx + k0, 1, fa, h->cd.lda, fh ? h->mu : nullptr))) return rc;
        if (fh)
            gauss_finish_kernel<<<(nk + 255) / 256, 256, 0, s>>>(h->mu, actor + LA.logstd, const_cast<float *>(d_actions), h->mu, h->logp,
                                                                nk, 1, 0, 0, 0, 0, idx + k0);
        else
            actor_head_kernel<<<grid_for_rows(h, nk, 8), 256, 0, s>>>(h->hh[h->la.nblk - 1], actor + LA.Wh, actor + LA.bh, actor + LA.logstd,
                                                                     const_cast<float *>(d_actions), h->mu, h->logp, nk, H, 1, 0,
                                                                     0, 0, 0, idx + k0);
        h->launches++;
        ppo_policy_loss_mb_kernel<<<(nk + 127) / 128, 128, 0, s>>>(h->mu, h->logp, d_actions, actor + LA.logstd, d_logp_old,
                                                                  d_returns, d_values, vn_snapshot(h), d_stats4, n_rows_global,
                                                                  idx + k0, h->dmu, grad_actor + LA.logstd, d_epoch_stats, nk, P);
        h->launches++;
        if ((rc = trunk_backward(h, LA, actor, 0, grad_actor, h->dmu, nk, s, fa, h->cd.lda))) return rc;
        // critic as the reference evaluates it here: one centralised row PER AGENT ROW of the minibatch (the N agent
        // rows of an env step land in different minibatches, so the once-per-env shortcut does not apply)
        for (long c0 = 0; c0 < nk; c0 += h->chunk_rows) {
            const int nc = (int)std::min<long>(h->chunk_rows, nk - c0);
            const long long *ci = idx + k0 + c0;
            if (cmp && (rc = compact_features(h, d_pv, d_en, nc, false, true, s, ci))) return rc;
            if ((rc = trunk_forward(h, LC, critic, 1, d_obs, nc, true, s, cmp ? nullptr : ci, N, fcr, h->cd.ldc, fh ? h->vnew : nullptr))) return rc;
            if (!fh) {
                critic_head_kernel<<<grid_for_rows(h, nc, 8), 256, 0, s>>>(h->hh[h->la.nblk - 1], critic + LC.Wh, critic + LC.bh, h->vnew, nc, H);
                h->launches++;
            }
            ppo_value_loss_kernel<<<(nc + 127) / 128, 128, 0, s>>>(d_returns, d_values, h->vnew, vn_now, h->dv, d_epoch_stats,
                                                                  nc, P, ci);
            h->launches++;
            if ((rc = trunk_backward(h, LC, critic, 1, grad_critic, h->dv, nc, s, fcr, h->cd.ldc))) return rc;
        }
    }
    if (cmp) {
        if ((rc = compact_finalize(h, LA, actor, 0, grad_actor, s))) return rc;
        if ((rc = compact_finalize(h, LC, critic, 1, grad_critic, s))) return rc;
    } else {
        if ((rc = ln0_finalize(h, LA, actor, grad_actor, s))) return rc;
        if ((rc = ln0_finalize(h, LC, critic, grad_critic, s))) return rc;
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_minibatch_grads(void *handle, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                              const float *d_obs, const float *d_actions, const float *d_logp_old, const float *d_values,
                              const float *d_returns, float *d_vn_state, const double *d_stats4, double n_rows_global,
                              const int64_t *d_row_index, int64_t n_index, const double *d_ret_sums,
                              double n_index_global, double *d_epoch_stats, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !actor || !critic || !grad_actor || !grad_critic || !d_obs || !d_actions || !d_logp_old || !d_values ||
        !d_returns || !d_stats4 || !d_row_index || !d_ret_sums || !d_epoch_stats || n_index < 1 ||
        !(n_rows_global >= 1.0) || !(n_index_global >= 1.0))
        return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N > 0) return DCC_ERR_UNSUPPORTED;      // recurrent handles: dcc_mappo_seq_grads
    DCC_DEVICE_GUARD(h->device);
    return minibatch_grads_impl(h, actor, critic, grad_actor, grad_critic, d_obs, nullptr, nullptr, d_actions, d_logp_old, d_values,
                                d_returns, d_vn_state, d_stats4, n_rows_global, d_row_index, n_index, d_ret_sums, n_index_global,
                                d_epoch_stats, static_cast<cudaStream_t>(stream));
}

int dcc_mappo_minibatch_grads_state(void *handle, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                                    const double *d_pos_vel, const uint8_t *d_energy, const float *d_actions,
                                    const float *d_logp_old, const float *d_values, const float *d_returns, float *d_vn_state,
                                    const double *d_stats4, double n_rows_global, const int64_t *d_row_index, int64_t n_index,
                                    const double *d_ret_sums, double n_index_global, double *d_epoch_stats, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !actor || !critic || !grad_actor || !grad_critic || !d_pos_vel || !d_energy || !d_actions || !d_logp_old ||
        !d_values || !d_returns || !d_stats4 || !d_row_index || !d_ret_sums || !d_epoch_stats || n_index < 1 ||
        !(n_rows_global >= 1.0) || !(n_index_global >= 1.0))
        return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    if (!h->compact) return DCC_ERR_UNSUPPORTED;     // dcc_mappo_set_env_layout first
    DCC_DEVICE_GUARD(h->device);
    return minibatch_grads_impl(h, actor, critic, grad_actor, grad_critic, nullptr, d_pos_vel, d_energy, d_actions, d_logp_old,
                                d_values, d_returns, d_vn_state, d_stats4, n_rows_global, d_row_index, n_index, d_ret_sums,
                                n_index_global, d_epoch_stats, static_cast<cudaStream_t>(stream));
}

// ---- recurrent policies ------------------------------------------------------------------------------------------------------
int dcc_mappo_rnn_pass_seqs(void *handle, int seq_len) {
    MappoHandle *h = as_mappo(handle);
    if (!h || seq_len < 1) return -1;
    return rnn_pass_seqs(h, seq_len);
}

int dcc_mappo_act_rnn(void *handle, const float *actor, const float *critic, const float *d_obs, int n_envs,
                      const float *d_h_actor, const float *d_h_critic, const float *d_masks, int mode, uint64_t seed,
                      uint64_t offset, int deterministic, float *d_actions, float *d_logp, float *d_values, float *d_h_actor_out,
                      float *d_h_critic_out, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !d_obs || !d_masks || n_envs < 1 || (!actor && !critic) || (mode != 0 && mode != 1)) return DCC_ERR_INVALID_ARG;
    if ((actor && (!d_actions || !d_h_actor || (mode == 1 && !d_logp))) || (critic && (!d_values || !d_h_critic))) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N < 1) return DCC_ERR_UNSUPPORTED;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int N = h->cfg.n_agents, D = h->cfg.obs_dim, H = h->cfg.hidden, RN = h->cfg.recurrent_N;
    int rc;
    if (actor && (rc = fold_ln0(h, h->la, actor, 0, false, s))) return rc;
    if (critic && (rc = fold_ln0(h, h->lc, critic, 1, false, s))) return rc;
    if (actor && (rc = rnn_prep_images(h, h->la, actor, 0, false, s))) return rc;
    if (critic && (rc = rnn_prep_images(h, h->lc, critic, 1, false, s))) return rc;
    for (int e0 = 0; e0 < n_envs; e0 += h->chunk_rows) {
        const int ne = min(h->chunk_rows, n_envs - e0);
        const float *x = d_obs + (size_t)e0 * N * D;
        if (actor) {
            const int rows = ne * N;
            rnn_mask_rows_kernel<<<(rows + 255) / 256, 256, 0, s>>>(d_masks + e0, nullptr, N, h->r_mask, rows);
            if ((rc = trunk_forward(h, h->la, actor, 0, x, rows, false, s))) return rc;
            const RnnSrc src{d_h_actor + (size_t)e0 * N * RN * H, nullptr, 1};
            if ((rc = rnn_forward(h, h->la, actor, 0, h->hh[h->la.nblk - 1], rows, 1, src, h->r_mask, s))) return rc;
            for (int l = 0; l < RN && d_h_actor_out; ++l)
                gru_store_state_kernel<<<ew_grid((size_t)rows * H), 256, 0, s>>>(h->r_Hout[l], d_h_actor_out, RN, l, rows, H, (size_t)e0 * N);
            actor_head_kernel<<<grid_for_rows(h, rows, 8), 256, 0, s>>>(
                h->r_Y, actor + h->la.Wh, actor + h->la.bh, actor + h->la.logstd, d_actions + (size_t)e0 * N * 2, nullptr,
                d_logp ? d_logp + (size_t)e0 * N : nullptr, rows, H, mode, deterministic, seed, offset, (uint64_t)e0 * N);
            h->launches += 2 + RN;
        }
        if (critic) {     // one row per env: the N agent rows of an env carry identical inputs, masks and hidden states
            rnn_mask_rows_kernel<<<(ne + 255) / 256, 256, 0, s>>>(d_masks + e0, nullptr, 1, h->r_mask, ne);
            if ((rc = trunk_forward(h, h->lc, critic, 1, x, ne, false, s))) return rc;
            const RnnSrc src{d_h_critic + (size_t)e0 * RN * H, nullptr, 1};
            if ((rc = rnn_forward(h, h->lc, critic, 1, h->hh[h->lc.nblk - 1], ne, 1, src, h->r_mask, s))) return rc;
            for (int l = 0; l < RN && d_h_critic_out; ++l)
                gru_store_state_kernel<<<ew_grid((size_t)ne * H), 256, 0, s>>>(h->r_Hout[l], d_h_critic_out, RN, l, ne, H, (size_t)e0);
            critic_head_kernel<<<grid_for_rows(h, ne, 8), 256, 0, s>>>(h->r_Y, critic + h->lc.Wh, critic + h->lc.bh, d_values + e0, ne, H);
            h->launches += 2 + RN;
        }
    }
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_seq_grads(void *handle, const float *actor, const float *critic, float *grad_actor, float *grad_critic,
                        const float *d_obs, const float *d_h_actor, const float *d_h_critic, const float *d_masks,
                        const float *d_actions, const float *d_logp_old, const float *d_values, const float *d_returns,
                        float *d_vn_state, const double *d_stats4, double n_rows_global, const int64_t *d_row_index, int64_t n_seq,
                        int seq_len, const double *d_ret_sums, double n_index_global, double *d_epoch_stats, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !actor || !critic || !grad_actor || !grad_critic || !d_obs || !d_h_actor || !d_h_critic || !d_masks || !d_actions ||
        !d_logp_old || !d_values || !d_returns || !d_stats4 || !d_row_index || !d_ret_sums || !d_epoch_stats || n_seq < 1 ||
        seq_len < 1 || !(n_rows_global >= 1.0) || !(n_index_global >= 1.0))
        return DCC_ERR_INVALID_ARG;
    if (h->cfg.use_valuenorm && !d_vn_state) return DCC_ERR_INVALID_ARG;
    if (h->cfg.recurrent_N < 1 || rnn_pass_seqs(h, seq_len) < 1) return DCC_ERR_UNSUPPORTED;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int N = h->cfg.n_agents, H = h->cfg.hidden;
    const NetLayout &LA = h->la, &LC = h->lc;
    int rc;
    if ((rc = grads_prologue(h, actor, critic, grad_actor, grad_critic, d_vn_state, d_ret_sums, n_index_global, d_epoch_stats, s, false)))
        return rc;
    if ((rc = rnn_prep_images(h, LA, actor, 0, true, s))) return rc;
    if ((rc = rnn_prep_images(h, LC, critic, 1, true, s))) return rc;
    const float *vn_now = h->cfg.use_valuenorm ? d_vn_state : nullptr;
    const PpoLossParams P = loss_params(h, n_index_global);
    const long Sc = rnn_pass_seqs(h, seq_len);
    const long long *idx = reinterpret_cast<const long long *>(d_row_index);
    for (long q0 = 0; q0 < n_seq; q0 += Sc) {
        const int S = (int)std::min<long>(Sc, n_seq - q0);
        const int rows = S * seq_len;
        const long long *pidx = idx + (size_t)q0 * seq_len;          // this pass: [seq_len, S] time-major agent-row indices
        // the step masks of the pass's rows (masks are stored once per env step: agent row / N)
        rnn_mask_rows_kernel<<<(rows + 255) / 256, 256, 0, s>>>(d_masks, pidx, N, h->r_mask, rows);
        h->launches++;
        // 1) actor: trunk -> GRU -> head -> policy loss -> backward through head / GRU (BPTT inside the pass) / trunk
        if ((rc = trunk_forward(h, LA, actor, 0, d_obs, rows, true, s, pidx, 1))) return rc;
        const RnnSrc sa{d_h_actor, pidx, 1};
        if ((rc = rnn_forward(h, LA, actor, 0, h->hh[LA.nblk - 1], S, seq_len, sa, h->r_mask, s))) return rc;
        actor_head_kernel<<<grid_for_rows(h, rows, 8), 256, 0, s>>>(h->r_Y, actor + LA.Wh, actor + LA.bh, actor + LA.logstd,
                                                                   const_cast<float *>(d_actions), h->mu, h->logp, rows, H, 1, 0, 0, 0, 0, pidx);
        ppo_policy_loss_mb_kernel<<<(rows + 127) / 128, 128, 0, s>>>(h->mu, h->logp, d_actions, actor + LA.logstd, d_logp_old, d_returns,
                                                                    d_values, vn_snapshot(h), d_stats4, n_rows_global, pidx, h->dmu,
                                                                    grad_actor + LA.logstd, d_epoch_stats, rows, P);
        h->launches += 2;
        if ((rc = rnn_backward(h, LA, actor, 0, grad_actor, h->hh[LA.nblk - 1], h->dmu, S, seq_len, h->r_mask, s))) return rc;
        if ((rc = trunk_backward(h, LA, actor, 0, grad_actor, nullptr, rows, s, nullptr, 0, h->r_dfeat))) return rc;
        // 2) critic, as the reference evaluates it here: one centralised row per AGENT row of the sequences (hidden states and
        //    masks are those of the env step: row / N)
        if ((rc = trunk_forward(h, LC, critic, 1, d_obs, rows, true, s, pidx, N))) return rc;
        const RnnSrc sc{d_h_critic, pidx, N};
        if ((rc = rnn_forward(h, LC, critic, 1, h->hh[LC.nblk - 1], S, seq_len, sc, h->r_mask, s))) return rc;
        critic_head_kernel<<<grid_for_rows(h, rows, 8), 256, 0, s>>>(h->r_Y, critic + LC.Wh, critic + LC.bh, h->vnew, rows, H);
        ppo_value_loss_kernel<<<(rows + 127) / 128, 128, 0, s>>>(d_returns, d_values, h->vnew, vn_now, h->dv, d_epoch_stats, rows, P, pidx);
        h->launches += 2;
        if ((rc = rnn_backward(h, LC, critic, 1, grad_critic, h->hh[LC.nblk - 1], h->dv, S, seq_len, h->r_mask, s))) return rc;
        if ((rc = trunk_backward(h, LC, critic, 1, grad_critic, nullptr, rows, s, nullptr, 0, h->r_dfeat))) return rc;
    }
    if ((rc = ln0_finalize(h, LA, actor, grad_actor, s))) return rc;
    if ((rc = ln0_finalize(h, LC, critic, grad_critic, s))) return rc;
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_mappo_apply(void *handle, int which, float *params, float *grads, float *adam_m, float *adam_v, float lr,
                    int64_t step, double *d_grad_norm_sq_out, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !params || !grads || !adam_m || !adam_v || step < 1 || (which != 0 && which != 1)) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const NetLayout &L = which == 0 ? h->la : h->lc;
    if (which == 0) {
        // d(-entropy_coef * dist_entropy)/dlogstd = -entropy_coef per action dim (act.py:172-184)
        add_scalar_kernel<<<1, 32, 0, s>>>(grads + L.logstd, h->cfg.act_dim, -h->cfg.entropy_coef);
        h->launches++;
    }
    double *sq = h->dsums + 4 + which;
    DCC_CUDA_TRY(cudaMemsetAsync(sq, 0, sizeof(double), s));
    const int blocks = (int)std::min<size_t>((L.total + 255) / 256, (size_t)h->sm_count * 4);
    sumsq_kernel<<<blocks, 256, 0, s>>>(grads, L.total, sq);
    const float b1 = h->cfg.adam_beta1, b2 = h->cfg.adam_beta2;
    const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
    const float bc2s = (float)sqrt(1.0 - pow((double)b2, (double)step));
    clip_adam_kernel<<<blocks, 256, 0, s>>>(params, grads, adam_m, adam_v, L.total, sq, h->cfg.max_grad_norm, lr, b1, b2,
                                            h->cfg.opti_eps, bc1, bc2s, 1.0f, h->cfg.use_max_grad_norm, h->cfg.weight_decay);
    h->launches += 2;
    if (d_grad_norm_sq_out) DCC_CUDA_TRY(cudaMemcpyAsync(d_grad_norm_sq_out, sq, sizeof(double), cudaMemcpyDeviceToDevice, s));
    DCC_CUDA_TRY(cudaGetLastError());
    return DCC_OK;
}

int dcc_op_gemm(void *handle, int backend, int ta, int tb, int M, int N, int K, const float *A, int lda, const float *B,
                int ldb, float *C, int ldc, int accumulate, dcc_stream_t stream) {
    MappoHandle *h = as_mappo(handle);
    if (!h || !A || !B || !C || backend < 0 || backend > 5 || M < 1 || N < 1 || K < 1) return DCC_ERR_INVALID_ARG;
    DCC_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (backend == 0) backend = h->backend;
    if (backend == 4 || backend == 5) {
        // test hooks of the TMA-fed operand paths: 4 = forward-shaped GEMM with a PRE-SPLIT A operand (X W^T and dZ W shapes),
        // 5 = weight-gradient GEMM dW = dZ^T X with a pre-split X operand (fp16-split kernel, max |dZ| from a separate pass)
        if (accumulate) return DCC_ERR_UNSUPPORTED;
        const bool wg = ta && !tb && M == tc::TC_N;
        if ((backend == 5) != wg || (backend == 4 && (ta || N != tc::TC_N))) return DCC_ERR_UNSUPPORTED;
        const float *src = wg ? B : A;                 // the operand that is pre-split: X [K rows, N cols] / A [M rows, K cols]
        const long srows = wg ? K : M;
        const int scols = wg ? N : K, sld = wg ? ldb : lda, ld16 = (scols + 7) / 8 * 8;
        __half *hi = nullptr, *lo = nullptr;
        float *img = nullptr;
        DCC_CUDA_TRY(cudaMalloc(&hi, (size_t)srows * ld16 * 2));
        DCC_CUDA_TRY(cudaMalloc(&lo, (size_t)srows * ld16 * 2));
        tc::tc_split16_kernel<<<h->sm_count * 8, 256, 0, s>>>(src, srows, scols, sld, hi, lo, ld16);
        const Split16 s16{hi, lo, ld16};
        int rc;
        if (wg) {
            DCC_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, s));
            rc = tc_gemm_wgrad(h, K, N, A, lda, nullptr, 0, C, ldc, s, true, &s16, false);
        } else {
            const int KT = (K + tc::TC_BK - 1) / tc::TC_BK;
            DCC_CUDA_TRY(cudaMalloc(&img, (size_t)KT * 2 * tc::TC_B_TILE_FLOATS * sizeof(float)));
            rc = tc_prep_weights(h, B, ldb, tb == 0, K, img, s, true);
            if (!rc) rc = tc_gemm_fwd(h, M, K, nullptr, 0, img, C, ldc, s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true,
                                      nullptr, nullptr, 0, nullptr, &s16, nullptr);
        }
        cudaStreamSynchronize(s);
        cudaFree(hi); cudaFree(lo); cudaFree(img);
        return rc;
    }
    if (backend == 2 || backend == 3) {
        // shapes the tensor-core kernels cover: X W^T and dZ W (weights on the B side, 256 output features);
        // backend 3 = the fp16-split forward kernel (|A| < 65504, |B| < 255)
        const bool f16 = backend == 3;
        if (ta && !tb && M == tc::TC_N) {   // dW = dZ^T X (backend 3: the experimental fp16-split weight-gradient kernel)
            if (!accumulate) DCC_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, s));
            return tc_gemm_wgrad(h, K, N, A, lda, B, ldb, C, ldc, s, f16);
        }
        if (ta || N != tc::TC_N || accumulate) return DCC_ERR_UNSUPPORTED;
        const int KT = (K + tc::TC_BK - 1) / tc::TC_BK;
        float *img = nullptr;
        DCC_CUDA_TRY(cudaMalloc(&img, (size_t)KT * 2 * tc::TC_B_TILE_FLOATS * sizeof(float)));
        int rc = tc_prep_weights(h, B, ldb, tb == 0, K, img, s, f16);
        if (!rc) rc = tc_gemm_fwd(h, M, K, A, lda, img, C, ldc, s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, f16);
        cudaStreamSynchronize(s);
        cudaFree(img);
        return rc;
    }
    const int saved = h->backend;
    h->backend = 1;
    const int rc = launch_gemm(h, ta != 0, tb != 0, M, N, K, A, lda, B, ldb, C, ldc, accumulate != 0, s);
    h->backend = saved;
    return rc;
}

}  // extern "C"

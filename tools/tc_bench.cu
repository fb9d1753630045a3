// tools/tc_bench.cu — standalone timing + in-kernel role profile of the tcgen05 GEMM kernels (tuning aid, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DDCC_TC_PROFILE -I include \
//        -I dynamic-coverage-control_b200/csrc -o tools/tc_bench.bin tools/tc_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "dcc_tc.cuh"
namespace dcc { void set_last_cuda_error(cudaError_t, const char *, const char *, int) {} }
using namespace dcc::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// weight-gradient kernel: tc_bench.bin wgrad R Nout [f16]   (G[256, Nout] += dZ[R,256]^T X[R, Nout])
static int bench_wgrad(int R, int Nout, bool f16, int split = 0) {
    const int ldx = (Nout + 31) / 32 * 32;
    float *dZ, *X, *G;
    CK(cudaMalloc(&dZ, (size_t)R * 256 * 4)); CK(cudaMalloc(&X, (size_t)R * ldx * 4)); CK(cudaMalloc(&G, (size_t)256 * Nout * 4));
    CK(cudaMemset(dZ, 0x11, (size_t)R * 256 * 4)); CK(cudaMemset(X, 0x11, (size_t)R * ldx * 4)); CK(cudaMemset(G, 0, (size_t)256 * Nout * 4));
    auto kern = f16 ? tc_gemm_wgrad_kernel<true> : tc_gemm_wgrad_kernel<false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM_BYTES));
    uint32_t *absmax;
    CK(cudaMalloc(&absmax, 4)); CK(cudaMemset(absmax, 0, 4));
    tc_absmax_bits_kernel<<<148 * 8, 256>>>(dZ, (long)R, 256, 256, absmax);
    const int bkw = f16 ? 64 : TC_BK;
    TcwParams p; memset(&p, 0, sizeof p);
    p.dZ = dZ; p.X = X; p.G = G; p.R = R; p.Nout = Nout; p.ldz = 256; p.ldx = ldx; p.ldg = Nout;
    p.n_tiles = (Nout + TC_N - 1) / TC_N;
    p.tile_n = ((Nout + p.n_tiles - 1) / p.n_tiles + 31) / 32 * 32;
    const int out_tiles = 2 * p.n_tiles, sms = 148;
    int ks = (2 * sms) / out_tiles;
    const int max_ks = (R + 8 * bkw - 1) / (8 * bkw);
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    p.rows_per_split = ((R + ks - 1) / ks + bkw - 1) / bkw * bkw;
    p.dz_absmax_bits = absmax;
    p.ksplits = (R + p.rows_per_split - 1) / p.rows_per_split;
    if (split && f16) {     // split 1: X pre-split (TMA); split 2: dZ pre-split too
        const int ld16 = (Nout + 7) / 8 * 8;
        void *xhi, *xlo, *zhi, *zlo;
        CK(cudaMalloc(&xhi, (size_t)R * ld16 * 2)); CK(cudaMalloc(&xlo, (size_t)R * ld16 * 2));
        CK(cudaMemset(xhi, 0x11, (size_t)R * ld16 * 2)); CK(cudaMemset(xlo, 0x01, (size_t)R * ld16 * 2));
        p.x_split = 1;
        if (!tc_make_map_2d_f16(&p.tmXhi, xhi, Nout, R, ld16, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc_make_map_2d_f16(&p.tmXlo, xlo, Nout, R, ld16, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("X maps failed\n"); return 1; }
        if (split > 1) {
            CK(cudaMalloc(&zhi, (size_t)R * 256 * 2)); CK(cudaMalloc(&zlo, (size_t)R * 256 * 2));
            CK(cudaMemset(zhi, 0x11, (size_t)R * 256 * 2)); CK(cudaMemset(zlo, 0x01, (size_t)R * 256 * 2));
            p.dz_split = 1;
            if (!tc_make_map_2d_f16(&p.tmZhi, zhi, 256, R, 256, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
                !tc_make_map_2d_f16(&p.tmZlo, zlo, 256, R, 256, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("Z maps failed\n"); return 1; }
        }
    }
    const int work = out_tiles * p.ksplits, grid = work < sms ? work : sms;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, TCF_THREADS, TCF_SMEM_BYTES>>>(p);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) kern<<<grid, TCF_THREADS, TCF_SMEM_BYTES>>>(p);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("wgrad%s R=%d Nout=%d (tile_n %d, %d work items, %d rows/split): %.1f us  %.1f TFLOP/s fp32-equivalent\n", f16 ? " [fp16 split]" : "", R, Nout, p.tile_n,
           work, p.rows_per_split, ms * 1e3, 2.0 * R * 256.0 * Nout / ms * 1e-9);
#ifdef DCC_TC_PROFILE
    unsigned long long prof[32];
    CK(cudaMemcpyFromSymbol(prof, g_tc_prof, sizeof prof));
    const double st = (double)prof[16];
    const int units = f16 ? 6 : 3;
    printf("CTA0: %llu stages | mma: wait_tempty %.0f wait_full %.0f issue %.0f, total %.0f cyc/stage | producer warp 0: wait_empty %.0f wait_cp.async %.0f (per unit, %d units/stage), total %.0f cyc/stage\n",
           prof[16], prof[17] / st, prof[18] / st, prof[19] / st, prof[20] / st, prof[21] / st, prof[22] / (units * st), units, prof[23] / st);
#endif
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 3 && !strcmp(argv[1], "wgrad")) return bench_wgrad(atoi(argv[2]), atoi(argv[3]), argc > 4 && atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 0);
    const int M = argc > 1 ? atoi(argv[1]) : 198408, K = argc > 2 ? atoi(argv[2]) : 352, epi = argc > 3 ? atoi(argv[3]) : 1;
    const bool f16 = argc > 6 && atoi(argv[6]);     // fp16 hi/lo split kernel (64 reduction elements per stage)
    const int KT = f16 ? (K + 63) / 64 : (K + 31) / 32;
    float *A, *W, *img, *C, *H, *bias, *mean, *rstd;
    CK(cudaMalloc(&A, (size_t)M * K * 4)); CK(cudaMalloc(&W, (size_t)256 * K * 4));
    CK(cudaMalloc(&img, (size_t)KT * 2 * TC_B_TILE_FLOATS * 4));
    CK(cudaMalloc(&C, (size_t)M * 256 * 4)); CK(cudaMalloc(&H, (size_t)M * 256 * 4));
    CK(cudaMalloc(&bias, 3 * 256 * 4)); CK(cudaMalloc(&mean, (size_t)M * 4)); CK(cudaMalloc(&rstd, (size_t)M * 4));
    std::vector<float> h((size_t)M * K);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    CK(cudaMemcpy(A, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(W, h.data(), (size_t)256 * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(bias, 0, 3 * 256 * 4));
    if (f16) tc_prep_weights_f16_kernel<<<(KT * 256 * 8 + 255) / 256, 256>>>(W, K, 0, K, KT, 256.f, img);
    else tc_prep_weights_kernel<<<(KT * 256 * 8 + 255) / 256, 256>>>(W, K, 0, K, KT, img);
    auto kern = f16 ? tc_gemm_fwd_kernel<true> : tc_gemm_fwd_kernel<false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM_BYTES));
    TcfParams p; memset(&p, 0, sizeof p);
    p.A = A; p.Bimg = img; p.C = C; p.M = M; p.K = K; p.KT = KT; p.lda = K; p.ldc = 256; p.splits = 1; p.kt_per_split = KT;
    p.out_scale = f16 ? 1.f / 256.f : 1.f;
    p.epi = epi; p.dbg = argc > 4 ? atoi(argv[4]) : 0; p.bias = bias; p.gamma = bias + 256; p.beta = bias + 512; p.H = H; p.mean = mean; p.rstd = rstd;
    if (argc > 5 && atoi(argv[5])) {   // TMA-store epilogue
        p.use_tma = (tc_make_store_map(&p.tmC, C, M, 256) && tc_make_store_map(&p.tmH, H, M, 256)) ? 1 : 0;
        printf("use_tma=%d  ", p.use_tma);
    }
    p.pf_dist = argc > 7 ? atoi(argv[7]) : 0;     // L2 prefetch distance of the activation operand, in stages
    const int head_out = argc > 8 ? atoi(argv[8]) : 0;   // fused output head (0 / 1 / 2)
    const int store_h = argc > 9 ? atoi(argv[9]) : 1;    // 0 = do not store h (update-mode last block)
    const int store_c = argc > 10 ? atoi(argv[10]) : 1;  // 0 = do not store a (rollout mode)
    float *hw, *hdst;
    CK(cudaMalloc(&hw, 4 * 256 * 4)); CK(cudaMemset(hw, 0, 4 * 256 * 4)); CK(cudaMalloc(&hdst, (size_t)M * 2 * 4));
    if (head_out > 0 && epi == 1) { p.head_out = head_out; p.head_fold = hw; p.head_dst = hdst; }
    if (!store_h) p.H = nullptr;
    if (!store_c) p.C = nullptr;
    // product configurations of round 2: pre-split A operand fetched by TMA (a_split), pre-split h output (h_split), unit affine (xhat mode)
    const int a_split = argc > 11 ? atoi(argv[11]) : 0, h_split = argc > 12 ? atoi(argv[12]) : 0, unit = argc > 13 ? atoi(argv[13]) : 0;
    if (a_split && f16) {
        const int ld16 = (K + 7) / 8 * 8;
        void *ahi, *alo;
        CK(cudaMalloc(&ahi, (size_t)M * ld16 * 2)); CK(cudaMalloc(&alo, (size_t)M * ld16 * 2));
        CK(cudaMemset(ahi, 0x11, (size_t)M * ld16 * 2)); CK(cudaMemset(alo, 0x01, (size_t)M * ld16 * 2));
        p.a_split = 1;
        if (!tc_make_map_2d_f16(&p.tmAhi, ahi, K, M, ld16, TC_BK16, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc_make_map_2d_f16(&p.tmAlo, alo, K, M, ld16, TC_BK16, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("A maps failed\n"); return 1; }
        p.pf_dist = 0;
    }
    if (h_split && p.H && p.use_tma) {
        void *hhi, *hlo;
        CK(cudaMalloc(&hhi, (size_t)M * 256 * 2)); CK(cudaMalloc(&hlo, (size_t)M * 256 * 2));
        p.h_split = 1;
        if (!tc_make_map_2d_f16(&p.tmHhi, hhi, TC_N, M, 256, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc_make_map_2d_f16(&p.tmHlo, hlo, TC_N, M, 256, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) { printf("H maps failed\n"); return 1; }
    }
    p.unit_affine = unit;
    printf("head=%d store_h=%d store_c=%d a_split=%d h_split=%d unit=%d  ", head_out, store_h, store_c, p.a_split, p.h_split, unit);
    if (p.pf_dist > 0 && !tc_make_prefetch_map(&p.tmA, A, M, K, K, f16 ? 64 : 32)) { printf("prefetch map failed\n"); p.pf_dist = 0; }
    printf("pf=%d  ", p.pf_dist);
    const int tiles = (M + 127) / 128, grid = tiles < 148 ? tiles : 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, TCF_THREADS, TCF_SMEM_BYTES>>>(p);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) kern<<<grid, TCF_THREADS, TCF_SMEM_BYTES>>>(p);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double flop = 2.0 * M * 256.0 * (KT * (f16 ? 64.0 : 32.0));
    printf("fwd%s M=%d K=%d epi=%d: %.1f us  %.1f TFLOP/s fp32-equivalent (x3 = %.0f TF MMA)\n", f16 ? " [fp16 split]" : "", M, K, epi, ms * 1e3,
           flop / ms * 1e-9, 3 * flop / ms * 1e-9);
#ifndef DCC_TC_PROFILE
    return 0;   // timing-only build (no -DDCC_TC_PROFILE): the role timers cost a few percent
#else
    unsigned long long prof[32];
    CK(cudaMemcpyFromSymbol(prof, g_tc_prof, sizeof prof));
    const double st = (double)prof[2];
    printf("CTA0: %llu stages | producer: wait_empty %.0f work %.0f cyc/stage | mma: wait_acc %.0f wait_full %.0f issue %.0f, total %.0f cyc/stage\n",
           prof[2], prof[0] / st, prof[1] / st, prof[4] / st, prof[5] / st, prof[6] / st, prof[7] / st);
    const double tl = st / KT;
    printf("      epilogue warp 4: wait_tfull %.0f drain %.0f cyc/stage, tile epilogue %.0f cyc/tile of which bias/relu/LN-stats(+head) %.0f [LN barriers %.0f, head %.0f], "
           "wait for staging reuse %.0f (%.1f tiles)\n", prof[8] / st,
           prof[9] / st, prof[10] / tl, prof[11] / tl, prof[14] / tl, prof[12] / tl, prof[13] / tl, tl);
    return 0;
#endif
}

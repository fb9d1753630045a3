// tools/mn16_probe.cu — pins the shared-memory descriptor semantics of MN-major 16-bit operands for tcgen05.mma
// kind::f16 (tuning aid for the planned fp16-split weight-gradient kernel; not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I dynamic-coverage-control_b200/csrc \
//        -o tools/mn16_probe.bin tools/mn16_probe.cu
//   tools/mn16_probe.bin [lbo_bytes sbo_bytes layout_type kstep_bytes]      (no arguments: sweeps the candidates)
// One CTA computes D[128, 256] = A^T B for A[64 k][128 m], B[64 k][256 n] (fp16, row-major in global memory = MN-major
// for this product) from the layout a TMA load with SWIZZLE_128B and a {64 features, 64 rows} box would produce:
// per 64-feature group a block of [64 k rows][128 B], 16-byte chunks XOR-ed with (k & 7).  The result is compared with a
// float64 reference on the host; the candidate (LBO, SBO, layout type, per-K-step descriptor advance) that reproduces it
// is the one to use.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "dcc_tc.cuh"
namespace dcc { void set_last_cuda_error(cudaError_t, const char *, const char *, int) {} }
using namespace dcc::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int PK = 64, PM = 128, PN = 256;
constexpr int GROUP_BYTES = PK * 128;                       // one 64-feature group: 64 k rows x 128 B
constexpr int A_BYTES = (PM / 64) * GROUP_BYTES, B_BYTES = (PN / 64) * GROUP_BYTES;

__global__ void __launch_bounds__(128, 1) mn16_probe_kernel(const __half *A, const __half *B, float *D, uint32_t lbo,
                                                            uint32_t sbo, uint32_t ltype, uint32_t kstep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + A_BYTES + B_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    // operand tiles: 16-byte chunk (row k, features 8c..8c+7) -> group (c >> 3), row k, chunk ((c & 7) ^ (k & 7))
    for (int i = threadIdx.x; i < PK * (PM / 8); i += blockDim.x) {
        const int k = i / (PM / 8), c = i % (PM / 8);
        const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)k * PM + c * 8);
        *reinterpret_cast<uint4 *>(smem + (c >> 3) * GROUP_BYTES + k * 128 + (((c & 7) ^ (k & 7)) << 4)) = v;
    }
    for (int i = threadIdx.x; i < PK * (PN / 8); i += blockDim.x) {
        const int k = i / (PN / 8), c = i % (PN / 8);
        const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)k * PN + c * 8);
        *reinterpret_cast<uint4 *>(smem + A_BYTES + (c >> 3) * GROUP_BYTES + k * 128 + (((c & 7) ^ (k & 7)) << 4)) = v;
    }
    dcc::fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(PM, PN, 1, 1);
        const uint64_t da = make_desc_sw128(smem_u32(smem), lbo, sbo, ltype);
        const uint64_t db = make_desc_sw128(smem_u32(smem + A_BYTES), lbo, sbo, ltype);
        for (int k = 0; k < PK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * kstep) >> 4);
            tc_mma_f16(tmem_base, da + adv, db + adv, idesc, k != 0 ? 1u : 0u);
        }
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int j = 0; j < PN / 32; ++j) {
        float v[32];
        tmem_ld32(tmem_base + j * 32 + ((uint32_t)(warp * 32) << 16), v);
        for (int i = 0; i < 32; ++i) D[(size_t)m * PN + j * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

static int run(const __half *dA, const __half *dB, float *dD, const std::vector<double> &ref, uint32_t lbo, uint32_t sbo,
               uint32_t ltype, uint32_t kstep) {
    CK(cudaMemset(dD, 0xff, (size_t)PM * PN * 4));
    mn16_probe_kernel<<<1, 128, A_BYTES + B_BYTES + 2048>>>(dA, dB, dD, lbo, sbo, ltype, kstep);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)PM * PN);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) {
        const double e = std::isfinite(D[i]) ? fabs(D[i] - ref[i]) : 1e30;
        if (e > worst) worst = e;
        if (e > 1e-3) ++bad;
    }
    printf("LBO %5u SBO %5u layout %u k-step %5u B: max |err| %.3e, %d of %d elements off  %s\n", lbo, sbo, ltype, kstep, worst, bad,
           PM * PN, bad == 0 ? "<== MATCH" : "");
    return 0;
}

int main(int argc, char **argv) {
    std::vector<__half> A((size_t)PK * PM), B((size_t)PK * PN);
    std::vector<float> Af(A.size()), Bf(B.size());
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 9) & 0x3ff) / 512.f - 1.f; };
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half(rnd()); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half(rnd()); Bf[i] = __half2float(B[i]); }
    std::vector<double> ref((size_t)PM * PN, 0.0);
    for (int k = 0; k < PK; ++k)
        for (int m = 0; m < PM; ++m)
            for (int n = 0; n < PN; ++n) ref[(size_t)m * PN + n] += (double)Af[(size_t)k * PM + m] * Bf[(size_t)k * PN + n];
    __half *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, (size_t)PM * PN * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mn16_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_BYTES + B_BYTES + 2048));
    if (argc > 4) return run(dA, dB, dD, ref, atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    // expected: LBO = distance between 64-feature groups (8192 B), SBO = distance between 8-row atoms (1024 B),
    // SWIZZLE_128B (layout type 2), one K = 16 step = two atoms = 2048 B; the others are the plausible misreadings
    const uint32_t cand[][4] = {{8192, 1024, 2, 2048}, {1024, 8192, 2, 2048}, {8192, 1024, 2, 1024}, {8192, 2048, 2, 2048},
                                {8192, 1024, 1, 2048}, {8192, 512, 1, 2048},  {16, 1024, 2, 2048},   {8192, 1024, 2, 256}};
    for (auto &c : cand)
        if (run(dA, dB, dD, ref, c[0], c[1], c[2], c[3])) return 1;
    return 0;
}

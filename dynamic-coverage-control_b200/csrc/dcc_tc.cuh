// dcc_tc.cuh — tcgen05 (5th-gen tensor core) GEMM kernels of the MAPPO learner path, sm_100a only.
//
// Precision: 3xTF32 split.  Every fp32 operand x is split into hi = tf32(x) and lo = tf32(x - hi); the product is
// accumulated in fp32 TMEM as  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the lo*lo term, ~2^-22 relative, is dropped).  This
// keeps logits / values within fp32 round-off of the reference's fp32 GEMMs (plain TF32 or bf16
// miss the 1e-5 parity bar by 1-3 orders of magnitude) at one third of the TF32 tensor rate.  Where the activation
// operand is a LayerNorm output (bounded), the forward kernel uses an fp16 hi/lo split instead (11 + 11 significand
// bits, same accuracy class): half the operand bytes and twice the MMA rate (tc_gemm_fwd_kernel<true>).
//
// Operand staging: tcgen05.mma reads both operands from shared memory in the canonical 128-byte-swizzled layouts
// (8 rows x 128 B atoms, 16-byte chunks XOR-ed with the row index inside the atom):
//   * activations are loaded by producer warps with coalesced 16-byte global loads, split into hi/lo in registers and
//     stored straight into the swizzled layout (no TMA round trip: the split has to touch every element anyway);
//   * weights are pre-split and pre-swizzled once per call into a global "image" of the shared-memory tiles, so a
//     whole K-tile (hi + lo, 64 KB) arrives with cp.async.bulk (TMA, SASS UBLKCP) on an mbarrier.
// Accumulators live in TMEM (2 x 256 columns, double-buffered so the epilogue of tile i overlaps the MMAs of tile
// i+1); one thread issues the MMAs; tcgen05.commit arrives on the mbarriers that recycle the stages.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <cuda_fp16.h>

#include "dcc_common.cuh"

namespace dcc {
namespace tc {

// ---- optional in-kernel role timing (tools/tc_bench.cu builds with -DDCC_TC_PROFILE; off in the library) ----------
#ifdef DCC_TC_PROFILE
__device__ unsigned long long g_tc_prof[32];
#define TC_PROF_DECL(...) unsigned long long __VA_ARGS__
#define TC_PROF_NOW(t) t = clock64()
#define TC_PROF_ADD(acc, t0, t1) acc += (t1) - (t0)
#define TC_PROF_OUT(cond, idx, val) do { if ((cond) && blockIdx.x == 0) g_tc_prof[idx] = (val); } while (0)
#else
#define TC_PROF_DECL(...)
#define TC_PROF_NOW(t)
#define TC_PROF_ADD(acc, t0, t1)
#define TC_PROF_OUT(cond, idx, val)
#endif

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space accesses: the stage pointers are derived by integer alignment of the dynamic shared-memory
// base, which makes the compiler fall back to generic LD/ST (slow address-space resolution) unless told otherwise
__device__ __forceinline__ void sts128(uint32_t saddr, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t saddr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}

// 16-byte read-only load that does not allocate in L1: the activation tiles stream through once (16-32 KB per K-stage),
// and with the shared-memory carve-out at its maximum the L1 is ~28 KB — allocating them there evicted the per-column
// vectors of the epilogue (bias, gamma, beta, head weights) on every stage, so each of THEIR loads paid an L2 round trip
// (measured: 10 k cycles per tile for the fused head, profiles/r02g_tc_fwd_role_profile.txt)
#ifndef DCC_TC_STREAM_LD
#define DCC_TC_STREAM_LD 0
#endif
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
#if DCC_TC_STREAM_LD
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// 16-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_load_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, kind::f16 (fp16 operands, 16 reduction elements per instruction, twice the tf32 rate)
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// round-to-nearest (ties away) to the 10-bit tf32 mantissa: what cvt.rna.tf32.f32 computes for finite inputs, without
// the NaN/Inf special-casing its lowering carries (activations and weights here are finite)
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(const float4 &x, float4 &hi, float4 &lo) {
    hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
    lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
}

// fp16 hi/lo split of 8 consecutive fp32 values: hi = fp16(x), lo = fp16(x - hi), packed in memory order (element 0 in
// the low half of the first word).  11 + 11 significand bits: the same 2^-22 class as the tf32 split, in half the bytes.
__device__ __forceinline__ uint32_t h2_bits(const __half2 &h) { return *reinterpret_cast<const uint32_t *>(&h); }
__device__ __forceinline__ void split_f16_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(a - f.x, b - f.y));
}
__device__ __forceinline__ void split_f16x8(const float4 &a, const float4 &b, uint4 &hi, uint4 &lo) {
    split_f16_pair(a.x, a.y, hi.x, lo.x);
    split_f16_pair(a.z, a.w, hi.y, lo.y);
    split_f16_pair(b.x, b.y, hi.z, lo.z);
    split_f16_pair(b.z, b.w, hi.w, lo.w);
}

// Shared-memory matrix descriptor (sm_100 UMMA), SWIZZLE_128B: start address, leading / stride byte offsets in
// 16-byte units, descriptor version 1, layout type 2 (bits 61-63).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                    uint64_t layout_type = 2) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (layout_type << 61);
}
// MN-major operands of 32-bit types have ONE legal swizzled layout, SWIZZLE_128B_BASE32B (layout type 1): rows of
// 128 B (32 features) per reduction index, atoms of 4 rows, the four 32-byte chunks of a row XOR-ed with (row & 3).
__device__ __forceinline__ uint32_t mn32_offset(int group, int k, int chunk16, int rows_per_group) {
    return (uint32_t)(group * rows_per_group * 128 + k * 128 + (((((chunk16 >> 1) ^ (k & 3)) << 1) | (chunk16 & 1)) << 4));
}
// Instruction descriptor, kind::tf32: D fp32 (bits 4-5 = 1), A/B tf32 (bits 7-9 / 10-12 = 2), majors (bit 15 / 16:
// 0 = K-major, 1 = MN-major), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with fp16 operands (format code 0) and fp32 accumulation
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ---- weight image: [KT] x { hi tile, lo tile }, each tile = 256 (n) rows x 32 (k) floats, K-major, 128B-swizzled ----
constexpr int TC_N = 256;            // output features per tile == hidden size
constexpr int TC_BK = 32;            // K per stage: one 128-byte swizzle row
constexpr int TC_BM = 128;           // rows per accumulator tile
constexpr int TC_B_TILE_FLOATS = TC_N * TC_BK;       // 8192 floats = 32 KB
constexpr int TC_A_TILE_FLOATS = TC_BM * TC_BK;      // 4096 floats = 16 KB

// B(k, n) = transposed ? W[k*ldw + n] : W[n*ldw + k], zero beyond K.  One thread per 16-byte chunk.
__global__ void tc_prep_weights_kernel(const float *__restrict__ W, int ldw, int transposed, int K, int KT,
                                       float *__restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (kt, n, c)
    if (idx >= KT * TC_N * 8) return;
    const int c = idx & 7, n = (idx >> 3) & (TC_N - 1), kt = idx >> 11;
    float4 x;
    float *xv = reinterpret_cast<float *>(&x);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k = kt * TC_BK + c * 4 + e;
        xv[e] = (k < K) ? (transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k]) : 0.f;
    }
    float4 hi, lo;
    split_tf32(x, hi, lo);
    float *tile = img + (size_t)kt * (2 * TC_B_TILE_FLOATS) + n * TC_BK + ((c ^ (n & 7)) << 2);
    *reinterpret_cast<float4 *>(tile) = hi;
    *reinterpret_cast<float4 *>(tile + TC_B_TILE_FLOATS) = lo;
}

// fp16-split image: [KT64] x { hi tile, lo tile }, each tile = 256 (n) rows x 64 (k) halves (the same 32 KB and the
// same 128-byte rows / swizzle as above, twice the reduction depth).  The weights are multiplied by `wscale` (a power of
// two, undone exactly in the epilogue) so that the lo halves stay in fp16's normal range; hi saturates at the largest
// finite fp16 instead of overflowing.
constexpr int TC_BK16 = 64;
__global__ void tc_prep_weights_f16_kernel(const float *__restrict__ W, int ldw, int transposed, int K, int KT,
                                           float wscale, float *__restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (kt, n, c)
    if (idx >= KT * TC_N * 8) return;
    const int c = idx & 7, n = (idx >> 3) & (TC_N - 1), kt = idx >> 11;
    float xv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kt * TC_BK16 + c * 8 + e;
        const float w = (k < K) ? (transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k]) : 0.f;
        xv[e] = fminf(fmaxf(w * wscale, -65504.f), 65504.f);
    }
    uint4 hi, lo;
    split_f16x8(make_float4(xv[0], xv[1], xv[2], xv[3]), make_float4(xv[4], xv[5], xv[6], xv[7]), hi, lo);
    float *tile = img + (size_t)kt * (2 * TC_B_TILE_FLOATS) + n * TC_BK + ((c ^ (n & 7)) << 2);
    *reinterpret_cast<uint4 *>(tile) = hi;
    *reinterpret_cast<uint4 *>(tile + TC_B_TILE_FLOATS) = lo;
}

// max |x| over a [rows, cols] fp32 matrix (cols % 4 == 0, 16-byte aligned rows) as float bits (atomicMax on the bits of
// non-negative floats orders them correctly); *out must be zeroed first
__global__ void tc_absmax_bits_kernel(const float *__restrict__ x, long rows, int cols, int ld, uint32_t *__restrict__ out) {
    const int c4n = cols >> 2;
    const long n = rows * c4n;
    float m = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / c4n;
        const int c4 = (int)(i - r * c4n);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * ld) + c4);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
// fp32 matrix [rows, cols] (row pitch ld floats) -> pre-split fp16 hi / lo matrices (row pitch ld16 halves, ld16 % 8 == 0,
// columns cols..ld16-1 zero).  Test hook / conversion helper for the TMA-fed operand paths (TcfParams::a_split).
__global__ void tc_split16_kernel(const float *__restrict__ x, long rows, int cols, int ld, __half *__restrict__ hi,
                                  __half *__restrict__ lo, int ld16) {
    const long n = rows * (ld16 >> 1);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (ld16 >> 1);
        const int c = (int)(i - r * (ld16 >> 1)) * 2;
        const float a = c < cols ? x[r * ld + c] : 0.f, b = c + 1 < cols ? x[r * ld + c + 1] : 0.f;
        uint32_t h, l;
        split_f16_pair(a, b, h, l);
        *reinterpret_cast<uint32_t *>(hi + r * ld16 + c) = h;
        *reinterpret_cast<uint32_t *>(lo + r * ld16 + c) = l;
    }
}

// power-of-two scale S with absmax * S in [2^13, 2^14) (1 when absmax is zero / denormal) and its inverse, from the bits
__device__ __forceinline__ void f16_scale_from_absmax(uint32_t bits, float &S, float &invS) {
    const uint32_t e = (bits >> 23) & 0xffu;
    if (e == 0) { S = 1.f; invS = 1.f; return; }
    uint32_t be = 267u - e;                 // biased exponent of 2^(13 - (e - 127))
    if (be > 254u) be = 254u;
    S = __uint_as_float(be << 23);
    invS = __uint_as_float((254u - be) << 23);
}


// ---- forward-shaped GEMM: C[M, 256] = A[M, K] * B^T, A row-major (K contiguous), B from a weight image ----------
//
// Accuracy: the tensor core adds into its TMEM accumulator with round-toward-zero, so one long accumulator chain
// drifts by ~0.1 ulp per MMA, always toward zero (measured rms 7e-9 * K relative — 2.4e-6 at K = 352, 1.9e-5 at
// K = 2720 — far from the 1.5e-7 of an fp32 FFMA chain).  The kernel therefore never lets a chain grow: each K-stage
// (32 elements = 12 MMAs) starts a FRESH accumulator (the two 256-column TMEM buffers ping-pong per stage), and the
// epilogue warps drain every finished stage into an fp32 register tile with round-to-nearest adds while the next
// stage's MMAs run.  Measured: rms 2.5e-7 relative, independent of K.
//
// Warp roles (512 threads, 4 warpgroups; registers re-split with setmaxnreg):
//   WG0 warps 0-3    producers: activation tile LDG -> hi/lo split -> swizzled STS; thread 0 issues the weight bulk copies
//   WG1 warps 4-7    accumulate / epilogue, columns   0-127 (TMEM lane quarter = warp % 4, one row per thread)
//   WG2 warps 8-11   accumulate / epilogue, columns 128-255
//   WG3 warp 12      MMA issuer (one thread) + TMEM allocation; warps 13-15 idle
constexpr int TCF_STAGES = 2;
#ifndef DCC_TCF_CHAIN
#define DCC_TCF_CHAIN 1
#endif
constexpr int TCF_CHAIN = DCC_TCF_CHAIN;   // K-stages accumulated in TMEM before the drain warps take the partial tile
constexpr int TCF_STAGE_BYTES = 2 * TC_A_TILE_FLOATS * 4 + 2 * TC_B_TILE_FLOATS * 4;   // 96 KB
constexpr int TCF_XPOSE_BYTES = 8 * 32 * 32 * 4;   // per-warp 32x32 store staging (XOR-swizzled columns: conflict-free)
constexpr int TCF_SMEM_BYTES = TCF_STAGES * TCF_STAGE_BYTES + TCF_XPOSE_BYTES + 1024 /*align*/ + 1280 /*barriers, row stats*/;

static_assert(TCF_SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int TCF_THREADS = 512;

enum { TCF_EPI_STORE = 0, TCF_EPI_BIAS_RELU_LN = 1 };

struct TcfParams {
    const float *A;      // [M, lda], lda % 4 == 0, 16-byte aligned
    const float *Bimg;   // [KT] x 64 KB
    float *C;            // [M, ldc]: raw product (EPI_STORE) or the post-ReLU activation a (EPI_BIAS_RELU_LN, may be NULL)
    int M, K, KT, lda, ldc;
    // split-K for load balance when there are few row tiles (partials combined with red.global.add; C zeroed first)
    int splits, kt_per_split;
    // fused epilogue (EPI_BIAS_RELU_LN): a = act(z + bias); h = LayerNorm(a) * gamma + beta (mlp.py:19-22)
    int epi;
    const float *bias, *gamma, *beta;
    float *H;            // [M, ldc]
    float *mean, *rstd;  // [M] (may be NULL)
    // fused output head of the LAST trunk block (EPI_BIAS_RELU_LN only): head_dst[row, o] = sum_c h[row, c] * Wh[o, c] + bh[o]
    // for o < head_out (2 = the actor's action mean, act.py:79-84 / distributions.py:83-92; 1 = the critic's value,
    // r_actor_critic.py:120).  With h = (a - mean) rstd gamma + beta this is
    //     rstd * (sum_c a_c gw[o,c] - mean * sgw[o]) + bw[o],   gw = gamma * Wh[o,:], sgw = sum_c gw, bw = beta . Wh[o,:] + bh[o]
    // (head_fold_kernel, once per call), so the thread that owns an accumulator row accumulates sum_c a_c gw[o,c] over its
    // 128 columns in the SAME pass that applies bias + activation, the partial sums ride along with the LayerNorm
    // statistics through shared memory (no extra barrier), and h itself need not leave the SM at all (H may then be NULL:
    // the backward pass rebuilds h from the saved activation a).
    // head_fold: [head_out][256] gw, then head_out x sgw, then head_out x bw (floats, 16-byte aligned)
    int head_out;
    const float *head_fold;
    float *head_dst;
    int dbg;             // tools/tc_bench only: 1 = skip the global stores of the epilogue
    // TMA store path of the epilogue (splits == 1): 2-D tensor maps over C and H ([M rows, 256 cols] fp32, box
    // 32 x 32, SWIZZLE_128B).  A thread owns one accumulator row, so it writes its row of a 32 x 32 box into the warp's
    // 4 KB staging buffer with conflict-free 16-byte stores (chunk ^ (row & 7), the layout the tensor map un-swizzles)
    // and ONE cp.async.bulk.tensor store per box leaves asynchronously: no shared-memory read-back, no st.global issue
    // or store-queue stalls in the epilogue warps, rows past M are clipped by the hardware.
    int use_tma;
    int act;             // trunk activation of the fused epilogue: 0 = ReLU, 1 = tanh (mappo.yaml use_ReLU)
    float out_scale;     // fp16-split kernel: 1 / wscale of the weight image, applied to the drained tile
    const uint32_t *a_absmax_bits;   // ASCALE variant: bits of max |A| (A is pre-scaled into fp16 range, see TcwParams)
    alignas(64) CUtensorMap tmC, tmH;
    // L2 prefetch of the activation operand: one cp.async.bulk.prefetch.tensor per K-stage, issued `pf_dist` stages ahead
    // of the producers' loads (box = 128 rows x one stage of columns), so that their LDGs hit L2 instead of waiting
    // ~2 k cycles for HBM with only 16 KB in flight per SM.  0 = off.
    int pf_dist;
    alignas(64) CUtensorMap tmA;
    // PRE-SPLIT operands (round 2).  An activation that only GEMMs consume — the inner trunk outputs h_k, the compact layer-1
    // features — is stored as TWO fp16 matrices (hi = fp16(x), lo = fp16(x - hi); 4 bytes per element like fp32) by the
    // kernel that produces it.  The consuming GEMMs then fetch their operand tiles with TMA tensor loads straight into
    // the swizzled shared-memory layout the tensor core reads (no producer warps, no conversion traffic through shared
    // memory); numerically identical to splitting in the consumer's producer warps.
    //   a_split: A comes from (tmAhi, tmAlo): [M, K] fp16, box 64 (k) x 128 rows, SWIZZLE_128B.  F16 kernel only.
    //   h_split: H leaves as (tmHhi, tmHlo): [M, 256] fp16, box 32 x 32, SWIZZLE_64B, staged hi | lo in the 4 KB buffer.
    int a_split, h_split;
    int unit_affine;     // the LayerNorm affine of this block is the identity (xhat mode): h = (a - mean) rstd, gamma / beta not read
    // optional reductions for the pre-split gradient path (float bits, atomicMax; zeroed by the caller): the largest rstd of the
    // rows this launch normalises (EPI_BIAS_RELU_LN) / the largest |C| it stores (EPI_STORE, after the scale) — the two factors of
    // the a-priori bound of |dz| that lets the LayerNorm-backward kernel write dz pre-scaled (relu_lnx_bwd_pipe_kernel<true, true>)
    uint32_t *rstd_max_out, *c_absmax_out;
    alignas(64) CUtensorMap tmAhi, tmAlo, tmHhi, tmHlo;
};

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no -lcuda); nullptr when unavailable
inline PFN_tensorMapEncodeTiled tc_tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
    }
    return enc;
}
// 2-D fp32 map: `cols` x `rows` elements, row pitch ld floats, box box_cols x box_rows.  Returns false when the encoder
// is unavailable or rejects the shape (callers then keep the st.global epilogue / skip the prefetch).
inline bool tc_make_map_2d(CUtensorMap *tm, const float *base, int cols, int rows, int ld, int box_cols, int box_rows,
                           CUtensorMapSwizzle swizzle) {
    PFN_tensorMapEncodeTiled enc = tc_tensor_map_encoder();
    if (!enc || rows < 1 || cols < 1 || (ld & 3) || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 2-D fp16 map: `cols` x `rows` elements, row pitch ld halves
inline bool tc_make_map_2d_f16(CUtensorMap *tm, const void *base, int cols, int rows, int ld, int box_cols, int box_rows,
                               CUtensorMapSwizzle swizzle) {
    PFN_tensorMapEncodeTiled enc = tc_tensor_map_encoder();
    if (!enc || rows < 1 || cols < 1 || (ld & 7) || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// TMA tensor load global -> shared (tile at column x, row y), completion counted in bytes on an mbarrier; elements outside the
// tensor are zero-filled
__device__ __forceinline__ void tma_load_2d(uint32_t sdst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sdst),
                 "l"((uint64_t)tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// [rows, 256] output matrix -> store map with a 128B-swizzled 32 x 32 box
inline bool tc_make_store_map(CUtensorMap *tm, const float *base, int rows, int ld) {
    return tc_make_map_2d(tm, base, 256, rows, ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
}
// [rows, K] activation matrix -> L2-prefetch map with a box of 128 rows x `box_cols` columns (one K-stage of a row tile)
inline bool tc_make_prefetch_map(CUtensorMap *tm, const float *base, int rows, int K, int ld, int box_cols) {
    return tc_make_map_2d(tm, base, K, rows, ld, box_cols, 128, CU_TENSOR_MAP_SWIZZLE_NONE);
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *tm, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)tm), "r"(x), "r"(y) : "memory");
}
// smem box (swizzled as the map says) -> global tile at (column x, row y), completion tracked by the bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t saddr, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tm), "r"(saddr),
                 "r"(x), "r"(y)
                 : "memory");
}

__device__ __forceinline__ void red_add_f32(float *addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// float4 held by lane `src` of the warp (compile-time lane): the epilogue keeps each per-column vector of its 128 columns as
// ONE coalesced float4 per lane (lane l = columns 4l..4l+3) and broadcasts element c4 with four shuffles instead of
// re-reading it from global memory in every unrolled iteration (32 broadcast loads per vector and tile, each an L2 round
// trip because the ~28 KB L1 left beside 227 KB of shared memory is flushed by the producers' tiles; measured upper bound of
// the gain with immediates: 10-12 % of the fused kernels, profiles/r02j_*).
__device__ __forceinline__ float4 lane_bcast4(const float4 &v, int src) {
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src),
                       __shfl_sync(0xffffffffu, v.w, src));
}

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// F16 = false: 3xTF32, 32 reduction elements per stage (p.KT = ceil(K / 32)).
// F16 = true:  fp16 hi/lo split, 64 reduction elements per stage in the SAME stage bytes, tile offsets, swizzle and 12
//              MMAs (kind::f16, K = 16 each) — half the shared-memory traffic and half the tensor time per reduction
//              element; p.KT = ceil(K / 64), p.Bimg from tc_prep_weights_f16_kernel.  |A| must stay below 65504
//              (LayerNorm outputs here).
// ASCALE (with F16, EXPERIMENTAL, DCC_TC_WGRAD_F16=1): A has an unbounded dynamic range (the backward dX = dZ W): it is
//              multiplied by the per-tensor power of two from p.a_absmax_bits before the split, undone in the epilogue.
template <bool F16, bool ASCALE = false>
__global__ void __launch_bounds__(TCF_THREADS, 1) tc_gemm_fwd_kernel(const __grid_constant__ TcfParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *xpose = reinterpret_cast<float *>(smem + TCF_STAGES * TCF_STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TCF_STAGES * TCF_STAGE_BYTES + TCF_XPOSE_BYTES);
    uint64_t *full = bars, *empty = bars + TCF_STAGES, *tfull = bars + 2 * TCF_STAGES, *tempty = bars + 2 * TCF_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TCF_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_work = ((p.M + TC_BM - 1) / TC_BM) * p.splits;   // work item = (row tile, K split)

    if (threadIdx.x == 0) {
        for (int s = 0; s < TCF_STAGES; ++s) {
            mbar_init(&full[s], p.a_split ? 1 : 4 + 1);   // 4 producer warps + the expect_tx arrival (pre-split A: TMA only)
            mbar_init(&empty[s], 1);      // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);      // tcgen05.commit
            mbar_init(&tempty[a], 8);     // 8 accumulate warps
        }
        fence_mbar_init();
    }
    if (warp == 12) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== producers =====
        setmaxnreg_dec<96>();
        const int t = threadIdx.x;            // 0..127
        const int c = t & 7, rsub = t >> 3;   // 16-byte chunk within the 128-byte row; row within a 16-row group
        uint32_t it = 0;
        TC_PROF_DECL(t0 = 0, t1 = 0, t2 = 0, p_wait = 0, p_work = 0);
        // The global loads of stage i+1 are issued before stage i is split and stored, so their latency overlaps the
        // wait for the shared-memory slot and the split of the previous stage (registers: two 8 x float4 sets).
        int w = blockIdx.x, kt = 0, kt1 = 0;
        auto set_work = [&](int ww) {
            kt = (ww % p.splits) * p.kt_per_split;
            kt1 = min(p.KT, kt + p.kt_per_split);
        };
        // prefetch cursor (thread 0): runs p.pf_dist stages ahead of the load cursor
        int wp = blockIdx.x, ktp = 0, ktp1 = 0;
        auto pf_set = [&]() {
            ktp = (wp % p.splits) * p.kt_per_split;
            ktp1 = min(p.KT, ktp + p.kt_per_split);
        };
        auto pf_step = [&]() {
            if (wp >= num_work) return;
            tma_prefetch_2d(&p.tmA, ktp * (F16 ? TC_BK16 : TC_BK), (wp / p.splits) * TC_BM);
            if (++ktp >= ktp1) {
                wp += gridDim.x;
                if (wp < num_work) pf_set();
            }
        };
        if (t == 0 && p.pf_dist > 0) {
            if (wp < num_work) pf_set();
            for (int i = 0; i < p.pf_dist; ++i) pf_step();
        }
        auto load_stage = [&](int ww, int kk, float4 (&v)[8]) {
            const int m0 = (ww / p.splits) * TC_BM, kcol = kk * TC_BK + c * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = m0 + i * 16 + rsub;
                v[i] = (row < p.M && kcol < p.K) ? ldg_stream(reinterpret_cast<const float4 *>(p.A + (size_t)row * p.lda + kcol))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (F16 && p.a_split) {
            // pre-split A: one thread feeds the pipeline — per stage two TMA tensor loads (A_hi, A_lo tiles, 16 KB each, rows
            // past M and columns past K zero-filled) and the two weight bulk copies, all counted on full[s]
            if (t == 0) {
                while (w < num_work) {
                    set_work(w);
                    const int m0 = (w / p.splits) * TC_BM;
                    for (; kt < kt1; ++kt, ++it) {
                        const int s = it % TCF_STAGES;
                        const uint32_t ph = (it / TCF_STAGES) & 1;
                        uint8_t *st = smem + s * TCF_STAGE_BYTES;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], 2 * TC_A_TILE_FLOATS * 4 + 2 * TC_B_TILE_FLOATS * 4);
                        tma_load_2d(smem_u32(st), &p.tmAhi, kt * TC_BK16, m0, &full[s]);
                        tma_load_2d(smem_u32(st) + TC_A_TILE_FLOATS * 4, &p.tmAlo, kt * TC_BK16, m0, &full[s]);
                        const float *src = p.Bimg + (size_t)kt * (2 * TC_B_TILE_FLOATS);
                        bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4, src, TC_B_TILE_FLOATS * 4, &full[s]);
                        bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4 + TC_B_TILE_FLOATS * 4, src + TC_B_TILE_FLOATS,
                                      TC_B_TILE_FLOATS * 4, &full[s]);
                    }
                    w += gridDim.x;
                }
            }
        } else if constexpr (F16) {
            // unit of work = HALF a stage (64 rows x 64 k = 8 float4 per thread), so that the register footprint of
            // the "loads of the next unit in flight while this one is split" scheme stays at two 8 x float4 sets
            auto load_unit = [&](int ww, int kk, int hf, float4 (&v)[8]) {
                const int m0 = (ww / p.splits) * TC_BM, kcol = kk * TC_BK16 + c * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = m0 + (hf * 4 + i) * 16 + rsub;
                    const float *src = p.A + (size_t)row * p.lda + kcol;
                    v[2 * i] = (row < p.M && kcol < p.K) ? ldg_stream(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[2 * i + 1] = (row < p.M && kcol + 4 < p.K) ? ldg_stream(reinterpret_cast<const float4 *>(src + 4))
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 vn[8];
            int hf = 0;
            float a_scale = 1.f;
            if constexpr (ASCALE) {
                float inv;
                f16_scale_from_absmax(__ldg(p.a_absmax_bits), a_scale, inv);
            }
            if (w < num_work) {
                set_work(w);
                load_unit(w, kt, 0, vn);
            }
            while (w < num_work) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = vn[i];
                const int kt_cur = kt, hf_cur = hf;
                if (hf == 0) {
                    hf = 1;
                } else {
                    hf = 0;
                    if (++kt >= kt1) {
                        w += gridDim.x;
                        if (w < num_work) set_work(w);
                    }
                }
                if (w < num_work) load_unit(w, kt, hf, vn);
                const int s = it % TCF_STAGES;
                const uint32_t ph = (it / TCF_STAGES) & 1;
                uint8_t *st = smem + s * TCF_STAGE_BYTES;
                const uint32_t st_u32 = smem_u32(st);
                TC_PROF_NOW(t0);
                if (hf_cur == 0) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (t == 0) {
                        if (p.pf_dist > 0) pf_step();
                        mbar_arrive_expect_tx(&full[s], 2 * TC_B_TILE_FLOATS * 4);
                        const float *src = p.Bimg + (size_t)kt_cur * (2 * TC_B_TILE_FLOATS);
                        bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4, src, TC_B_TILE_FLOATS * 4, &full[s]);
                        bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4 + TC_B_TILE_FLOATS * 4, src + TC_B_TILE_FLOATS,
                                      TC_B_TILE_FLOATS * 4, &full[s]);
                    }
                }
                TC_PROF_NOW(t1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = (hf_cur * 4 + i) * 16 + rsub;
                    uint4 hi, lo;
                    if constexpr (ASCALE) {
                        v[2 * i].x *= a_scale; v[2 * i].y *= a_scale; v[2 * i].z *= a_scale; v[2 * i].w *= a_scale;
                        v[2 * i + 1].x *= a_scale; v[2 * i + 1].y *= a_scale; v[2 * i + 1].z *= a_scale; v[2 * i + 1].w *= a_scale;
                    }
                    split_f16x8(v[2 * i], v[2 * i + 1], hi, lo);
                    const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                    sts128u(st_u32 + off, hi);
                    sts128u(st_u32 + TC_A_TILE_FLOATS * 4 + off, lo);
                }
                if (hf_cur == 1) {
                    fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    ++it;
                }
                TC_PROF_NOW(t2);
                TC_PROF_ADD(p_wait, t0, t1);
                TC_PROF_ADD(p_work, t1, t2);
            }
        } else {
            float4 vn[8];
            if (w < num_work) {
                set_work(w);
                load_stage(w, kt, vn);
            }
            while (w < num_work) {
                float4 v[8];
    #pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = vn[i];
                const int kt_cur = kt;
                // advance to the next stage of this CTA and start its loads
                if (++kt >= kt1) {
                    w += gridDim.x;
                    if (w < num_work) set_work(w);
                }
                if (w < num_work) load_stage(w, kt, vn);
                const int s = it % TCF_STAGES;
                const uint32_t ph = (it / TCF_STAGES) & 1;
                uint8_t *st = smem + s * TCF_STAGE_BYTES;
                const uint32_t st_u32 = smem_u32(st);
                TC_PROF_NOW(t0);
                mbar_wait(&empty[s], ph ^ 1);
                TC_PROF_NOW(t1);
                if (t == 0) {
                    if (p.pf_dist > 0) pf_step();
                    mbar_arrive_expect_tx(&full[s], 2 * TC_B_TILE_FLOATS * 4);
                    const float *src = p.Bimg + (size_t)kt_cur * (2 * TC_B_TILE_FLOATS);
                    bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4, src, TC_B_TILE_FLOATS * 4, &full[s]);
                    bulk_load_g2s(st + 2 * TC_A_TILE_FLOATS * 4 + TC_B_TILE_FLOATS * 4, src + TC_B_TILE_FLOATS,
                                  TC_B_TILE_FLOATS * 4, &full[s]);
                }
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 16 + rsub;
                    float4 hi, lo;
                    split_tf32(v[i], hi, lo);
                    const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                    sts128(st_u32 + off, hi);
                    sts128(st_u32 + TC_A_TILE_FLOATS * 4 + off, lo);
                }
                fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                TC_PROF_NOW(t2);
                TC_PROF_ADD(p_wait, t0, t1);
                TC_PROF_ADD(p_work, t1, t2);
                ++it;
            }
        }
        TC_PROF_OUT(t == 0, 0, p_wait);
        TC_PROF_OUT(t == 0, 1, p_work);
        TC_PROF_OUT(t == 0, 2, (unsigned long long)it);
    } else if (warp >= 12) {
        // ===== MMA issuer: one thread of warp 12 =====
        setmaxnreg_dec<32>();
        if (warp == 12 && lane == 0) {
            constexpr uint32_t idesc = F16 ? make_idesc_f16(TC_BM, TC_N, 0, 0) : make_idesc_tf32(TC_BM, TC_N, 0, 0);
            auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
                if constexpr (F16) tc_mma_f16(d, a, b, idesc, acc);
                else tc_mma_tf32(d, a, b, idesc, acc);
            };
            uint32_t it = 0, cit = 0;     // shared-memory stage counter, accumulator-chain counter
            TC_PROF_DECL(t0 = 0, t1 = 0, t2 = 0, t3 = 0, m_wacc = 0, m_wfull = 0, m_issue = 0, m_begin = 0);
            TC_PROF_NOW(m_begin);
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                const int kt0 = (w % p.splits) * p.kt_per_split, kt1 = min(p.KT, kt0 + p.kt_per_split);
                for (int kt = kt0; kt < kt1; ++kt, ++it) {
                    const int s = it & 1;                          // shared-memory stage
                    const uint32_t ph = (it >> 1) & 1;
                    const int j = kt - kt0;                        // stage within the tile
                    const bool chain_first = (j % TCF_CHAIN) == 0; // a chain = TCF_CHAIN stages into one accumulator
                    const bool chain_last = (j % TCF_CHAIN) == TCF_CHAIN - 1 || kt == kt1 - 1;
                    const int ab = cit & 1;                        // TMEM accumulator buffer of this chain
                    TC_PROF_NOW(t0);
                    if (chain_first) mbar_wait(&tempty[ab], ((cit >> 1) & 1) ^ 1);   // drained by the epilogue warps
                    TC_PROF_NOW(t1);
                    mbar_wait(&full[s], ph);           // operands landed
                    TC_PROF_NOW(t2);
                    tc_fence_after();
                    const uint32_t d = tmem_base + ab * TC_N;
                    const uint32_t sa = smem_u32(smem + s * TCF_STAGE_BYTES);
                    const uint64_t a_hi = make_desc_sw128(sa, 16, 1024);
                    const uint64_t a_lo = make_desc_sw128(sa + TC_A_TILE_FLOATS * 4, 16, 1024);
                    const uint64_t b_hi = make_desc_sw128(sa + 2 * TC_A_TILE_FLOATS * 4, 16, 1024);
                    const uint64_t b_lo = make_desc_sw128(sa + 2 * TC_A_TILE_FLOATS * 4 + TC_B_TILE_FLOATS * 4, 16, 1024);
                    // The 8 small cross terms go first, while the fresh accumulator is still ~2^-11 of its final
                    // magnitude (their round-toward-zero losses are then negligible); only the 4 hi*hi MMAs add at
                    // full magnitude.  Measured rms error vs float64: 1e-7-class instead of 2.5e-7 with interleaving.
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 8 tf32 / 16 fp16 = 32 bytes along the swizzled row
                        mma(d, a_lo + adv, b_hi + adv, (k != 0 || !chain_first) ? 1u : 0u);   // fresh accumulator per chain
                        mma(d, a_hi + adv, b_lo + adv, 1);
                    }
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        mma(d, a_hi + adv, b_hi + adv, 1);
                    }
                    tc_commit(&empty[s]);      // smem stage reusable once these MMAs have read it
                    if (chain_last) {
                        tc_commit(&tfull[ab]); // accumulator ready to drain
                        ++cit;
                    }
                    TC_PROF_NOW(t3);
                    TC_PROF_ADD(m_wacc, t0, t1);
                    TC_PROF_ADD(m_wfull, t1, t2);
                    TC_PROF_ADD(m_issue, t2, t3);
                }
            }
            TC_PROF_OUT(true, 4, m_wacc);
            TC_PROF_OUT(true, 5, m_wfull);
            TC_PROF_OUT(true, 6, m_issue);
            TC_PROF_OUT(true, 7, t3 - m_begin);
        }
        __syncwarp();
    } else {
        // ===== accumulate / epilogue warps 4-11 =====
        setmaxnreg_inc<192>();
        const int q = warp & 3;                 // TMEM lane quarter
        const int half = (warp - 4) >> 2;       // column half: 0 -> 0..127, 1 -> 128..255
        const int ew = warp - 4;
        const uint32_t xp_u32 = smem_u32(xpose + ew * (32 * 32));
        const uint32_t xq_u32 = smem_u32(xpose + (ew ^ 4) * (32 * 32));   // staging buffer of the warp that owns the other column half
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        uint32_t it = 0;
        TC_PROF_DECL(t0 = 0, t1 = 0, t2 = 0, t3 = 0, e_wait = 0, e_drain = 0, e_tile = 0, e_stats = 0);
        TC_PROF_DECL(t4 = 0, t5 = 0, t6 = 0, e_head = 0, e_wread = 0, e_lnbar = 0);
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int kt0 = (w % p.splits) * p.kt_per_split, kt1 = min(p.KT, kt0 + p.kt_per_split);
            float acc[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) acc[i] = 0.f;
            const int nchains = (kt1 - kt0 + TCF_CHAIN - 1) / TCF_CHAIN;
            for (int c = 0; c < nchains; ++c, ++it) {      // `it` counts accumulator chains here
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                TC_PROF_NOW(t0);
                mbar_wait(&tfull[s], ph);
                TC_PROF_NOW(t1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[32];
                    tmem_ld32(tmem_base + s * TC_N + half * 128 + j * 32 + lane_bits, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[s]);
                TC_PROF_NOW(t2);
                TC_PROF_ADD(e_wait, t0, t1);
                TC_PROF_ADD(e_drain, t1, t2);
            }
            TC_PROF_NOW(t0);
            // per-column vectors of this warp's 128 columns, one float4 per lane (see lane_bcast4); issued before the
            // scaling pass so their latency is covered
            float4 vbias = make_float4(0.f, 0.f, 0.f, 0.f), vg = vbias, vb = vbias, vw0 = vbias, vw1 = vbias;
            // The vectors of the statistics pass (bias, folded head weights) are then parked in this warp's own staging buffer
            // (bytes 512..2047; idle here: the previous tile's boxes have been read by the TMA engine long ago, the exchange
            // area is bytes 0..511) and every unrolled iteration fetches its four values with ONE broadcast LDS.128 instead of
            // four shuffles: the shuffle unit delivers one warp instruction per clock per SM, and 8 warps x 384 shuffles per
            // tile (bias + two head vectors) were ~3 k of the ~14 k cycles of a tile epilogue.
            const uint32_t vec_u32 = xp_u32 + 512;
            if (p.epi == TCF_EPI_BIAS_RELU_LN) {
                vbias = __ldg(reinterpret_cast<const float4 *>(p.bias + half * 128) + lane);
                if (p.H && !p.unit_affine) {
                    vg = __ldg(reinterpret_cast<const float4 *>(p.gamma + half * 128) + lane);
                    vb = __ldg(reinterpret_cast<const float4 *>(p.beta + half * 128) + lane);
                }
                if (p.head_out > 0) {
                    vw0 = __ldg(reinterpret_cast<const float4 *>(p.head_fold + half * 128) + lane);
                    vw1 = __ldg(reinterpret_cast<const float4 *>(p.head_fold + (p.head_out > 1 ? TC_N : 0) + half * 128) + lane);
                }
                if (p.use_tma) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                sts128(vec_u32 + lane * 16, vbias);
                if (p.head_out > 0) { sts128(vec_u32 + 512 + lane * 16, vw0); sts128(vec_u32 + 1024 + lane * 16, vw1); }
                __syncwarp();
            }
            if constexpr (F16) {
                float sc = p.out_scale;         // undo the power-of-two scale of the weight image (exact)
                if constexpr (ASCALE) {
                    float up, inv;
                    f16_scale_from_absmax(__ldg(p.a_absmax_bits), up, inv);
                    sc *= inv;
                }
#pragma unroll
                for (int i = 0; i < 128; ++i) acc[i] *= sc;
            }
            // ---- tile epilogue from registers: thread = row (q*32 + lane), 128 columns [half*128, +128) ----
            const int row0 = (w / p.splits) * TC_BM + q * 32;
            const int rl = q * 32 + lane;
            float mean = 0.f, rstd = 0.f;
            if (p.epi == TCF_EPI_BIAS_RELU_LN) {
                float sum = 0.f, d0 = 0.f, d1 = 0.f;
                const int hout = p.head_out;
                // Separate, compact unrolled loops per (activation, head) case: one merged loop with the tanh code inside
                // every unrolled iteration was 2.7x slower for ReLU (instruction-cache misses on each skipped block).
                if (p.act == 0 && hout == 0) {
#pragma unroll
                    for (int c4 = 0; c4 < 32; ++c4) {
                        const float4 bv = lds128(vec_u32 + c4 * 16);
                        acc[4 * c4 + 0] = fmaxf(acc[4 * c4 + 0] + bv.x, 0.f);
                        acc[4 * c4 + 1] = fmaxf(acc[4 * c4 + 1] + bv.y, 0.f);
                        acc[4 * c4 + 2] = fmaxf(acc[4 * c4 + 2] + bv.z, 0.f);
                        acc[4 * c4 + 3] = fmaxf(acc[4 * c4 + 3] + bv.w, 0.f);
                        sum += (acc[4 * c4 + 0] + acc[4 * c4 + 1]) + (acc[4 * c4 + 2] + acc[4 * c4 + 3]);
                    }
                } else if (p.act == 0) {
                    // (one output: the second dot product repeats the first and is ignored)
#pragma unroll
                    for (int c4 = 0; c4 < 32; ++c4) {
                        const float4 bv = lds128(vec_u32 + c4 * 16), w0 = lds128(vec_u32 + 512 + c4 * 16), w1 = lds128(vec_u32 + 1024 + c4 * 16);
                        acc[4 * c4 + 0] = fmaxf(acc[4 * c4 + 0] + bv.x, 0.f);
                        acc[4 * c4 + 1] = fmaxf(acc[4 * c4 + 1] + bv.y, 0.f);
                        acc[4 * c4 + 2] = fmaxf(acc[4 * c4 + 2] + bv.z, 0.f);
                        acc[4 * c4 + 3] = fmaxf(acc[4 * c4 + 3] + bv.w, 0.f);
                        sum += (acc[4 * c4 + 0] + acc[4 * c4 + 1]) + (acc[4 * c4 + 2] + acc[4 * c4 + 3]);
                        d0 = fmaf(acc[4 * c4 + 0], w0.x, d0); d0 = fmaf(acc[4 * c4 + 1], w0.y, d0);
                        d0 = fmaf(acc[4 * c4 + 2], w0.z, d0); d0 = fmaf(acc[4 * c4 + 3], w0.w, d0);
                        d1 = fmaf(acc[4 * c4 + 0], w1.x, d1); d1 = fmaf(acc[4 * c4 + 1], w1.y, d1);
                        d1 = fmaf(acc[4 * c4 + 2], w1.z, d1); d1 = fmaf(acc[4 * c4 + 3], w1.w, d1);
                    }
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 32; ++c4) {
                        const float4 bv = lds128(vec_u32 + c4 * 16);
                        acc[4 * c4 + 0] = tanhf(acc[4 * c4 + 0] + bv.x);
                        acc[4 * c4 + 1] = tanhf(acc[4 * c4 + 1] + bv.y);
                        acc[4 * c4 + 2] = tanhf(acc[4 * c4 + 2] + bv.z);
                        acc[4 * c4 + 3] = tanhf(acc[4 * c4 + 3] + bv.w);
                        sum += (acc[4 * c4 + 0] + acc[4 * c4 + 1]) + (acc[4 * c4 + 2] + acc[4 * c4 + 3]);
                    }
                    if (hout > 0) {
#pragma unroll
                        for (int c4 = 0; c4 < 32; ++c4) {
                            const float4 w0 = lds128(vec_u32 + 512 + c4 * 16), w1 = lds128(vec_u32 + 1024 + c4 * 16);
                            d0 = fmaf(acc[4 * c4 + 0], w0.x, d0); d0 = fmaf(acc[4 * c4 + 1], w0.y, d0);
                            d0 = fmaf(acc[4 * c4 + 2], w0.z, d0); d0 = fmaf(acc[4 * c4 + 3], w0.w, d0);
                            d1 = fmaf(acc[4 * c4 + 0], w1.x, d1); d1 = fmaf(acc[4 * c4 + 1], w1.y, d1);
                            d1 = fmaf(acc[4 * c4 + 2], w1.z, d1); d1 = fmaf(acc[4 * c4 + 3], w1.w, d1);
                        }
                    }
                }
                // exchange with the partner warp that owns the other 128 columns of the same 32 rows (warp w <-> w + 4): a
                // 64-thread named barrier per row quarter instead of one over all 256 epilogue threads; the head's partial
                // dot products ride along with the row sum (rowstat: [4 values][2 halves][128 rows])
                // The exchange area is the first 512 bytes of each warp's own staging buffer ([4 values][32 lanes]; idle here once
                // the previous tile's boxes have been read by the TMA engine).
                TC_PROF_NOW(t4);
                if (p.use_tma) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                const uint32_t mine = xp_u32 + lane * 4, other = xq_u32 + lane * 4;
                sts32(mine, sum);
                if (hout > 0) { sts32(mine + 128, d0); if (hout > 1) sts32(mine + 256, d1); }
                named_bar_sync(1 + q, 64);
                mean = (sum + lds32(other)) * (1.f / TC_N);
                float p0 = 0.f, p1 = 0.f;
                if (hout > 0) { p0 = d0 + lds32(other + 128); if (hout > 1) p1 = d1 + lds32(other + 256); }
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < 128; ++i) { const float d = acc[i] - mean; sq = fmaf(d, d, sq); }
                sts32(mine + 384, sq);
                named_bar_sync(1 + q, 64);
                rstd = rsqrtf((sq + lds32(other + 384)) * (1.f / TC_N) + 1e-5f);
                named_bar_sync(1 + q, 64);      // the partner has read my values: the staging buffer may be reused for the stores
                TC_PROF_NOW(t5);
                TC_PROF_ADD(e_lnbar, t4, t5);
                if (p.rstd_max_out && half == 0) {
                    float rm = (row0 + lane < p.M) ? rstd : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
                    if (lane == 0 && rm > 0.f) atomicMax(p.rstd_max_out, __float_as_uint(rm));
                }
                if (half == 0 && row0 + lane < p.M) {
                    if (p.mean) p.mean[row0 + lane] = mean;
                    if (p.rstd) p.rstd[row0 + lane] = rstd;
                    if (hout > 0) {
                        const float *tailv = p.head_fold + hout * TC_N;      // sgw[hout], bw[hout]
                        float *dst = p.head_dst + (size_t)(row0 + lane) * hout;
                        dst[0] = fmaf(rstd, p0 - mean * __ldg(tailv), __ldg(tailv + hout));
                        if (hout > 1) dst[1] = fmaf(rstd, p1 - mean * __ldg(tailv + 1), __ldg(tailv + hout + 1));
                    }
                }
            }
            if (p.c_absmax_out && p.epi == TCF_EPI_STORE) {      // rows past M hold exact zeros (zero-filled operand rows)
                float cm = 0.f;
#pragma unroll
                for (int i = 0; i < 128; ++i) cm = fmaxf(cm, fabsf(acc[i]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                if (lane == 0 && cm > 0.f) atomicMax(p.c_absmax_out, __float_as_uint(cm));
            }
            TC_PROF_NOW(t3);
            TC_PROF_ADD(e_stats, t0, t3);
            if (p.dbg & 2) {
                // direct stores: thread = row, 16 bytes per instruction (no shared-memory staging)
                const int row = row0 + lane;
                if (row < p.M && !(p.dbg & 1)) {
                    if (p.epi == TCF_EPI_STORE) {
                        float *dst = p.C + (size_t)row * p.ldc + half * 128;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (p.splits > 1) red_add_v4(dst + 4 * c, acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                            else *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                        }
                    } else {
                        if (p.C) {
                            float *dst = p.C + (size_t)row * p.ldc + half * 128;
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                        }
                        float *dsth = p.H + (size_t)row * p.ldc + half * 128;
#pragma unroll
                        for (int c = 0; p.H && c < 32; ++c) {
                            const float4 g = __ldg(reinterpret_cast<const float4 *>(p.gamma + half * 128) + c);
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.beta + half * 128) + c);
                            float4 hv;
                            hv.x = fmaf((acc[4 * c] - mean) * rstd, g.x, b.x); hv.y = fmaf((acc[4 * c + 1] - mean) * rstd, g.y, b.y);
                            hv.z = fmaf((acc[4 * c + 2] - mean) * rstd, g.z, b.z); hv.w = fmaf((acc[4 * c + 3] - mean) * rstd, g.w, b.w);
                            *reinterpret_cast<float4 *>(dsth + 4 * c) = hv;
                        }
                    }
                }
            } else if (p.use_tma) {
                // TMA store path: per 32-column block, stage this warp's 32 rows x 32 columns (thread = row) and hand the
                // box to the TMA engine; the staging buffer is reused once the engine has READ it (wait_group.read).
                // (Two half-size boxes per buffer were measured: the wait disappears but the 8 extra fence + issue sequences
                // per tile cost more, 15.7 k -> 16.7 k cycles per tile epilogue; profiles/r02i_*.)
                const bool ln = p.epi == TCF_EPI_BIAS_RELU_LN;
                const uint32_t myrow = xp_u32 + lane * 128;
                const int sw = lane & 7;
                if (row0 < p.M && !(p.dbg & 1)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col0 = half * 128 + j * 32;
                        if (p.C) {
                            TC_PROF_NOW(t4);
                            if (lane == 0) bulk_wait_read<0>();
                            __syncwarp();
                            TC_PROF_NOW(t5);
                            TC_PROF_ADD(e_wread, t4, t5);
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                sts128(myrow + ((c ^ sw) << 4), make_float4(acc[j * 32 + 4 * c], acc[j * 32 + 4 * c + 1],
                                                                            acc[j * 32 + 4 * c + 2], acc[j * 32 + 4 * c + 3]));
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(&p.tmC, xp_u32, col0, row0); bulk_commit(); }
                        }
                        if (ln && p.H && p.h_split) {
                            // pre-split output: h leaves as fp16 hi | lo (2 KB boxes side by side in the staging buffer,
                            // 64-byte rows in the tensor maps' SWIZZLE_64B layout: chunk c of row r at r*64 + ((c ^ ((r>>1)&3)) << 4))
                            TC_PROF_NOW(t4);
                            if (lane == 0) bulk_wait_read<0>();
                            __syncwarp();
                            TC_PROF_NOW(t5);
                            TC_PROF_ADD(e_wread, t4, t5);
                            const uint32_t r64 = xp_u32 + lane * 64;
                            const int sw2 = (lane >> 1) & 3;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int e0 = j * 32 + 8 * c;
                                float4 h0, h1;
                                if (p.unit_affine) {      // xhat mode: the un-affined normalisation itself (same values as gamma = 1, beta = 0)
                                    h0.x = (acc[e0 + 0] - mean) * rstd; h0.y = (acc[e0 + 1] - mean) * rstd;
                                    h0.z = (acc[e0 + 2] - mean) * rstd; h0.w = (acc[e0 + 3] - mean) * rstd;
                                    h1.x = (acc[e0 + 4] - mean) * rstd; h1.y = (acc[e0 + 5] - mean) * rstd;
                                    h1.z = (acc[e0 + 6] - mean) * rstd; h1.w = (acc[e0 + 7] - mean) * rstd;
                                } else {
                                    const float4 g0 = lane_bcast4(vg, j * 8 + 2 * c), b0 = lane_bcast4(vb, j * 8 + 2 * c);
                                    const float4 g1 = lane_bcast4(vg, j * 8 + 2 * c + 1), b1 = lane_bcast4(vb, j * 8 + 2 * c + 1);
                                    h0.x = fmaf((acc[e0 + 0] - mean) * rstd, g0.x, b0.x); h0.y = fmaf((acc[e0 + 1] - mean) * rstd, g0.y, b0.y);
                                    h0.z = fmaf((acc[e0 + 2] - mean) * rstd, g0.z, b0.z); h0.w = fmaf((acc[e0 + 3] - mean) * rstd, g0.w, b0.w);
                                    h1.x = fmaf((acc[e0 + 4] - mean) * rstd, g1.x, b1.x); h1.y = fmaf((acc[e0 + 5] - mean) * rstd, g1.y, b1.y);
                                    h1.z = fmaf((acc[e0 + 6] - mean) * rstd, g1.z, b1.z); h1.w = fmaf((acc[e0 + 7] - mean) * rstd, g1.w, b1.w);
                                }
                                uint4 hi, lo;
                                split_f16x8(h0, h1, hi, lo);
                                const uint32_t off = (uint32_t)((c ^ sw2) << 4);
                                sts128u(r64 + off, hi);
                                sts128u(r64 + 2048 + off, lo);
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&p.tmHhi, xp_u32, col0, row0);
                                tma_store_2d(&p.tmHlo, xp_u32 + 2048, col0, row0);
                                bulk_commit();
                            }
                        } else if (ln && p.H) {
                            TC_PROF_NOW(t4);
                            if (lane == 0) bulk_wait_read<0>();
                            __syncwarp();
                            TC_PROF_NOW(t5);
                            TC_PROF_ADD(e_wread, t4, t5);
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 g = lane_bcast4(vg, j * 8 + c), b = lane_bcast4(vb, j * 8 + c);
                                float4 hv;
                                hv.x = fmaf((acc[j * 32 + 4 * c] - mean) * rstd, g.x, b.x);
                                hv.y = fmaf((acc[j * 32 + 4 * c + 1] - mean) * rstd, g.y, b.y);
                                hv.z = fmaf((acc[j * 32 + 4 * c + 2] - mean) * rstd, g.z, b.z);
                                hv.w = fmaf((acc[j * 32 + 4 * c + 3] - mean) * rstd, g.w, b.w);
                                sts128(myrow + ((c ^ sw) << 4), hv);
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(&p.tmH, xp_u32, col0, row0); bulk_commit(); }
                        }
                    }
                }
            } else {
            // Stores: a thread holds 128 columns of ONE row, so a direct store would scatter 16-byte pieces over 32 rows.
            // Rows are staged 8 at a time through the warp's 4 KB buffer (16-byte chunks XOR-ed with the row: conflict
            // free both ways) and leave as whole 512-byte row segments, one STG.128 per row per warp.
            const int colw = half * 128 + lane * 4;     // this lane's 4 columns on the way out
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
            if (p.epi == TCF_EPI_BIAS_RELU_LN) {
                g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma + colw));
                b4 = __ldg(reinterpret_cast<const float4 *>(p.beta + colw));
            }
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                const int rr = lane & 7;
                if ((lane >> 3) == g) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        sts128(xp_u32 + ((rr * 32 + (c ^ rr)) << 4), make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int row = row0 + g * 8 + r;
                    if (row < p.M && !(p.dbg & 1)) {
                        const float4 v = lds128(xp_u32 + ((r * 32 + (lane ^ r)) << 4));
                        if (p.epi == TCF_EPI_STORE) {
                            float *dst = p.C + (size_t)row * p.ldc + colw;
                            if (p.splits > 1) red_add_v4(dst, v.x, v.y, v.z, v.w);
                            else *reinterpret_cast<float4 *>(dst) = v;
                        } else {
                            if (p.C) *reinterpret_cast<float4 *>(p.C + (size_t)row * p.ldc + colw) = v;
                            // the row statistics of the 8 staged rows sit in lanes g*8 + r
                            const float mr = __shfl_sync(0xffffffffu, mean, g * 8 + r), rs = __shfl_sync(0xffffffffu, rstd, g * 8 + r);
                            float4 hv;
                            hv.x = fmaf((v.x - mr) * rs, g4.x, b4.x); hv.y = fmaf((v.y - mr) * rs, g4.y, b4.y);
                            hv.z = fmaf((v.z - mr) * rs, g4.z, b4.z); hv.w = fmaf((v.w - mr) * rs, g4.w, b4.w);
                            if (p.H) *reinterpret_cast<float4 *>(p.H + (size_t)row * p.ldc + colw) = hv;
                        }
                    }
                }
                __syncwarp();
            }
            }
            TC_PROF_NOW(t1);
            TC_PROF_ADD(e_tile, t0, t1);
        }
        if (p.use_tma && lane == 0) bulk_wait_all<0>();   // staged boxes fully written before the CTA's smem goes away
        __syncwarp();
        TC_PROF_OUT(threadIdx.x == 128, 8, e_wait);
        TC_PROF_OUT(threadIdx.x == 128, 9, e_drain);
        TC_PROF_OUT(threadIdx.x == 128, 10, e_tile);
        TC_PROF_OUT(threadIdx.x == 128, 11, e_stats);
        TC_PROF_OUT(threadIdx.x == 128, 12, e_head);
        TC_PROF_OUT(threadIdx.x == 128, 13, e_wread);
        TC_PROF_OUT(threadIdx.x == 128, 14, e_lnbar);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- weight-gradient GEMM: G[256, Nout] += dZ[R, 256]^T X[R, Nout]  (reduction over the R batch rows) -----------
//
// Both operands are activations stored row-major [row][feature], i.e. MN-major for this product (the reduction index
// is the slow one).  tcgen05 reads MN-major operands natively (instruction-descriptor major bits = 1) from the
// canonical MN-major layout for 32-bit types (SWIZZLE_128B_BASE32B): per 32-feature group a block of [k rows][128 B],
// 4-row atoms, 32-byte chunks XOR-ed with (k & 3); groups LBO apart.  A coalesced 128-byte global row segment is exactly one shared-memory
// row, so the producers need no transpose.
// A CTA owns a 128 (dZ features) x <=256 (X features) output tile and a contiguous range of batch rows; work items
// that share a row range are adjacent in the grid so the second reader of a dZ / X tile hits L2.  Per 32-row stage the
// accumulator chain is fresh and drained into fp32 registers (see tc_gemm_fwd_kernel); the CTA's partial tile is
// added to G with red.global.add.f32 (coalesced through the per-warp transpose).
struct TcwParams {
    const float *dZ;     // [R, ldz], 256 features
    const float *X;      // [R, ldx], Nout valid features (ldx % 4 == 0)
    float *G;            // [256, ldg]
    int R, Nout, ldz, ldx, ldg;
    int n_tiles;         // column tiles; output tiles = 2 * n_tiles
    int tile_n;          // columns per tile: a multiple of 32, <= 256, chosen to balance the tiles (338 -> 192 + 146)
    int ksplits, rows_per_split;   // rows_per_split % 32 == 0 (% 64 for the fp16-split kernel)
    // fp16-split kernel only: bits of max |dZ| over the tensor (tc_absmax_bits_kernel); dZ is multiplied by the power of
    // two that brings this maximum into [2^13, 2^14) before the split and the partial tile is scaled back before it is added
    const uint32_t *dz_absmax_bits;
    // fp16-split kernel, pre-split X (see TcfParams): the X operand comes from (tmXhi, tmXlo) — [R, Nout] fp16, box 64 features
    // x 64 rows, SWIZZLE_128B, which IS the MN-major 16-bit operand layout (64-feature groups of [64 k rows][128 B], 16-byte
    // chunks XOR (k & 7)) — fetched by TMA tensor loads; the producer warps then convert only dZ (2 of the 6 units of a stage).
    int x_split;
    // dz_split (with x_split): dZ is pre-split as well — two fp16 matrices [R, 256] already multiplied by the power of two that
    // f16_scale_from_absmax derives from *dz_absmax_bits (an upper BOUND of |dZ| known before dZ is written, see dcc_mappo.cu) —
    // and arrives as four 64 x 64 boxes per stage (tmZhi, tmZlo); no producer warp touches the data: one thread feeds the pipeline.
    int dz_split;
    alignas(64) CUtensorMap tmXhi, tmXlo, tmZhi, tmZlo;
};

//
// Round-2 note: a variant with EIGHT producer warps (640 threads, setmaxnreg 72 / 152 / 24) was measured — no gain in isolation
// (236 vs 235 us) and slower inside the update (dW2 227 -> 255 us, profiles/r02l_*): the producers are not issue-bound by
// their warp count; a 32-row stage moves ~336 KB through shared memory (cp.async ring in + out, hi/lo stores, tensor-core
// reads), i.e. ~2 600 cycles at 128 B/clk against 1 536 cycles of MMA.
// F16 = true (off by default: DCC_TC_WGRAD_F16=1; validated, not faster): fp16 hi/lo split of both operands, 64 batch rows per
// stage in the same stage bytes.  16-bit MN-major operands use the ordinary SWIZZLE_128B layout: per 64-feature group a
// block of [64 k rows][128 B], 16-byte chunks XOR-ed with (k & 7), groups 8 KB apart (descriptor LBO = 8192, SBO = 1024,
// 2048 B per K = 16 instruction: pinned on the GPU with tools/mn16_probe.cu).  dZ is pre-scaled by one power of two per
// tensor (TcwParams::dz_absmax_bits); X must be a LayerNorm output.
template <bool F16>
__global__ void __launch_bounds__(TCF_THREADS, 1) tc_gemm_wgrad_kernel(const __grid_constant__ TcwParams p) {
    constexpr int BKW = F16 ? 64 : TC_BK;          // batch rows per stage
    const bool xs = F16 && p.x_split;              // X operand pre-split in global memory, fetched by TMA
    const int UNITS = F16 ? (xs ? 2 : 6) : 3;      // 16 KB raw load units per stage: dZ 1 (2), X 2 (4; 0 when pre-split)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *xpose = reinterpret_cast<float *>(smem + TCF_STAGES * TCF_STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TCF_STAGES * TCF_STAGE_BYTES + TCF_XPOSE_BYTES);
    uint64_t *full = bars, *empty = bars + TCF_STAGES, *tfull = bars + 2 * TCF_STAGES, *tempty = bars + 2 * TCF_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TCF_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int out_tiles = 2 * p.n_tiles;
    const int num_work = out_tiles * p.ksplits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TCF_STAGES; ++s) {
            // 4 producer warps (+ the expect_tx arrival of the X tiles); with dZ pre-split too, only the expect_tx arrival
            mbar_init(&full[s], (F16 && p.x_split) ? (p.dz_split ? 1 : 5) : 4);
            mbar_init(&empty[s], 1);      // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 8);
        }
        fence_mbar_init();
    }
    if (warp == 12) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // smem stage layout: A_hi 16 KB | A_lo 16 KB | B_hi 32 KB | B_lo 32 KB; every 32-feature group = 32 rows x 128 B
    if (warp < 4 && F16 && p.x_split && p.dz_split) {
        // both operands pre-split in global memory: one thread issues the stage's TMA loads, nothing is converted here
        setmaxnreg_dec<96>();
        if (threadIdx.x == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                const int ot = w % out_tiles, ks = w / out_tiles;
                const int mh = ot & 1, n0 = (ot >> 1) * p.tile_n;
                const int nv = min(p.tile_n, (p.Nout - n0 + 15) & ~15);
                const int ngroups = (nv + 63) >> 6;
                const int r_beg = ks * p.rows_per_split, r_end = min(p.R, r_beg + p.rows_per_split);
                for (int r0 = r_beg; r0 < r_end; r0 += BKW, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const uint32_t st_u32 = smem_u32(smem + s * TCF_STAGE_BYTES);
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(4 + 2 * ngroups) * 8192u);
                    for (int g = 0; g < 2; ++g) {      // dZ features [mh * 128 + 64 g, +64), rows [r0, r0 + 64); rows past R are zero-filled
                        tma_load_2d(st_u32 + g * 8192, &p.tmZhi, mh * 128 + g * 64, r0, &full[s]);
                        tma_load_2d(st_u32 + TC_A_TILE_FLOATS * 4 + g * 8192, &p.tmZlo, mh * 128 + g * 64, r0, &full[s]);
                    }
                    const uint32_t sb = st_u32 + 2 * TC_A_TILE_FLOATS * 4;
                    for (int g = 0; g < ngroups; ++g) {
                        tma_load_2d(sb + g * 8192, &p.tmXhi, n0 + g * 64, r0, &full[s]);
                        tma_load_2d(sb + TC_B_TILE_FLOATS * 4 + g * 8192, &p.tmXlo, n0 + g * 64, r0, &full[s]);
                    }
                }
            }
        }
    } else if (warp < 4) {
        setmaxnreg_dec<96>();
        const int t = threadIdx.x;
        // Work is cut into load units of 8 x 16 bytes per thread: per 32-row stage one unit of dZ (A) and two units of X
        // (B).  Raw units are fetched with 16-byte cp.async (LDGSTS) into a 2-slot ring in shared memory two units
        // ahead of their use, so the global-load latency (~1 us) overlaps the split / store work of two units and the
        // wait for the stage slot; each thread reads back only the chunks it fetched itself (no barrier needed).
        const uint32_t ring_u32 = smem_u32(xpose);     // 2 slots x 16 KB (the epilogue of this kernel uses no staging)
        int w = blockIdx.x, r0 = 0, r_end = 0, mh = 0, n0 = 0, ngroups = 0;
        auto set_work = [&](int ww) {
            const int ot = ww % out_tiles, ks = ww / out_tiles;
            mh = ot & 1; n0 = (ot >> 1) * p.tile_n;
            const int nv = min(p.tile_n, (p.Nout - n0 + 15) & ~15);
            ngroups = F16 ? (nv + 63) >> 6 : (nv + 31) >> 5;
            r0 = ks * p.rows_per_split; r_end = min(p.R, r0 + p.rows_per_split);
        };
        // ring offset of a thread's i-th 16-byte piece; in the fp16 form pieces 2j, 2j+1 are the two halves of one
        // 8-feature group (32 contiguous bytes in global memory) that becomes ONE 16-byte fp16 chunk
        auto ring_off = [](int i) { return F16 ? (uint32_t)((i & 1) * 8192 + (i >> 1) * 2048) : (uint32_t)(i * 2048); };
        float dz_scale = 1.f;
        if constexpr (F16) {
            float inv;
            f16_scale_from_absmax(__ldg(p.dz_absmax_bits), dz_scale, inv);
        }
        auto next_nonempty = [&]() {
            while (w < num_work) {
                set_work(w);
                if (r0 < r_end) return;
                w += gridDim.x;
            }
        };
        // fetch cursor (runs two units ahead of the consume cursor)
        int f_kind = 0;
        auto fetch_unit = [&](int slot) {      // issues the cp.asyncs of the unit at the fetch cursor, then advances it
            if (w < num_work) {
                const uint32_t dst = ring_u32 + slot * 16384 + t * 16;
                if constexpr (F16) {
                    if (f_kind < 2) {          // dZ rows [32 * f_kind, +32) of the stage: 16 chunks of 8 features per row
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int pi = t + 128 * (i >> 1), k = f_kind * 32 + (pi >> 4), c = pi & 15;
                            const bool ok = r0 + k < r_end;
                            cp_async16(dst + ring_off(i), p.dZ + (size_t)(ok ? r0 + k : 0) * p.ldz + mh * 128 + c * 8 + (i & 1) * 4,
                                       ok ? 16u : 0u);
                        }
                    } else {                   // X rows [16 * (f_kind - 2), +16): 32 chunks of 8 features per row
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int pi = t + 128 * (i >> 1), k = (f_kind - 2) * 16 + (pi >> 5), c = pi & 31;
                            const int col = n0 + c * 8 + (i & 1) * 4;
                            const bool ok = (c >> 3) < ngroups && r0 + k < r_end && col < p.Nout;
                            cp_async16(dst + ring_off(i), p.X + (size_t)(ok ? r0 + k : 0) * p.ldx + (ok ? col : 0), ok ? 16u : 0u);
                        }
                    }
                } else if (f_kind == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ci = t + 128 * i, k = ci >> 5, mc = ci & 31;
                        const bool ok = r0 + k < r_end;
                        cp_async16(dst + i * 2048, p.dZ + (size_t)(ok ? r0 + k : 0) * p.ldz + mh * 128 + mc * 4, ok ? 16u : 0u);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ci = t + 128 * ((f_kind - 1) * 8 + i), k = ci >> 6, nc = ci & 63;
                        const int col = n0 + nc * 4;
                        const bool ok = (nc >> 3) < ngroups && r0 + k < r_end && col < p.Nout;
                        cp_async16(dst + i * 2048, p.X + (size_t)(ok ? r0 + k : 0) * p.ldx + (ok ? col : 0), ok ? 16u : 0u);
                    }
                }
                if (++f_kind == UNITS) {
                    f_kind = 0;
                    r0 += BKW;
                    if (r0 >= r_end) {
                        w += gridDim.x;
                        next_nonempty();
                    }
                }
            }
            cp_async_commit();     // one group per unit, also when empty: keeps the wait_group accounting uniform
        };
        next_nonempty();
        // consume cursor: the same unit sequence, replayed from a copy of the initial cursor
        int cw = w, c_r0 = r0, c_rend = r_end, c_ngroups = ngroups, c_kind = 0;
        auto c_set_work = [&](int ww) {
            const int ot = ww % out_tiles, ks = ww / out_tiles;
            const int nn0 = (ot >> 1) * p.tile_n;
            const int nv = min(p.tile_n, (p.Nout - nn0 + 15) & ~15);
            c_ngroups = F16 ? (nv + 63) >> 6 : (nv + 31) >> 5;
            c_r0 = ks * p.rows_per_split; c_rend = min(p.R, c_r0 + p.rows_per_split);
        };
        fetch_unit(0);
        fetch_unit(1);
        uint32_t it = 0, u = 0;
        TC_PROF_DECL(p_wait_acc = 0, p_cp_acc = 0, pc0 = 0, pc1 = 0, p_begin = 0, p_end = 0);
        TC_PROF_NOW(p_begin);
        while (cw < num_work) {
            const int slot = u & 1;
            TC_PROF_NOW(pc0);
            cp_async_wait<1>();                  // unit u has landed (unit u+1 may still be in flight)
            TC_PROF_NOW(pc1);
            TC_PROF_ADD(p_cp_acc, pc0, pc1);
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = lds128(ring_u32 + slot * 16384 + t * 16 + ring_off(i));
            fetch_unit(slot);                    // refill this slot with unit u+2
            const int cur_kind = c_kind, cur_ngroups = c_ngroups;
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const uint32_t st_u32 = smem_u32(smem + s * TCF_STAGE_BYTES);
            if (cur_kind == 0) {
                TC_PROF_DECL(pw0 = 0, pw1 = 0);
                TC_PROF_NOW(pw0);
                mbar_wait(&empty[s], ph ^ 1);
                TC_PROF_NOW(pw1);
                TC_PROF_ADD(p_wait_acc, pw0, pw1);
                if (xs && t == 0) {
                    // the stage's X tiles: one 64 x 64 box per 64-feature group and half (hi, lo), 8 KB each
                    const int ot = cw % out_tiles;
                    const int xn0 = (ot >> 1) * p.tile_n;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)cur_ngroups * 2u * 8192u);
                    const uint32_t sb = st_u32 + 2 * TC_A_TILE_FLOATS * 4;
                    for (int g = 0; g < cur_ngroups; ++g) {
                        tma_load_2d(sb + g * 8192, &p.tmXhi, xn0 + g * 64, c_r0, &full[s]);
                        tma_load_2d(sb + TC_B_TILE_FLOATS * 4 + g * 8192, &p.tmXlo, xn0 + g * 64, c_r0, &full[s]);
                    }
                }
            }
            if constexpr (F16) {
                if (cur_kind < 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pi = t + 128 * j, k = cur_kind * 32 + (pi >> 4), c = pi & 15;
                        float4 a = v[2 * j], b = v[2 * j + 1];
                        a.x *= dz_scale; a.y *= dz_scale; a.z *= dz_scale; a.w *= dz_scale;
                        b.x *= dz_scale; b.y *= dz_scale; b.z *= dz_scale; b.w *= dz_scale;
                        uint4 hi, lo;
                        split_f16x8(a, b, hi, lo);
                        const uint32_t off = (uint32_t)((c >> 3) * 8192 + k * 128 + (((c & 7) ^ (k & 7)) << 4));
                        sts128u(st_u32 + off, hi);
                        sts128u(st_u32 + TC_A_TILE_FLOATS * 4 + off, lo);
                    }
                    if (xs && cur_kind == 1) {      // pre-split X: the stage is complete once both dZ units are stored
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                        ++it;
                    }
                } else {
                    const uint32_t sb_u32 = st_u32 + 2 * TC_A_TILE_FLOATS * 4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pi = t + 128 * j, k = (cur_kind - 2) * 16 + (pi >> 5), c = pi & 31;
                        if ((c >> 3) < cur_ngroups) {
                            uint4 hi, lo;
                            split_f16x8(v[2 * j], v[2 * j + 1], hi, lo);
                            const uint32_t off = (uint32_t)((c >> 3) * 8192 + k * 128 + (((c & 7) ^ (k & 7)) << 4));
                            sts128u(sb_u32 + off, hi);
                            sts128u(sb_u32 + TC_B_TILE_FLOATS * 4 + off, lo);
                        }
                    }
                    if (cur_kind == UNITS - 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                        ++it;
                    }
                }
            } else if (cur_kind == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int ci = t + 128 * i, k = ci >> 5, mc = ci & 31;
                    float4 hi, lo;
                    split_tf32(v[i], hi, lo);
                    const uint32_t off = mn32_offset(mc >> 3, k, mc & 7, TC_BK);
                    sts128(st_u32 + off, hi);
                    sts128(st_u32 + TC_A_TILE_FLOATS * 4 + off, lo);
                }
            } else {
                const uint32_t sb_u32 = st_u32 + 2 * TC_A_TILE_FLOATS * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int ci = t + 128 * ((cur_kind - 1) * 8 + i), k = ci >> 6, nc = ci & 63;
                    if ((nc >> 3) < cur_ngroups) {
                        float4 hi, lo;
                        split_tf32(v[i], hi, lo);
                        const uint32_t off = mn32_offset(nc >> 3, k, nc & 7, TC_BK);
                        sts128(sb_u32 + off, hi);
                        sts128(sb_u32 + TC_B_TILE_FLOATS * 4 + off, lo);
                    }
                }
                if (cur_kind == 2) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    ++it;
                }
            }
            // advance the consume cursor
            ++u;
            if (++c_kind == UNITS) {
                c_kind = 0;
                c_r0 += BKW;
                if (c_r0 >= c_rend) {
                    cw += gridDim.x;
                    while (cw < num_work) {
                        c_set_work(cw);
                        if (c_r0 < c_rend) break;
                        cw += gridDim.x;
                    }
                }
            }
        }
        cp_async_wait<0>();
        TC_PROF_NOW(p_end);
        TC_PROF_OUT(t == 0, 21, p_wait_acc);
        TC_PROF_OUT(t == 0, 22, p_cp_acc);
        TC_PROF_OUT(t == 0, 23, p_end - p_begin);
    } else if (warp >= 12) {
        setmaxnreg_dec<32>();
        if (warp == 12 && lane == 0) {
            uint32_t it = 0;
            TC_PROF_DECL(t0 = 0, t1 = 0, t2 = 0, t3 = 0, m_wacc = 0, m_wfull = 0, m_issue = 0, m_begin = 0);
            TC_PROF_NOW(m_begin);
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                const int ot = w % out_tiles, ks = w / out_tiles;
                const int n0 = (ot >> 1) * p.tile_n;
                const int nv = min(p.tile_n, (p.Nout - n0 + 15) & ~15);
                const uint32_t idesc = F16 ? make_idesc_f16(TC_BM, nv, 1, 1) : make_idesc_tf32(TC_BM, nv, 1, 1);
                const int r_beg = ks * p.rows_per_split, r_end = min(p.R, r_beg + p.rows_per_split);
                for (int r0 = r_beg; r0 < r_end; r0 += BKW, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    TC_PROF_NOW(t0);
                    mbar_wait(&tempty[s], ph ^ 1);
                    TC_PROF_NOW(t1);
                    mbar_wait(&full[s], ph);
                    TC_PROF_NOW(t2);
                    tc_fence_after();
                    const uint32_t d = tmem_base + s * TC_N;
                    const uint32_t sa = smem_u32(smem + s * TCF_STAGE_BYTES);
                    // MN-major.  32-bit: LBO = distance between 32-feature groups (4096 B), SBO = between 4-row atoms
                    // (512 B), layout SWIZZLE_128B_BASE32B; 16-bit: 64-feature groups 8192 B apart, 8-row atoms (1024 B),
                    // plain SWIZZLE_128B
                    constexpr uint32_t LBO = F16 ? 8192 : 4096, SBO = F16 ? 1024 : 512;
                    constexpr uint64_t LT = F16 ? 2 : 1;
                    const uint64_t a_hi = make_desc_sw128(sa, LBO, SBO, LT);
                    const uint64_t a_lo = make_desc_sw128(sa + TC_A_TILE_FLOATS * 4, LBO, SBO, LT);
                    const uint64_t b_hi = make_desc_sw128(sa + 2 * TC_A_TILE_FLOATS * 4, LBO, SBO, LT);
                    const uint64_t b_lo = make_desc_sw128(sa + 2 * TC_A_TILE_FLOATS * 4 + TC_B_TILE_FLOATS * 4, LBO, SBO, LT);
                    // one instruction consumes 8 (tf32) / 16 (fp16) reduction rows = two atoms: 1024 / 2048 bytes
                    constexpr int KSTEP = F16 ? 2048 : 1024;
                    auto mma = [&](uint64_t a, uint64_t b, uint32_t acc) {
                        if constexpr (F16) tc_mma_f16(d, a, b, idesc, acc);
                        else tc_mma_tf32(d, a, b, idesc, acc);
                    };
#pragma unroll
                    for (int k = 0; k < 4; ++k) {     // cross terms first (see tc_gemm_fwd_kernel)
                        const uint64_t adv = (uint64_t)((k * KSTEP) >> 4);
                        mma(a_lo + adv, b_hi + adv, k != 0 ? 1u : 0u);
                        mma(a_hi + adv, b_lo + adv, 1);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * KSTEP) >> 4);
                        mma(a_hi + adv, b_hi + adv, 1);
                    }
                    tc_commit(&empty[s]);
                    tc_commit(&tfull[s]);
                    TC_PROF_NOW(t3);
                    TC_PROF_ADD(m_wacc, t0, t1);
                    TC_PROF_ADD(m_wfull, t1, t2);
                    TC_PROF_ADD(m_issue, t2, t3);
                }
            }
            TC_PROF_OUT(true, 16, (unsigned long long)it);
            TC_PROF_OUT(true, 17, m_wacc);
            TC_PROF_OUT(true, 18, m_wfull);
            TC_PROF_OUT(true, 19, m_issue);
            TC_PROF_OUT(true, 20, t3 - m_begin);
        }
        __syncwarp();
    } else {
        setmaxnreg_inc<192>();
        const int q = warp & 3, half = (warp - 4) >> 2, ew = warp - 4;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int ot = w % out_tiles, ks = w / out_tiles;
            const int mh = ot & 1, n0 = (ot >> 1) * p.tile_n;
            const int nv = min(p.tile_n, (p.Nout - n0 + 15) & ~15);
            const int r_beg = ks * p.rows_per_split, r_end = min(p.R, r_beg + p.rows_per_split);
            float acc[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) acc[i] = 0.f;
            for (int r0 = r_beg; r0 < r_end; r0 += BKW, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&tfull[s], ph);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (half * 128 + j * 32 < nv) {      // warp-uniform: columns the MMA actually wrote
                        float v[32];
                        tmem_ld32(tmem_base + s * TC_N + half * 128 + j * 32 + lane_bits, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[s]);
            }
            if constexpr (F16) {
                float sc, inv;
                f16_scale_from_absmax(__ldg(p.dz_absmax_bits), sc, inv);   // undo the power-of-two scale of dZ (exact)
#pragma unroll
                for (int i = 0; i < 128; ++i) acc[i] *= inv;
            }
            if (r_beg < r_end) {
                // the CTA's partial tile goes to G with 16-byte vector atomics straight from the registers (thread =
                // one dZ feature row, 4 consecutive X features per instruction); this happens once per ~40 stages
                float *grow = p.G + (size_t)(mh * 128 + q * 32 + lane) * p.ldg;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (half * 128 + j * 32 < nv) {
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const int col = n0 + half * 128 + j * 32 + c4 * 4;
                            float *dst = grow + col;
                            if (col + 3 < p.Nout && ((p.ldg & 3) == 0)) {
                                red_add_v4(dst, acc[j * 32 + c4 * 4], acc[j * 32 + c4 * 4 + 1], acc[j * 32 + c4 * 4 + 2], acc[j * 32 + c4 * 4 + 3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (col + e < p.Nout) red_add_f32(dst + e, acc[j * 32 + c4 * 4 + e]);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
}  // namespace dcc

// dcc_common.cuh — shared helpers for the sm_100a kernels of the dcc hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dcc_b200.h"

namespace dcc {

// ---- error plumbing ---------------------------------------------------------------------------
void set_last_cuda_error(cudaError_t e, const char *what, const char *file, int line);

#define DCC_CUDA_TRY(expr)                                                   \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) {                                             \
            ::dcc::set_last_cuda_error(_e, #expr, __FILE__, __LINE__);       \
            return DCC_ERR_CUDA;                                             \
        }                                                                    \
    } while (0)

// Makes `dev` the current device for the scope of an entry point and restores the caller's device on the way out
// (a handle is bound to one GPU; the caller's current device may be another one).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != dev) {
            err = cudaSetDevice(dev);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define DCC_DEVICE_GUARD(dev)                                                          \
    ::dcc::DeviceGuard _dcc_dev_guard(dev);                                            \
    do {                                                                               \
        if (_dcc_dev_guard.err != cudaSuccess) {                                       \
            ::dcc::set_last_cuda_error(_dcc_dev_guard.err, "cudaSetDevice(handle device)", __FILE__, __LINE__); \
            return DCC_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- separately rounded float64 / float32 arithmetic -------------------------------------------
// The env state math must reproduce NumPy's operation-by-operation rounding, so nothing here may be
// contracted into an FMA by the compiler: the *_rn intrinsics are never fused.
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// squared 2-norm exactly as this image's numpy evaluates ddot(x,x) inside np.linalg.norm:
// x0*x0 rounded, then one fused multiply-add with x1 (see oracle/dcc_env_oracle.c header).
__device__ __forceinline__ double sqnorm2(double x, double y) { return __fma_rn(y, y, __dmul_rn(x, x)); }

// ---- bulk async copy (TMA, non-tensor form): shared::cta -> global ------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace dcc

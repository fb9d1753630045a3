// tools/store_test.cu — per-SM global-store throughput of the GEMM epilogue's access pattern, in isolation (tuning aid).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) store_kernel(float *C, int tiles, int mode, unsigned long long *cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    unsigned long long t0 = clock64();
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        float *base = C + (size_t)tile * 128 * 256;
        if (mode == 0) {   // 512-byte row segments: one STG.128 per row per warp (the kernel's pattern)
            for (int r = 0; r < 32; ++r)
                *reinterpret_cast<float4 *>(base + (size_t)(q * 32 + r) * 256 + half * 128 + lane * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
        } else if (mode == 1) {   // whole 1 KB rows: two STG.128 per row per warp, warps split rows 16-wise
            for (int r = 0; r < 16; ++r)
                for (int h = 0; h < 2; ++h)
                    *reinterpret_cast<float4 *>(base + (size_t)(warp * 16 + r) * 256 + h * 128 + lane * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
        } else {           // thread-per-row 16-byte pieces
            for (int c = 0; c < 32; ++c)
                *reinterpret_cast<float4 *>(base + (size_t)(q * 32 + lane) * 256 + half * 128 + c * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
        }
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    const int tiles = 1550;
    float *C; unsigned long long *cyc, h;
    cudaMalloc(&C, (size_t)tiles * 128 * 256 * 4); cudaMalloc(&cyc, 8);
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        store_kernel<<<148, 256>>>(C, tiles, mode, cyc); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) store_kernel<<<148, 256>>>(C, tiles, mode, cyc);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %.1f us for %.0f MB = %.2f TB/s; CTA0 %.0f cycles per 128 KB tile\n", mode, ms * 1e3, tiles * 131072.0 / 1e6,
               tiles * 131072.0 / ms * 1e-9, (double)h / ((tiles + 147) / 148));
    }
    return 0;
}

#include <cstdio>
#include <cstdint>
extern __shared__ uint8_t dyn[];
__global__ void k(unsigned *out) { if (threadIdx.x == 0) out[blockIdx.x] = (unsigned)__cvta_generic_to_shared(dyn); }
__global__ void k2(unsigned *out) { __shared__ float s[37]; s[threadIdx.x % 37] = 1.f; __syncthreads(); if (threadIdx.x == 0) out[blockIdx.x] = (unsigned)__cvta_generic_to_shared(dyn) + (s[3] > 5.f); }
int main() {
    unsigned *d, h[4];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 231680);
    k<<<2, 32, 231680>>>(d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("dyn smem base (no static): 0x%x 0x%x\n", h[0], h[1]);
    k<<<2, 32, 1000>>>(d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("dyn smem base (small): 0x%x\n", h[0]);
    k2<<<2, 64, 1000>>>(d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); printf("dyn smem base (with 148 B static): 0x%x  err=%d\n", h[0], (int)cudaGetLastError());
    return 0;
}

// dcc_runtime.cu — status strings, last-error slot, pinned host memory: the small runtime around the kernels.
#include <string.h>

#include "dcc_common.cuh"

namespace dcc {
static thread_local char g_last_err[512] = "";

void set_last_cuda_error(cudaError_t e, const char *what, const char *file, int line) {
    snprintf(g_last_err, sizeof g_last_err, "%s: %s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), what,
             file, line);
    (void)cudaGetLastError();  // clear the sticky-free error slot so later calls report their own failures
}
}  // namespace dcc

extern "C" {

const char *dcc_status_string(int status) {
    switch (status) {
        case DCC_OK: return "ok";
        case DCC_ERR_INVALID_ARG: return "invalid argument";
        case DCC_ERR_CUDA: return "CUDA runtime error (see dcc_last_cuda_error)";
        case DCC_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case DCC_ERR_ALLOC: return "allocation failed";
        case DCC_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

const char *dcc_last_cuda_error(void) { return dcc::g_last_err; }

int dcc_abi_version(void) { return DCC_ABI_VERSION; }

int dcc_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return DCC_ERR_INVALID_ARG;
    *ptr = nullptr;
    DCC_CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return DCC_OK;
}

int dcc_host_free(void *ptr) {
    if (!ptr) return DCC_OK;
    DCC_CUDA_TRY(cudaFreeHost(ptr));
    return DCC_OK;
}

}  // extern "C"

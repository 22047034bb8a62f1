// Library-level bookkeeping of the C-ABI: version, last CUDA error, error strings.
#include "sw_common.cuh"

static thread_local int g_last_cuda_error = 0;
void sw_set_last_cuda_error(int e) { g_last_cuda_error = e; }

extern "C" int sw_abi_version(void) { return 2; }
extern "C" int sw_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" const char* sw_error_string(int code) {
    switch (code) {
        case SW_OK: return "ok";
        case SW_ERR_ARG: return "invalid argument (null pointer or non-positive size)";
        case SW_ERR_CUDA: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
        case SW_ERR_UNSUPPORTED: return "shape outside the supported range of this kernel";
        default: return "unknown error code";
    }
}

// ABI bookkeeping for libglare_b200.so (include/glare_b200.h).
#include "common.cuh"

GLARE_API int glare_abi_version(void) { return 1; }

GLARE_API const char* glare_error_string(int code) {
    if (code == 0) return "ok";
    if (code == GLARE_ERR_BAD_ARG) return "glare_b200: bad argument (null pointer, negative size or unsupported shape)";
    if (code == GLARE_ERR_UNSUPPORTED) return "glare_b200: configuration not supported by this kernel";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "glare_b200: unknown error";
}

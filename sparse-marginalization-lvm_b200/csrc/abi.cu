// Library-level entry points of the C ABI (include/smlvm.h): version, error strings and the
// cached device query shared by the kernels' launch heuristics.
#include "common.cuh"

namespace smlvm {

int device_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSMsB200;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace smlvm

extern "C" int smlvm_version(void) { return SMLVM_ABI_VERSION; }

extern "C" const char* smlvm_error_string(int code)
{
    switch (code) {
        case SMLVM_OK: return "ok";
        case SMLVM_ERR_ARG: return "invalid argument (null pointer, negative size or bad enum)";
        case SMLVM_ERR_UNSUPPORTED: return "shape outside the limits the kernels are built for";
        case SMLVM_ERR_WORKSPACE: return "workspace missing or too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown smlvm error";
}

// cabi.cu -- library-level entry points of the C ABI (include/bevgen.h).
#include "bevgen_common.cuh"

namespace bevgen {
long long g_launch_count = 0;
}

extern "C" int bevgen_abi_version(void) { return BEVGEN_ABI_VERSION; }

extern "C" void bevgen_constants(double* out) {
    using namespace bevgen;
    out[0] = kH; out[1] = kW; out[2] = kVoxel; out[3] = kOffX; out[4] = kOffY;
    out[5] = kZMin; out[6] = kZMax; out[7] = kZ; out[8] = kMapC; out[9] = kZSpan;
}

extern "C" const char* bevgen_error_string(int code) {
    if (code == 0) return "ok";
    if (code == BEVGEN_E_INVALID_ARG) return "bevgen: invalid argument";
    if (code == BEVGEN_E_WORKSPACE) return "bevgen: workspace missing or too small";
    if (code == BEVGEN_E_UNSUPPORTED) return "bevgen: unsupported size";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "bevgen: unknown error";
}

extern "C" int64_t bevgen_launch_count(void) { return bevgen::g_launch_count; }

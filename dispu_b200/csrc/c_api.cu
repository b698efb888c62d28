// ABI housekeeping of libdispu_b200.so: version, error strings, launch counter.
#include <atomic>

#include "common.cuh"

namespace dispu {

static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int finish_launch() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DISPU_OK : (int)e;
}

}  // namespace dispu

extern "C" int dispu_abi_version(void) { return DISPU_ABI_VERSION; }

extern "C" uint64_t dispu_launch_count(void) { return dispu::g_launches.load(std::memory_order_relaxed); }

extern "C" const char *dispu_error_string(int rc) {
    switch (rc) {
        case DISPU_OK: return "ok";
        case DISPU_EINVAL: return "invalid argument (dimension or attribute out of range)";
        case DISPU_ENULL: return "required pointer is NULL";
        case DISPU_EUNSUPPORTED: return "size outside what the sm_100a kernels are instantiated for";
        case DISPU_EWORKSPACE: return "scratch buffer missing or too small";
        default: break;
    }
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    return "unknown dispu error";
}

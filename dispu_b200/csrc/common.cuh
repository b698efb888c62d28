// Shared helpers of the sm_100a kernels: launch bookkeeping, pinned FP32 arithmetic, warp
// reductions.  Everything in this directory is compiled into libdispu_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dispu_b200.h"

namespace dispu {

// ---- launch bookkeeping ---------------------------------------------------------------------
void count_launch(int n = 1);  // c_api.cu
int finish_launch();           // cudaGetLastError() -> ABI return code; c_api.cu

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#define DISPU_REQUIRE(cond, code) \
    do {                          \
        if (!(cond)) return (code); \
    } while (0)

// ---- pinned arithmetic -------------------------------------------------------------------------
// Squared distance exactly as nvcc contracts the reference CUDA kernels' expression
// (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) for sm_100a (tf_sampling_g.cu:142,
// tf_grouping_g.cu:24, tf_nndistance_g.cu:26): FMUL y*y ; FFMA x*x+. ; FFMA z*z+.
__device__ __forceinline__ float d2_ref_gpu(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// Squared distance as the reference HOST loops evaluate it (g++ -O2 x86-64, no contraction):
// ((dx*dx)+(dy*dy))+(dz*dz)  (tf_interpolate.cpp:76, nanoflann.hpp:323-348).
__device__ __forceinline__ float d2_ref_cpu(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- warp helpers ---------------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace dispu

// Small helper kernels of the generator's dense path (sm_100a):
//  * dispu_attention_pack — the K and V operands of the single-kernel attention (attention_tc.cu) straight from the
//    key | value columns of the projected features, ONE pass: every (cloud, 128-key block) tile is read once (coalesced,
//    staged in shared memory) and written as the two pre-split [hi | lo] TF32 UMMA operands (K: 128 keys x 64 channels,
//    V: 64 channels x 128 keys, both K-major, tc_common.cuh layout) with fully coalesced 16-byte stores.  Replaces two
//    launches of the generic strided pack kernel (PointNonLocalCell, Common/ops.py:326-339 has no such step: it is the
//    price of feeding tensor cores fp32 data at fp32 accuracy).
//  * dispu_linear_small — y = act(x . W + b) for K <= 8 input channels (the 3 -> 24 "layer0" of the feature extractor,
//    Common/ops.py:1449-1452): a 3-term dot product per output does not belong on tensor cores; FP32 FMA chain in
//    ascending k, one thread per (row, 4 columns).
#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kApThreads = 256;
constexpr int kApStride = 132;  // floats per staged row (128 + 4: 16-byte aligned rows, conflict-free LDS.128 across rows)

__global__ void __launch_bounds__(kApThreads)
attention_pack_kernel(int m, const float *__restrict__ kv, long long ld, uint8_t *__restrict__ k_packed,
                      uint8_t *__restrict__ v_packed) {
    extern __shared__ __align__(16) float ap_s[];  // [128 keys][kApStride]: key 64 | value 64
    const int nb = m / 128;
    const int cloud = blockIdx.x / nb, blk = blockIdx.x - cloud * nb;
    const float *src = kv + ((long long)cloud * m + (long long)blk * 128) * ld;
    for (int e = threadIdx.x; e < 128 * 32; e += kApThreads) {  // 128 rows x 32 float4
        const int r = e >> 5, c4 = e & 31;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (long long)r * ld) + c4);
        *reinterpret_cast<float4 *>(ap_s + r * kApStride + c4 * 4) = v;
    }
    __syncthreads();
    uint8_t *ko = k_packed + (size_t)blockIdx.x * 65536;
    uint8_t *vo = v_packed + (size_t)blockIdx.x * 65536;
    // K operand: N = 128 keys, K = 64 channels = 2 chunks x [hi 16 KB | lo 16 KB]; a float4 = 4 consecutive channels of one key
    for (int e = threadIdx.x; e < 2048; e += kApThreads) {
        const int chunk = e >> 10, kk4 = (e & 1023) >> 7, r = e & 127;
        const float4 w = *reinterpret_cast<const float4 *>(ap_s + r * kApStride + chunk * 32 + kk4 * 4);
        const float h0 = tf32_round(w.x), h1 = tf32_round(w.y), h2 = tf32_round(w.z), h3 = tf32_round(w.w);
        uint8_t *dst = ko + (size_t)chunk * 32768 + (size_t)kk4 * 2048 + (size_t)r * 16;
        *reinterpret_cast<float4 *>(dst) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4 *>(dst + 16384) = make_float4(tf32_round(__fsub_rn(w.x, h0)), tf32_round(__fsub_rn(w.y, h1)),
                                                               tf32_round(__fsub_rn(w.z, h2)), tf32_round(__fsub_rn(w.w, h3)));
    }
    // V operand: N = 64 channels, K = 128 keys = 4 chunks x [hi 8 KB | lo 8 KB]; a float4 = 4 consecutive keys of one channel
    for (int e = threadIdx.x; e < 2048; e += kApThreads) {
        const int chunk = e >> 9, kk4 = (e & 511) >> 6, r = e & 63;
        const float *p = ap_s + (chunk * 32 + kk4 * 4) * kApStride + 64 + r;
        const float w0 = p[0], w1 = p[kApStride], w2 = p[2 * kApStride], w3 = p[3 * kApStride];
        const float h0 = tf32_round(w0), h1 = tf32_round(w1), h2 = tf32_round(w2), h3 = tf32_round(w3);
        uint8_t *dst = vo + (size_t)chunk * 16384 + (size_t)kk4 * 1024 + (size_t)r * 16;
        *reinterpret_cast<float4 *>(dst) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4 *>(dst + 8192) = make_float4(tf32_round(__fsub_rn(w0, h0)), tf32_round(__fsub_rn(w1, h1)),
                                                              tf32_round(__fsub_rn(w2, h2)), tf32_round(__fsub_rn(w3, h3)));
    }
}

template <int KMAX>
__global__ void __launch_bounds__(256)
linear_small_kernel(long long M, int K, int N, const float *__restrict__ X, long long ldx, const float *__restrict__ W,
                    const float *__restrict__ bias, int relu, float *__restrict__ Y, long long ldy) {
    const int ng = (N + 3) >> 2;  // column groups of 4
    const long long total = M * ng;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / ng;
        const int c0 = (int)(e - row * ng) * 4;
        float x[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) x[k] = k < K ? __ldg(X + row * ldx + k) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u;
            if (c < N) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < K) acc = __fmaf_rn(x[k], __ldg(W + (size_t)k * N + c), acc);
                if (bias) acc = __fadd_rn(acc, __ldg(bias + c));
                Y[row * ldy + c] = relu ? fmaxf(acc, 0.f) : acc;
            }
        }
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_attention_pack(int b, int m, const float *kv, long long ld, void *k_packed, void *v_packed,
                                    dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && m > 0 && ld >= 128, DISPU_EINVAL);
    DISPU_REQUIRE(m % 128 == 0 && ld % 4 == 0, DISPU_EUNSUPPORTED);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(kv && k_packed && v_packed, DISPU_ENULL);
    DISPU_REQUIRE(((reinterpret_cast<uintptr_t>(kv) | reinterpret_cast<uintptr_t>(k_packed) | reinterpret_cast<uintptr_t>(v_packed)) & 15u) == 0,
                  DISPU_EUNSUPPORTED);
    const long long n_blocks = (long long)b * (m / 128);
    DISPU_REQUIRE(n_blocks <= 0x7fffffffLL, DISPU_EUNSUPPORTED);
    const size_t smem = sizeof(float) * 128 * kApStride;  // 66 KB
    const cudaError_t e = cudaFuncSetAttribute(attention_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    attention_pack_kernel<<<(unsigned)n_blocks, kApThreads, smem, (cudaStream_t)stream>>>(
        m, kv, ld, reinterpret_cast<uint8_t *>(k_packed), reinterpret_cast<uint8_t *>(v_packed));
    return finish_launch();
}

extern "C" int dispu_linear_small(long long M, int K, int N, const float *X, long long ldx, const float *W, const float *bias,
                                  int relu, float *Y, long long ldy, dispu_stream_t stream) {
    DISPU_REQUIRE(M >= 0 && K > 0 && N > 0, DISPU_EINVAL);
    DISPU_REQUIRE(K <= 8, DISPU_EUNSUPPORTED);
    if (M == 0) return DISPU_OK;
    DISPU_REQUIRE(X && W && Y, DISPU_ENULL);
    if (ldx <= 0) ldx = K;
    if (ldy <= 0) ldy = N;
    DISPU_REQUIRE(ldx >= K && ldy >= N, DISPU_EINVAL);
    const long long total = M * ((N + 3) / 4);
    long long grid = ceil_div64(total, 256);
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    count_launch();
    if (K <= 4)
        linear_small_kernel<4><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(M, K, N, X, ldx, W, bias, relu, Y, ldy);
    else
        linear_small_kernel<8><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(M, K, N, X, ldx, W, bias, relu, Y, ldy);
    return finish_launch();
}

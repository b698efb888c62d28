// Tensor-core path of the fused set-abstraction body (group -> centre-subtract -> concat -> MLP ->
// max over nsample) for sm_100a: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM.
// Reference chain replaced: Common/pointnet_util.py:46-53,117-133 + Common/tf_util.py:120-185
// (see sa_mlp.cu).  The grouped MLP is the one tensor-core-eligible contraction of the path.
//
// One persistent CTA per SM (128 threads); a tile is 128 grouped rows = 128/nsample whole groups.
//   setup   every layer's weights are staged ONCE per CTA into shared memory, already split into
//           TF32 hi / lo planes and laid out as K-major UMMA operands (tc_common.cuh); TMEM
//           accumulator (<= 256 columns) allocated once.
//   tile    (1) gather: thread r builds grouped row r (points[idx] | xyz[idx]-centre | 0-pad)
//               straight into the A operand planes (hi, lo) in shared memory;
//           (2) per layer: one thread issues K/8 x {3 | 1} tcgen05.mma (M=128, N=C_out, K=8)
//               and commits to an mbarrier; all threads wait, then thread r reads accumulator
//               ROW r with tcgen05.ld (lane = row), applies bias + ReLU, re-splits into hi/lo and
//               writes the next layer's A operand — activations never leave the SM;
//           (3) last layer: bias + ReLU, then the max over the nsample rows of a group is ONE
//               `redux.sync.max.u32` per column (rows of a group are lanes of a warp; post-ReLU
//               floats order like their bit patterns), and lanes store (group, column) coalesced.
// Precision: DISPU_MLP_TF32X3 issues a_lo*b_hi + a_hi*b_lo + a_hi*b_hi per K-step ("3xTF32"),
// which reproduces the FP32 product to ~2^-21; DISPU_MLP_TF32 issues a_hi*b_hi only.
#include "sa_mlp.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTcRows = 128;
constexpr uint32_t kTmemCols = 256;
constexpr int kTcMaxSmem = 220 * 1024;

struct TcPlan {
    int n_layers;
    int K[kMlpMaxLayers];     // padded input width (multiple of 8)
    int N[kMlpMaxLayers];     // output width (multiple of 16, <= 256)
    int cin[kMlpMaxLayers];   // true input width
    uint32_t offB[kMlpMaxLayers];     // byte offset of the hi plane of layer l's weights; lo follows
    uint32_t offBias[kMlpMaxLayers];  // byte offset of the bias (N floats)
    uint32_t offA;            // A hi plane; lo follows at +planeA
    uint32_t planeA;          // 128 * Kmax * 4
    const float *w[kMlpMaxLayers];
    const float *bias[kMlpMaxLayers];
    int c_pts, c_pts_pad, layout;
};

template <bool kX3, int NS>
__global__ void __launch_bounds__(kTcThreads, 1)
sa_mlp_tc_kernel(int n, int m, const float *__restrict__ xyz, const float *__restrict__ points,
                 const float *__restrict__ new_xyz, const int *__restrict__ idx, TcPlan plan,
                 float *__restrict__ out, long long total_rows, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t smem_base = smem_u32(smem);

    // ---- one-time setup ------------------------------------------------------------------------
    if (t == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);

    for (int l = 0; l < plan.n_layers; ++l) {
        const int K = plan.K[l], N = plan.N[l], cin = plan.cin[l];
        uint8_t *bhi = smem + plan.offB[l];
        uint8_t *blo = bhi + (size_t)K * N * 4;
        const float *W = plan.w[l];
        for (int e = t; e < K * N; e += kTcThreads) {
            const int k = e / N, nn = e - k * N;
            int row = -1;  // reference weight row feeding internal channel k
            if (l == 0) {
                if (k < plan.c_pts) row = plan.layout == 0 ? 3 + k : k;
                else if (k >= plan.c_pts_pad && k < plan.c_pts_pad + 3)
                    row = plan.layout == 0 ? k - plan.c_pts_pad : plan.c_pts + (k - plan.c_pts_pad);
            } else if (k < cin) {
                row = k;
            }
            const float w = row >= 0 ? __ldg(W + (size_t)row * N + nn) : 0.f;
            const float hi = tf32_round(w);
            const uint32_t o = op_offset(N, nn, k);
            *reinterpret_cast<float *>(bhi + o) = hi;
            if (kX3) *reinterpret_cast<float *>(blo + o) = __fsub_rn(w, hi);
        }
        float *sb = reinterpret_cast<float *>(smem + plan.offBias[l]);
        for (int e = t; e < N; e += kTcThreads) sb[e] = plan.bias[l] ? __ldg(plan.bias[l] + e) : 0.f;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = s_tmem;
    uint32_t phase = 0;

    uint8_t *a_hi = smem + plan.offA;
    uint8_t *a_lo = a_hi + plan.planeA;
    const uint32_t a_hi_addr = smem_base + plan.offA;
    const uint32_t a_lo_addr = a_hi_addr + plan.planeA;
    const uint32_t row_off = (uint32_t)(t >> 3) * 128u + (uint32_t)(t & 7) * 16u;  // my row inside a K-chunk
    constexpr uint32_t kLboA = (kTcRows / 8) * 128;                               // 2048
    constexpr int ns = NS;  // rows per group: 8, 16 or 32 (a group never straddles a warp)
    const unsigned gmask = NS == 32 ? 0xffffffffu : (((1u << (NS & 31)) - 1u) << ((lane / NS) * NS));

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- (1) gather my grouped row into the A operand of layer 0 ----------------------------
        {
            const long long R = (long long)tile * kTcRows + t;
            const bool valid = R < total_rows;
            const int K0 = plan.K[0], c_pts = plan.c_pts, cpp = plan.c_pts_pad;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            const float *pf = nullptr;
            if (valid) {
                const long long g = R / ns;
                const long long cloud = g / m;
                const int ii = __ldg(idx + R);
                const float *pxyz = xyz + (cloud * n + ii) * 3;
                const float *ctr = new_xyz + g * 3;
                gx = __fsub_rn(__ldg(pxyz), __ldg(ctr));
                gy = __fsub_rn(__ldg(pxyz + 1), __ldg(ctr + 1));
                gz = __fsub_rn(__ldg(pxyz + 2), __ldg(ctr + 2));
                if (c_pts) pf = points + (cloud * n + ii) * (long long)c_pts;
            }
            const bool vec = (c_pts & 3) == 0 && (reinterpret_cast<uintptr_t>(points) & 15u) == 0;
            for (int c = 0; c < K0 / 4; ++c) {
                const int k0 = c * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    if (k0 + 3 < c_pts && vec) {
                        v = __ldg(reinterpret_cast<const float4 *>(pf + k0));
                    } else {
                        float e4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = k0 + u;
                            float x = 0.f;
                            if (k < c_pts) x = __ldg(pf + k);
                            else if (k == cpp) x = gx;
                            else if (k == cpp + 1) x = gy;
                            else if (k == cpp + 2) x = gz;
                            e4[u] = x;
                        }
                        v = make_float4(e4[0], e4[1], e4[2], e4[3]);
                    }
                }
                const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                const uint32_t o = (uint32_t)c * kLboA + row_off;
                *reinterpret_cast<float4 *>(a_hi + o) = h;
                if (kX3)
                    *reinterpret_cast<float4 *>(a_lo + o) =
                        make_float4(__fsub_rn(v.x, h.x), __fsub_rn(v.y, h.y), __fsub_rn(v.z, h.z), __fsub_rn(v.w, h.w));
            }
        }

        for (int l = 0; l < plan.n_layers; ++l) {
            const int K = plan.K[l], N = plan.N[l];
            // operands written by the generic proxy -> visible to the tensor core
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            // ---- (2) issue the layer's MMAs -------------------------------------------------------
            if (t == 0) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kTcRows, N);
                const uint32_t lboB = (uint32_t)(N / 8) * 128u;
                const uint32_t b_hi_addr = smem_base + plan.offB[l];
                const uint32_t b_lo_addr = b_hi_addr + (uint32_t)K * N * 4u;
                for (int ks = 0; ks < K / 8; ++ks) {
                    const uint64_t ah = make_smem_desc(a_hi_addr + ks * 2 * kLboA, kLboA, 128);
                    const uint64_t bh = make_smem_desc(b_hi_addr + ks * 2 * lboB, lboB, 128);
                    if (kX3) {
                        const uint64_t al = make_smem_desc(a_lo_addr + ks * 2 * kLboA, kLboA, 128);
                        const uint64_t bl = make_smem_desc(b_lo_addr + ks * 2 * lboB, lboB, 128);
                        mma_tf32_ss(tmem_d, al, bh, idesc, ks > 0);
                        mma_tf32_ss(tmem_d, ah, bl, idesc, 1);
                        mma_tf32_ss(tmem_d, ah, bh, idesc, 1);
                    } else {
                        mma_tf32_ss(tmem_d, ah, bh, idesc, ks > 0);
                    }
                }
                mma_commit(&s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
            tc_fence_after_sync();

            // ---- (3) epilogue: my accumulator row ------------------------------------------------
            const float *sb = reinterpret_cast<const float *>(smem + plan.offBias[l]);
            const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
            const bool last = l == plan.n_layers - 1;
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                const int cols = min(32, N - c0);  // 32 or 16
                if (cols == 32) tmem_ld32(trow + c0, v); else tmem_ld16(trow + c0, v);
                tmem_ld_wait();
                if (!last) {
                    // bias + ReLU -> hi/lo split -> next layer's A operand (K-chunks of 4 columns)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (j < cols) {
                            float x[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                x[u] = fmaxf(__fadd_rn(__uint_as_float(v[j + u]), sb[c0 + j + u]), 0.f);
                            const float4 h = make_float4(tf32_round(x[0]), tf32_round(x[1]), tf32_round(x[2]), tf32_round(x[3]));
                            const uint32_t o = (uint32_t)((c0 + j) >> 2) * kLboA + row_off;
                            *reinterpret_cast<float4 *>(a_hi + o) = h;
                            if (kX3)
                                *reinterpret_cast<float4 *>(a_lo + o) = make_float4(
                                    __fsub_rn(x[0], h.x), __fsub_rn(x[1], h.y), __fsub_rn(x[2], h.z), __fsub_rn(x[3], h.w));
                        }
                    }
                } else {
                    // bias + ReLU, max over the group's lanes, coalesced store of (group, column)
                    const long long g = ((long long)tile * kTcRows + t) / ns;
                    const bool gvalid = (long long)tile * kTcRows + (t / ns) * ns < total_rows;
                    float keep[32 / NS];  // window w holds column c0 + w*ns + lane%ns
#pragma unroll
                    for (int w = 0; w < 32 / NS; ++w) keep[w] = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < cols) {
                            const float x = fmaxf(__fadd_rn(__uint_as_float(v[j]), sb[c0 + j]), 0.f);
                            const uint32_t mx = __reduce_max_sync(gmask, __float_as_uint(x));  // x >= 0: bit order == value order
                            if ((j % NS) == (lane % NS)) keep[j / NS] = __uint_as_float(mx);
                        }
                    }
                    if (gvalid) {
#pragma unroll
                        for (int w = 0; w < 32 / NS; ++w) {
                            const int col = c0 + w * ns + (lane % ns);
                            if (col < c0 + cols) out[g * N + col] = keep[w];
                        }
                    }
                }
            }
            // all TMEM reads / smem writes of this layer done before the next MMA overwrites them
            tc_fence_before_sync();
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kTmemCols);
}

// D (128 x N) = A (128 x K) . W (K x N): descriptor / layout self-test used by tests (not a product op).
template <bool kX3>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_selftest_kernel(int K, int N, const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5;
    uint8_t *a_hi = smem, *a_lo = a_hi + 128 * K * 4, *b_hi = a_lo + 128 * K * 4, *b_lo = b_hi + K * N * 4;
    if (t == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);
    for (int e = t; e < 128 * K; e += kTcThreads) {
        const int r = e / K, k = e - r * K;
        const float x = A[e], h = tf32_round(x);
        *reinterpret_cast<float *>(a_hi + op_offset(128, r, k)) = h;
        *reinterpret_cast<float *>(a_lo + op_offset(128, r, k)) = x - h;
    }
    for (int e = t; e < K * N; e += kTcThreads) {
        const int k = e / N, nn = e - k * N;
        const float x = W[e], h = tf32_round(x);
        *reinterpret_cast<float *>(b_hi + op_offset(N, nn, k)) = h;
        *reinterpret_cast<float *>(b_lo + op_offset(N, nn, k)) = x - h;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = s_tmem;
    if (t == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t lboA = 2048, lboB = (uint32_t)(N / 8) * 128u;
        const uint32_t ah0 = smem_u32(a_hi), al0 = smem_u32(a_lo), bh0 = smem_u32(b_hi), bl0 = smem_u32(b_lo);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ah = make_smem_desc(ah0 + ks * 2 * lboA, lboA, 128), bh = make_smem_desc(bh0 + ks * 2 * lboB, lboB, 128);
            if (kX3) {
                const uint64_t al = make_smem_desc(al0 + ks * 2 * lboA, lboA, 128), bl = make_smem_desc(bl0 + ks * 2 * lboB, lboB, 128);
                mma_tf32_ss(tmem_d, al, bh, idesc, ks > 0);
                mma_tf32_ss(tmem_d, ah, bl, idesc, 1);
                mma_tf32_ss(tmem_d, ah, bh, idesc, 1);
            } else {
                mma_tf32_ss(tmem_d, ah, bh, idesc, ks > 0);
            }
        }
        mma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after_sync();
    const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const int cols = min(32, N - c0);
        if (cols == 32) tmem_ld32(trow + c0, v); else tmem_ld16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < cols) D[(size_t)t * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kTmemCols);
}

int sa_mlp_tc_launch(int b, int n, int m, int ns, int c_pts, int layout, const float *xyz, const float *points,
                     const float *new_xyz, const int *idx, const MlpParams &prm, float *out, int precision,
                     cudaStream_t stream, bool *handled) {
    *handled = false;
    if (!(ns == 8 || ns == 16 || ns == 32)) return DISPU_EUNSUPPORTED;
    TcPlan plan;
    plan.n_layers = prm.n_layers;
    plan.c_pts = c_pts;
    plan.c_pts_pad = (c_pts + 3) & ~3;
    plan.layout = layout;
    const bool x3 = precision == DISPU_MLP_TF32X3;
    const int planes = x3 ? 2 : 1;
    uint32_t off = 0;
    int kmax = 0;
    for (int l = 0; l < prm.n_layers; ++l) {
        const int cin = prm.widths[l], N = prm.widths[l + 1];
        if (N % 16 != 0 || N > 256) return DISPU_EUNSUPPORTED;
        const int K = l == 0 ? ((plan.c_pts_pad + 3 + 7) & ~7) : cin;  // hidden widths are multiples of 16 already
        plan.K[l] = K;
        plan.N[l] = N;
        plan.cin[l] = cin;
        plan.w[l] = prm.w[l];
        plan.bias[l] = prm.bias[l];
        plan.offB[l] = off;
        off += (uint32_t)K * N * 4 * planes;
        if (K > kmax) kmax = K;
    }
    for (int l = 0; l < prm.n_layers; ++l) {
        plan.offBias[l] = off;
        off += (uint32_t)plan.N[l] * 4;
    }
    off = (off + 127u) & ~127u;
    plan.offA = off;
    plan.planeA = (uint32_t)kTcRows * kmax * 4;
    off += plan.planeA * planes;
    if (off > (uint32_t)kTcMaxSmem) return DISPU_EUNSUPPORTED;
    *handled = true;
    const long long total_rows = (long long)b * m * ns;
    const long long n_tiles_ll = ceil_div64(total_rows, kTcRows);
    if (n_tiles_ll > 0x7fffffffLL) return DISPU_EUNSUPPORTED;
    const int n_tiles = (int)n_tiles_ll;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < sms ? n_tiles : sms;
    count_launch();
#define DISPU_TC_LAUNCH(X3, NSV)                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = cudaFuncSetAttribute(sa_mlp_tc_kernel<X3, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              (int)off);                                                           \
        if (e_ != cudaSuccess) return (int)e_;                                                                     \
        sa_mlp_tc_kernel<X3, NSV><<<grid, kTcThreads, off, stream>>>(n, m, xyz, points, new_xyz, idx, plan, out,    \
                                                                     total_rows, n_tiles);                         \
    } while (0)
    if (x3) {
        if (ns == 8) DISPU_TC_LAUNCH(true, 8); else if (ns == 16) DISPU_TC_LAUNCH(true, 16); else DISPU_TC_LAUNCH(true, 32);
    } else {
        if (ns == 8) DISPU_TC_LAUNCH(false, 8); else if (ns == 16) DISPU_TC_LAUNCH(false, 16); else DISPU_TC_LAUNCH(false, 32);
    }
#undef DISPU_TC_LAUNCH
    return finish_launch();
}

}  // namespace dispu

// Hidden test hook (not declared in include/dispu_b200.h): D(128xN) = A(128xK).W(KxN) on tcgen05.
extern "C" int dispu__tc_gemm_selftest(int K, int N, const float *A, const float *W, float *D, int x3,
                                       dispu_stream_t stream) {
    using namespace dispu;
    if (K % 8 || N % 16 || N > 256 || K <= 0 || N <= 0) return DISPU_EINVAL;
    const size_t smem = (size_t)(128 * K + K * N) * 4 * 2;
    if (smem > (size_t)kTcMaxSmem) return DISPU_EUNSUPPORTED;
    count_launch();
    if (x3) {
        cudaFuncSetAttribute(tc_gemm_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_selftest_kernel<true><<<1, kTcThreads, smem, (cudaStream_t)stream>>>(K, N, A, W, D);
    } else {
        cudaFuncSetAttribute(tc_gemm_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_selftest_kernel<false><<<1, kTcThreads, smem, (cudaStream_t)stream>>>(K, N, A, W, D);
    }
    return finish_launch();
}

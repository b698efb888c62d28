// Tensor-core path of the fused set-abstraction body (group -> centre-subtract -> concat -> MLP ->
// max over nsample) for sm_100a: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM.
// Reference chain replaced: Common/pointnet_util.py:46-53,117-133 + Common/tf_util.py:120-185
// (see sa_mlp.cu).  The grouped MLP is the one tensor-core-eligible contraction of the path.
//
// One persistent CTA per SM; a tile is 128 grouped rows = 128/nsample whole groups.  The CTA runs
// up to four warpgroups, each with its own tile stream, TMEM region and mbarrier, so one
// warpgroup's tensor-core phase overlaps the others' gather / epilogue phases.
//   setup   every layer's weights are staged ONCE per CTA into shared memory, already split into
//           TF32 hi / lo planes and laid out as K-major UMMA operands (tc_common.cuh).
//   tile    (1) gather: thread r builds grouped row r (points[idx] | xyz[idx]-centre | 0-pad) and
//               writes it as the A operand of layer 0 into TENSOR MEMORY (lane r) with tcgen05.st;
//           (2) per layer: one thread issues K/8 x {3 | 1} tcgen05.mma (A from TMEM, B from smem,
//               M=128, N=C_out, K=8) and commits to the warpgroup's mbarrier; thread r then reads
//               accumulator ROW r with tcgen05.ld, applies bias + ReLU, re-splits into hi/lo and
//               stores the next layer's A operand back to TMEM — activations never touch shared or
//               global memory;
//           (3) last layer: bias + ReLU, then the max over the nsample rows of a group is ONE
//               `redux.sync.max.u32` per column (rows of a group are lanes of a warp; post-ReLU
//               floats order like their bit patterns), and lanes store (group, column) coalesced.
// Precision: DISPU_MLP_TF32X3 issues a_lo*b_hi + a_hi*b_lo + a_hi*b_hi per K-step ("3xTF32"),
// which reproduces the FP32 product to ~2^-21; DISPU_MLP_TF32 issues a_hi*b_hi only.
#include "sa_mlp.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kTcThreads = 128;  // self-test kernels
constexpr int kTcRows = 128;
constexpr uint32_t kTmemCols = 256;
constexpr int kTcMaxSmem = 220 * 1024;

struct TcPlan {
    int n_layers;
    int K[kMlpMaxLayers];     // padded input width (multiple of 8)
    int N[kMlpMaxLayers];     // output width (multiple of 16, <= 256)
    int cin[kMlpMaxLayers];   // true input width
    int dcol[kMlpMaxLayers];  // TMEM column (inside a warpgroup's region) of the layer's accumulator
    uint32_t offB[kMlpMaxLayers];     // smem byte offset of the hi plane of layer l's weights; lo follows
    uint32_t offBias[kMlpMaxLayers];  // smem byte offset of the bias (N floats)
    const float *w[kMlpMaxLayers];
    const float *bias[kMlpMaxLayers];
    int c_pts, c_pts_pad, layout;
    int wg_cols;  // TMEM columns per warpgroup region
    int n_wg;     // warpgroups per CTA (each runs its own tile stream)
};

__device__ __forceinline__ void wg_barrier(int wg, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "r"(threads) : "memory");
}

// Hidden-layer epilogue for COLS accumulator columns starting at c0 of my row: bias + ReLU, TF32 hi/lo
// split, and the result becomes columns [c0, c0+COLS) of the next layer's A operand in TMEM.
template <bool kX3, int COLS>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_row, int dcol, int N, int c0, const float *sb) {
    uint32_t v[32];
    if (COLS == 32) tmem_ld32(tmem_row + dcol + c0, v); else tmem_ld16(tmem_row + dcol + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < COLS; j += 8) {
        const float4 b0 = *reinterpret_cast<const float4 *>(sb + c0 + j);
        const float4 b1 = *reinterpret_cast<const float4 *>(sb + c0 + j + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t h[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float x = fmaxf(__fadd_rn(__uint_as_float(v[j + u]), bb[u]), 0.f);
            const float hh = tf32_round(x);
            h[u] = __float_as_uint(hh);
            lo[u] = __float_as_uint(__fsub_rn(x, hh));
        }
        tmem_st8(tmem_row + c0 + j, h);
        if (kX3) tmem_st8(tmem_row + N + c0 + j, lo);
    }
}

// Last-layer epilogue: max over the NS rows (= lanes) of my group for COLS columns, then bias + ReLU on
// the winner only.  max_rows(relu(x+b)) == relu(max_rows(x)+b) exactly (fp add is monotone in x), so
// the reduction runs on the RAW accumulator through an order-preserving float->int key — ONE
// redux.sync.max.s32 per column when a group is a whole warp, a shuffle butterfly otherwise; lane (c % NS) of the group
// keeps column c and stores it coalesced.
template <int NS, int COLS>
__device__ __forceinline__ void epilogue_pool(uint32_t tmem_d_row, int N, int c0, const float *sb, unsigned gmask,
                                              int lane, bool gvalid, float *out_row) {
    uint32_t v[32];
    if (COLS == 32) tmem_ld32(tmem_d_row + c0, v); else tmem_ld16(tmem_d_row + c0, v);
    tmem_ld_wait();
    constexpr int W = COLS / NS > 0 ? COLS / NS : 1;
    int keep[W];
#pragma unroll
    for (int w = 0; w < W; ++w) keep[w] = 0;
    const int lm = lane % NS;
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
        const int b = (int)v[j];
        int mx = b ^ ((b >> 31) & 0x7fffffff);  // signed-int order == float order
        if (NS == 32) {
            mx = __reduce_max_sync(gmask, mx);
        } else {
            // groups smaller than a warp: xor-shuffle butterfly over the NS lanes of the group.  (redux.sync with one member
            // mask per group returns through a UNIFORM register, so the compiler serialises the groups of a warp behind
            // divergence code — measured ~10x slower than the shuffles in edgeconv_tc.cu's trace)
#pragma unroll
            for (int off = NS / 2; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        }
        if ((j % NS) == lm) keep[j / NS] = mx;
    }
    if (gvalid) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int col = c0 + w * NS + lm;
            if (w * NS + lm < COLS) {
                const int key = keep[w];
                const float x = __int_as_float(key ^ ((key >> 31) & 0x7fffffff));
                out_row[col] = fmaxf(__fadd_rn(x, sb[col]), 0.f);
            }
        }
    }
    (void)N;
}

// NS: rows per group (8, 16, 32).  kX3: 3xTF32 (hi/lo planes) or single-pass TF32.
//
// Activations live in TENSOR MEMORY for the whole MLP: a warpgroup's thread r owns TMEM lane r =
// grouped row r.  The A operand of every layer is read by the tensor core straight from TMEM
// ("TS" tcgen05.mma), the accumulator is read back with tcgen05.ld, bias/ReLU/TF32-split happen in
// registers and the next layer's A operand is written back with tcgen05.st; shared memory holds
// only the weights.  Column plan inside a warpgroup's region for layer l (P = planes):
//   A_hi [0,K_l)  A_lo [K_l,2K_l)  D [dcol_l, dcol_l+N_l)  with dcol_l >= P*K_l and >= P*N_l (next A).
// kSplit: warps per TMEM lane quarter inside a warpgroup (1 or 2).  With 2, a warpgroup is 256 threads and the
// two warps that share a quarter take alternate 8-column gather chunks / 16-column epilogue chunks of the same
// rows, halving the serial length of every gather / epilogue phase (used when the column budget only allows
// two tile streams per SM).
template <bool kX3, int NS, int kSplit>
__global__ void __launch_bounds__(512, 1)
sa_mlp_tc_kernel(int n, int m, const float *__restrict__ xyz, const float *__restrict__ points,
                 const float *__restrict__ new_xyz, const int *__restrict__ idx, TcPlan plan,
                 float *__restrict__ out, long long total_rows, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[4];
    __shared__ uint32_t s_tmem;
    constexpr int P = kX3 ? 2 : 1;

    const int tid = threadIdx.x;
    constexpr int kWgThreads = 128 * kSplit;
    const int wg = tid / kWgThreads, t = tid & 127, lane = tid & 31, q = (tid >> 5) & 3;
    const int half = (tid / 128) % kSplit;  // which of the kSplit warps of my lane quarter I am
    const uint32_t smem_base = smem_u32(smem);

    // ---- one-time setup: barriers, TMEM, weights (hi/lo planes, K-major UMMA layout) ------------
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s_bar[i], 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&s_tmem, 512);
    for (int l = 0; l < plan.n_layers; ++l) {
        const int K = plan.K[l], N = plan.N[l], cin = plan.cin[l];
        uint8_t *bhi = smem + plan.offB[l];
        uint8_t *blo = bhi + (size_t)K * N * 4;
        const float *W = plan.w[l];
        for (int e = tid; e < K * N; e += blockDim.x) {
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
        for (int e = tid; e < N; e += blockDim.x) sb[e] = plan.bias[l] ? __ldg(plan.bias[l] + e) : 0.f;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    const uint32_t tmem_wg = s_tmem + (uint32_t)(wg * plan.wg_cols);       // lane 0 of my warpgroup's region
    const uint32_t tmem_row = tmem_wg + ((uint32_t)(q * 32) << 16);        // my warp's 32 lanes
    uint32_t phase = 0;
    const unsigned gmask = NS == 32 ? 0xffffffffu : (((1u << (NS & 31)) - 1u) << ((lane / NS) * NS));
    const int c_pts = plan.c_pts, cpp = plan.c_pts_pad, K0 = plan.K[0];
    const bool vec = (c_pts & 3) == 0 && (reinterpret_cast<uintptr_t>(points) & 15u) == 0;
    const int tile_stride = gridDim.x * plan.n_wg;

    int tile = blockIdx.x * plan.n_wg + wg;
    int ii_next = 0;
    if (tile < n_tiles) {
        const long long R = (long long)tile * kTcRows + t;
        if (R < total_rows) ii_next = __ldg(idx + R);
    }
    for (; tile < n_tiles; tile += tile_stride) {
        // ---- (1) gather my grouped row -> A operand of layer 0 (TMEM) ---------------------------
        const long long R = (long long)tile * kTcRows + t;
        const bool valid = R < total_rows;
        const int ii = ii_next;
        {
            const long long Rn = R + (long long)tile_stride * kTcRows;  // prefetch next tile's index
            if (tile + tile_stride < n_tiles && Rn < total_rows) ii_next = __ldg(idx + Rn);
        }
        {
            float gx = 0.f, gy = 0.f, gz = 0.f;
            const float *pf = points;
            if (valid) {
                const long long g = R / NS;
                const long long cloud = g / m;
                const float *pxyz = xyz + (cloud * n + ii) * 3;
                const float *ctr = new_xyz + g * 3;
                gx = __fsub_rn(__ldg(pxyz), __ldg(ctr));
                gy = __fsub_rn(__ldg(pxyz + 1), __ldg(ctr + 1));
                gz = __fsub_rn(__ldg(pxyz + 2), __ldg(ctr + 2));
                if (c_pts) pf = points + (cloud * n + ii) * (long long)c_pts;
            }
            for (int k0 = half * 8; k0 < K0; k0 += 8 * kSplit) {
                float x[8];
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int kk = k0 + hlf * 4;
                    if (valid && vec && kk + 3 < c_pts) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(pf + kk));
                        x[hlf * 4] = v.x; x[hlf * 4 + 1] = v.y; x[hlf * 4 + 2] = v.z; x[hlf * 4 + 3] = v.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = kk + u;
                            float e = 0.f;
                            if (valid) {
                                if (k < c_pts) e = __ldg(pf + k);
                                else if (k == cpp) e = gx;
                                else if (k == cpp + 1) e = gy;
                                else if (k == cpp + 2) e = gz;
                            }
                            x[hlf * 4 + u] = e;
                        }
                    }
                }
                uint32_t h[8], lo[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float hh = tf32_round(x[u]);
                    h[u] = __float_as_uint(hh);
                    lo[u] = __float_as_uint(__fsub_rn(x[u], hh));
                }
                tmem_st8(tmem_row + k0, h);
                if (kX3) tmem_st8(tmem_row + K0 + k0, lo);
            }
        }

        for (int l = 0; l < plan.n_layers; ++l) {
            const int K = plan.K[l], N = plan.N[l], dcol = plan.dcol[l];
            tmem_st_wait();  // my A-operand stores have landed
            tc_fence_before_sync();
            wg_barrier(wg, kWgThreads);
            // ---- (2) one thread of the warpgroup issues the layer's MMAs ------------------------
            if (t == 0 && half == 0) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kTcRows, N);
                const uint32_t lboB = (uint32_t)(N / 8) * 128u;
                const uint32_t b_hi_addr = smem_base + plan.offB[l];
                const uint32_t b_lo_addr = b_hi_addr + (uint32_t)K * N * 4u;
                const uint32_t d = tmem_wg + dcol;
                for (int ks = 0; ks < K / 8; ++ks) {
                    const uint64_t bh = make_smem_desc(b_hi_addr + ks * 2 * lboB, lboB, 128);
                    if (kX3) {
                        const uint64_t bl = make_smem_desc(b_lo_addr + ks * 2 * lboB, lboB, 128);
                        mma_tf32_ts(d, tmem_wg + K + ks * 8, bh, idesc, ks > 0);  // a_lo * b_hi
                        mma_tf32_ts(d, tmem_wg + ks * 8, bl, idesc, 1);           // a_hi * b_lo
                        mma_tf32_ts(d, tmem_wg + ks * 8, bh, idesc, 1);           // a_hi * b_hi
                    } else {
                        mma_tf32_ts(d, tmem_wg + ks * 8, bh, idesc, ks > 0);
                    }
                }
                mma_commit(&s_bar[wg]);
            }
            // only the issuing warp polls the mbarrier; the rest of the warpgroup sleeps on a hardware barrier
            // (128-512 threads spinning on try_wait would eat the issue slots the other warpgroups' epilogues need)
            if ((tid % kWgThreads) < 32) mbar_wait(&s_bar[wg], phase);
            phase ^= 1;
            wg_barrier(wg, kWgThreads);
            tc_fence_after_sync();

            // ---- (3) epilogue on my accumulator row -------------------------------------------------
            const float *sb = reinterpret_cast<const float *>(smem + plan.offBias[l]);
            if (kSplit == 1) {
                if (l != plan.n_layers - 1) {
                    for (int c0 = 0; c0 + 32 <= N; c0 += 32) epilogue_hidden<kX3, 32>(tmem_row, dcol, N, c0, sb);
                    if (N & 16) epilogue_hidden<kX3, 16>(tmem_row, dcol, N, N - 16, sb);
                } else {
                    const long long g = R / NS;
                    const bool gvalid = (long long)tile * kTcRows + (t / NS) * NS < total_rows;
                    for (int c0 = 0; c0 + 32 <= N; c0 += 32)
                        epilogue_pool<NS, 32>(tmem_row + dcol, N, c0, sb, gmask, lane, gvalid, out + g * N);
                    if (N & 16) epilogue_pool<NS, 16>(tmem_row + dcol, N, N - 16, sb, gmask, lane, gvalid, out + g * N);
                }
            } else {  // alternate 16-column chunks between the warps sharing my lane quarter
                if (l != plan.n_layers - 1) {
                    for (int c0 = half * 16; c0 < N; c0 += 16 * kSplit) epilogue_hidden<kX3, 16>(tmem_row, dcol, N, c0, sb);
                } else {
                    const long long g = R / NS;
                    const bool gvalid = (long long)tile * kTcRows + (t / NS) * NS < total_rows;
                    for (int c0 = half * 16; c0 < N; c0 += 16 * kSplit)
                        epilogue_pool<NS, 16>(tmem_row + dcol, N, c0, sb, gmask, lane, gvalid, out + g * N);
                }
            }
        }
        // the next gather overwrites [0, P*K0): only my own lanes, after my own tcgen05.wait::ld
    }
    (void)P;
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(s_tmem, 512);
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

// Same product with A held in TMEM ("TS" form): rows are TMEM lanes, K runs along columns.
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_selftest_ts_kernel(int K, int N, const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5;
    uint8_t *b_hi = smem, *b_lo = b_hi + K * N * 4;
    if (t == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
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
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    const uint32_t a_hi_col = 256, a_lo_col = 256 + K;  // D at [0,N), A planes from column 256
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float x = A[(size_t)t * K + k0 + u], hh = tf32_round(x);
            h[u] = __float_as_uint(hh);
            l[u] = __float_as_uint(x - hh);
        }
        tmem_st8(tmem_d + lane_base + a_hi_col + k0, h);
        tmem_st8(tmem_d + lane_base + a_lo_col + k0, l);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t lboB = (uint32_t)(N / 8) * 128u;
        const uint32_t bh0 = smem_u32(b_hi), bl0 = smem_u32(b_lo);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bh = make_smem_desc(bh0 + ks * 2 * lboB, lboB, 128), bl = make_smem_desc(bl0 + ks * 2 * lboB, lboB, 128);
            mma_tf32_ts(tmem_d, tmem_d + a_lo_col + ks * 8, bh, idesc, ks > 0);
            mma_tf32_ts(tmem_d, tmem_d + a_hi_col + ks * 8, bl, idesc, 1);
            mma_tf32_ts(tmem_d, tmem_d + a_hi_col + ks * 8, bh, idesc, 1);
        }
        mma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const int cols = min(32, N - c0);
        if (cols == 32) tmem_ld32(tmem_d + lane_base + c0, v); else tmem_ld16(tmem_d + lane_base + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < cols) D[(size_t)t * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 512);
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
    int need = 0;
    for (int l = 0; l < prm.n_layers; ++l) {
        const int cin = prm.widths[l], N = prm.widths[l + 1];
        if (N % 16 != 0 || N > 256) return DISPU_EUNSUPPORTED;
        const int K = l == 0 ? ((plan.c_pts_pad + 3 + 7) & ~7) : cin;  // hidden widths are multiples of 16 already
        const bool last = l == prm.n_layers - 1;
        int dcol = planes * K;
        if (!last && planes * N > dcol) dcol = planes * N;
        plan.K[l] = K;
        plan.N[l] = N;
        plan.cin[l] = cin;
        plan.dcol[l] = dcol;
        plan.w[l] = prm.w[l];
        plan.bias[l] = prm.bias[l];
        plan.offB[l] = off;
        off += (uint32_t)K * N * 4 * planes;
        if (dcol + N > need) need = dcol + N;
    }
    if (need > 512) return DISPU_EUNSUPPORTED;
    for (int l = 0; l < prm.n_layers; ++l) {
        plan.offBias[l] = off;
        off += (uint32_t)plan.N[l] * 4;
    }
    if (off > (uint32_t)kTcMaxSmem) return DISPU_EUNSUPPORTED;
    *handled = true;
    plan.wg_cols = (need + 31) & ~31;
    plan.n_wg = 512 / plan.wg_cols;
    if (plan.n_wg > 4) plan.n_wg = 4;
    const long long total_rows = (long long)b * m * ns;
    const long long n_tiles_ll = ceil_div64(total_rows, kTcRows);
    if (n_tiles_ll > 0x7fffffffLL) return DISPU_EUNSUPPORTED;
    const int n_tiles = (int)n_tiles_ll;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid_ll = ceil_div64(n_tiles, plan.n_wg);
    const int grid = grid_ll < sms ? (int)grid_ll : sms;
    const int split = plan.n_wg <= 2 ? 2 : 1;
    const int threads = plan.n_wg * 128 * split;
    count_launch();
#define DISPU_TC_LAUNCH(X3, NSV)                                                                                   \
    do {                                                                                                           \
        auto kern_ = split == 2 ? sa_mlp_tc_kernel<X3, NSV, 2> : sa_mlp_tc_kernel<X3, NSV, 1>;                        \
        cudaError_t e_ = cudaFuncSetAttribute(kern_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off);       \
        if (e_ != cudaSuccess) return (int)e_;                                                                     \
        kern_<<<grid, threads, off, stream>>>(n, m, xyz, points, new_xyz, idx, plan, out, total_rows, n_tiles);     \
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
    if (x3 == 2) {  // TS form (A in TMEM), 3xTF32
        if (K > 128) return DISPU_EUNSUPPORTED;
        const size_t sm2 = (size_t)K * N * 4 * 2;
        cudaFuncSetAttribute(tc_gemm_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        tc_gemm_selftest_ts_kernel<<<1, kTcThreads, sm2, (cudaStream_t)stream>>>(K, N, A, W, D);
        return finish_launch();
    }
    if (x3) {
        cudaFuncSetAttribute(tc_gemm_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_selftest_kernel<true><<<1, kTcThreads, smem, (cudaStream_t)stream>>>(K, N, A, W, D);
    } else {
        cudaFuncSetAttribute(tc_gemm_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_selftest_kernel<false><<<1, kTcThreads, smem, (cudaStream_t)stream>>>(K, N, A, W, D);
    }
    return finish_launch();
}

// Dense layer on the 5th-generation tensor cores with FP32-level accuracy:
//     Y[M,N] = act(X[M,K] . W[K,N] + bias)        act = ReLU or identity, N <= 256
// tcgen05.mma kind::tf32 with the 3xTF32 split (x = hi + lo, hi*w_hi + lo*w_hi + hi*w_lo), FP32 accumulators in
// TMEM.  This is the unit the grouped point-feature MLPs of the live generator reduce to once their first
// layer is hoisted (Common/ops.py:1055-1072: conv1 on (b,4N,16,128) rows and the [1,128]-conv = a 2048->256
// contraction per point; tf_util.conv2d, Common/tf_util.py:120-185), with bias and ReLU fused in the epilogue
// instead of separate bias_add / relu passes.
//
// Warp-specialised persistent kernel, one CTA per SM, 13 warps:
//   warps 0-7   producers: two threads per row stream the X tile in 32-column chunks, split into TF32 hi/lo and
//               write both planes in the K-major UMMA layout.  Small weights (K*N_pad*8 B + two stages fit) stay
//               RESIDENT in shared memory for the CTA's lifetime; otherwise the producers also copy the matching
//               pre-packed weight chunk (already split + laid out by dispu_dense_pack) into the stage;
//   warps 8-11  epilogue: accumulator row r -> registers (tcgen05.ld), + bias, ReLU, 16-byte stores;
//   warp  12    one thread issues 4 K-steps x 3 tcgen05.mma per chunk (A and B from shared memory, M=128,
//               N=N_pad) and releases the stage with tcgen05.commit.
// Up to four smem stages (full/empty mbarriers) and two TMEM accumulators (full/empty mbarriers) keep loads,
// MMAs and the epilogue of consecutive tiles overlapped.
#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kDnKC = 32;        // K columns per pipeline chunk
constexpr int kDnMaxStages = 4;
constexpr int kDnProducerWarps = 8;
constexpr int kDnThreads = (kDnProducerWarps + 4 + 1) * 32;  // producers | epilogue | MMA issuer
constexpr int kDnSmemBudget = 216 * 1024;

__host__ __device__ __forceinline__ int dn_npad(int N) { return (N + 15) & ~15; }
__host__ __device__ __forceinline__ int dn_chunks(int K) { return (K + kDnKC - 1) / kDnKC; }
// bytes of one packed weight chunk: hi plane then lo plane, each N_pad rows x 32 k (K-major UMMA layout)
__host__ __device__ __forceinline__ size_t dn_chunk_bytes(int N) { return (size_t)2 * dn_npad(N) * kDnKC * 4; }

// W[k][n] = Wsrc[k*sk + n*sn] (row-major (K,N): sk=N, sn=1) -> per K-chunk [hi | lo] planes in UMMA K-major
// layout, zero padded in K and N.  blockIdx.y = batch (source and destination advance by their batch strides).
__global__ void __launch_bounds__(256)
dense_pack_kernel(int K, int N, const float *__restrict__ W, long long sk, long long sn, long long w_batch_stride,
                  uint8_t *__restrict__ packed, long long packed_batch_stride) {
    const int npad = dn_npad(N);
    const int chunk = blockIdx.x;
    W += (long long)blockIdx.y * w_batch_stride;
    uint8_t *hi = packed + (long long)blockIdx.y * packed_batch_stride + (size_t)chunk * dn_chunk_bytes(N);
    uint8_t *lo = hi + (size_t)npad * kDnKC * 4;
    const bool k_fast = sk < sn;  // walk the contiguous source axis with consecutive threads
    for (int e = threadIdx.x; e < npad * kDnKC; e += 256) {
        int kk, nn;
        if (k_fast) { nn = e / kDnKC; kk = e - nn * kDnKC; }
        else { kk = e / npad; nn = e - kk * npad; }
        const int k = chunk * kDnKC + kk;
        const float w = (k < K && nn < N) ? __ldg(W + (long long)k * sk + (long long)nn * sn) : 0.f;
        const float h = tf32_round(w);
        const uint32_t o = op_offset(npad, nn, kk);
        *reinterpret_cast<float *>(hi + o) = h;
        *reinterpret_cast<float *>(lo + o) = tf32_round(__fsub_rn(w, h));
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Everything one launch needs (kernel parameter; mirrors dispu_dense_args of the C ABI).
struct DenseParams {
    long long M;
    int K, N;
    const float *X;
    long long ldx;
    const uint8_t *packed;
    const float *bias;
    int relu;
    float *Y;
    long long ldy;
    // gather front-end: row r of the A operand = relu(X[cloud(r), gidx[r], :] + sign * Q[r / g_ns, :])
    const int *gidx;
    int g_n, g_ns;
    long long g_rows_per_cloud;
    const float *Q;
    long long ldq;
    float sign;
    // residuals added after the activation
    const float *R1;
    long long ldr1;
    const float *R2;
    long long ldr2;
    // batched weights: rows [i*rows_per_batch, (i+1)*rows_per_batch) use packed + i*packed_batch_stride
    long long rows_per_batch, packed_batch_stride;
    int n_tiles, n_stages;
};

// four consecutive floats k..k+3 of a row (zero beyond K)
__device__ __forceinline__ float4 dn_load4(const float *__restrict__ row, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) {
        v = __ldg(reinterpret_cast<const float4 *>(row + k));
    } else {
        if (k < K) v.x = __ldg(row + k);
        if (k + 1 < K) v.y = __ldg(row + k + 1);
        if (k + 2 < K) v.z = __ldg(row + k + 2);
        if (k + 3 < K) v.w = __ldg(row + k + 3);
    }
    return v;
}

// kResident: the whole packed weight stays in shared memory for the CTA's lifetime (small K*N) and the pipeline
// stages hold only X chunks; otherwise every stage carries its X chunk and the matching weight chunk.
template <bool kResident>
__global__ void __launch_bounds__(kDnThreads, 1)
dense_tf32x3_kernel(const __grid_constant__ DenseParams prm) {
    const long long M = prm.M;
    const int K = prm.K, N = prm.N;
    const float *__restrict__ X = prm.X;
    const uint8_t *__restrict__ packed = prm.packed;
    const float *__restrict__ bias = prm.bias;
    const int relu = prm.relu;
    float *__restrict__ Y = prm.Y;
    const int n_tiles = prm.n_tiles, n_stages = prm.n_stages;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[kDnMaxStages], s_empty[kDnMaxStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[256];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npad = dn_npad(N);
    const int nchunks = dn_chunks(K);
    const uint32_t a_plane = 128u * kDnKC * 4u;  // 16 KB per plane
    const uint32_t b_plane = (uint32_t)npad * kDnKC * 4u;
    const uint32_t stage_bytes = kResident ? 2 * a_plane : 2 * a_plane + 2 * b_plane;  // A_hi | A_lo [| B_hi | B_lo]
    const uint32_t w_base = (uint32_t)n_stages * stage_bytes;                           // resident weights live here
    const uint32_t smem_base = smem_u32(smem);
    constexpr int kMmaWarp = kDnProducerWarps + 4;

    if (tid == 0) {
        for (int s = 0; s < kDnMaxStages; ++s) {
            mbar_init(&s_full[s], kDnProducerWarps);  // one arrival per producer warp
            mbar_init(&s_empty[s], 1);                // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&s_tfull[a], 1);   // tcgen05.commit
            mbar_init(&s_tempty[a], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 512);
    for (int e = tid; e < 256; e += kDnThreads) s_bias[e] = (bias && e < N) ? __ldg(bias + e) : 0.f;
    if (kResident) {
        const float4 *src = reinterpret_cast<const float4 *>(packed);
        float4 *dst = reinterpret_cast<float4 *>(smem + w_base);
        const int n16 = (int)((size_t)nchunks * 2 * b_plane / 16);
        for (int e = tid; e < n16; e += kDnThreads) dst[e] = __ldg(src + e);
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem;

    if (warp < kDnProducerWarps) {
        // ================= producers: thread (r, h) owns float4 columns h*4..h*4+3 of row r's chunk =================
        const int r = tid & 127, h = tid >> 7;
        uint32_t it = 0;  // global chunk counter -> stage / phase
        const uint32_t row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
        const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (prm.ldx % 4 == 0);
        const bool qvec = ((reinterpret_cast<uintptr_t>(prm.Q) & 15u) == 0) && (prm.ldq % 4 == 0);
        const int *__restrict__ gidx = prm.gidx;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long row = (long long)tile * 128 + r;
            const bool rvalid = row < M;
            const long long rowc = rvalid ? row : 0;
            const float *xr;
            const float *qr = nullptr;
            if (gidx) {
                const long long cloud = rowc / prm.g_rows_per_cloud;
                xr = X + (cloud * prm.g_n + __ldg(gidx + rowc)) * prm.ldx;
                if (prm.Q) qr = prm.Q + (rowc / prm.g_ns) * prm.ldq;
            } else {
                xr = X + rowc * prm.ldx;
            }
            const uint8_t *packed_t = packed;
            if (!kResident && prm.rows_per_batch > 0)
                packed_t = packed + ((long long)tile * 128 / prm.rows_per_batch) * prm.packed_batch_stride;
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int s = it % n_stages;
                const uint32_t ph = (it / n_stages) & 1;
                // issue my loads before waiting for the stage: they only need registers
                const int k0 = c * kDnKC + h * 16;
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + j * 4;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rvalid) {
                        v[j] = dn_load4(xr, k, K, vec);
                        if (gidx) {
                            if (qr) {
                                const float4 q = dn_load4(qr, k, K, qvec);
                                v[j].x = __fmaf_rn(prm.sign, q.x, v[j].x);
                                v[j].y = __fmaf_rn(prm.sign, q.y, v[j].y);
                                v[j].z = __fmaf_rn(prm.sign, q.z, v[j].z);
                                v[j].w = __fmaf_rn(prm.sign, q.w, v[j].w);
                            }
                            v[j].x = fmaxf(v[j].x, 0.f);
                            v[j].y = fmaxf(v[j].y, 0.f);
                            v[j].z = fmaxf(v[j].z, 0.f);
                            v[j].w = fmaxf(v[j].w, 0.f);
                        }
                    }
                }
                mbar_wait(&s_empty[s], ph ^ 1);  // stage free (first pass: passes immediately)
                uint8_t *st = smem + (size_t)s * stage_bytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 hi = make_float4(tf32_round(v[j].x), tf32_round(v[j].y), tf32_round(v[j].z), tf32_round(v[j].w));
                    const uint32_t o = (uint32_t)(h * 4 + j) * 2048u + row_off;
                    *reinterpret_cast<float4 *>(st + o) = hi;
                    // lo is rounded (not left to the tensor core's truncation): halves the error of the lo*w_hi term
                    *reinterpret_cast<float4 *>(st + a_plane + o) =
                        make_float4(tf32_round(__fsub_rn(v[j].x, hi.x)), tf32_round(__fsub_rn(v[j].y, hi.y)),
                                    tf32_round(__fsub_rn(v[j].z, hi.z)), tf32_round(__fsub_rn(v[j].w, hi.w)));
                }
                if (!kResident) {  // weight chunk: contiguous copy of the pre-packed [hi | lo] planes
                    const float4 *src = reinterpret_cast<const float4 *>(packed_t + (size_t)c * dn_chunk_bytes(N));
                    float4 *dst = reinterpret_cast<float4 *>(st + 2 * a_plane);
                    const int n16 = (int)(2 * b_plane / 16);
                    for (int e = tid; e < n16; e += kDnProducerWarps * 32) dst[e] = __ldg(src + e);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_full[s]);
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, npad);
            const uint32_t lboB = (uint32_t)(npad / 8) * 128u;
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const int a = tcount & 1;
                const uint32_t aph = (tcount >> 1) & 1;
                mbar_wait(&s_tempty[a], aph ^ 1);  // accumulator drained by the epilogue (first two tiles: immediate)
                tc_fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)a * 256u;
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (it / n_stages) & 1;
                    mbar_wait(&s_full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t ah = smem_base + (uint32_t)s * stage_bytes, al = ah + a_plane;
                    const uint32_t bh = kResident ? smem_base + w_base + (uint32_t)c * 2 * b_plane : ah + 2 * a_plane;
                    const uint32_t bl = bh + b_plane;
#pragma unroll
                    for (int ks = 0; ks < kDnKC / 8; ++ks) {
                        const uint64_t dah = make_smem_desc(ah + ks * 2 * 2048u, 2048u, 128u);
                        const uint64_t dal = make_smem_desc(al + ks * 2 * 2048u, 2048u, 128u);
                        const uint64_t dbh = make_smem_desc(bh + ks * 2 * lboB, lboB, 128u);
                        const uint64_t dbl = make_smem_desc(bl + ks * 2 * lboB, lboB, 128u);
                        mma_tf32_ss(d, dal, dbh, idesc, (c | ks) != 0);
                        mma_tf32_ss(d, dah, dbl, idesc, 1);
                        mma_tf32_ss(d, dah, dbh, idesc, 1);
                    }
                    mma_commit(&s_empty[s]);  // stage reusable once these MMAs have read it
                }
                mma_commit(&s_tfull[a]);  // accumulator complete
            }
        }
    } else {
        // ================= epilogue (4 warps -> TMEM lane quarters 0-3) =================
        const int q = (warp - kDnProducerWarps) & 3;  // warp % 4 must equal the lane quarter it reads
        const int r = q * 32 + lane;
        uint32_t tcount = 0;
        const bool yvec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15u) == 0) && (prm.ldy % 4 == 0);
        const float *__restrict__ R1 = prm.R1;
        const float *__restrict__ R2 = prm.R2;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(&s_tfull[a], aph);
            tc_fence_after_sync();
            const long long row = (long long)tile * 128 + r;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * 256u;
            float *yr = Y + row * prm.ldy;
            const float *r1 = R1 ? R1 + row * prm.ldr1 : nullptr;
            const float *r2 = R2 ? R2 + row * prm.ldr2 : nullptr;
            for (int c0 = 0; c0 < npad; c0 += 16) {
                uint32_t v[32];
                tmem_ld16(trow + c0, v);
                tmem_ld_wait();
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float o[4];
                        const int col = c0 + j;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float x = __fadd_rn(__uint_as_float(v[j + u]), s_bias[col + u]);
                            x = relu ? fmaxf(x, 0.f) : x;
                            if (col + u < N) {
                                if (r1) x = __fadd_rn(x, __ldg(r1 + col + u));
                                if (r2) x = __fadd_rn(x, __ldg(r2 + col + u));
                            }
                            o[u] = x;
                        }
                        if (col + 3 < N && yvec) {
                            *reinterpret_cast<float4 *>(yr + col) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (col + u < N) yr[col + u] = o[u];
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[a]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_dense_packed_bytes(int K, int N) {
    if (K <= 0 || N <= 0 || N > 256) return 0;
    return (size_t)dn_chunks(K) * dn_chunk_bytes(N);
}

extern "C" int dispu_dense_pack_strided(int K, int N, const float *W, long long stride_k, long long stride_n, int batch,
                                        long long w_batch_stride, void *packed, long long packed_batch_stride,
                                        dispu_stream_t stream) {
    DISPU_REQUIRE(K > 0 && N > 0 && batch >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(N <= 256 && batch <= 65535, DISPU_EUNSUPPORTED);
    if (batch == 0) return DISPU_OK;
    DISPU_REQUIRE(W && packed, DISPU_ENULL);
    DISPU_REQUIRE(batch == 1 || packed_batch_stride >= (long long)dispu_dense_packed_bytes(K, N), DISPU_EINVAL);
    DISPU_REQUIRE(packed_batch_stride % 16 == 0, DISPU_EINVAL);
    count_launch();
    dense_pack_kernel<<<dim3(dn_chunks(K), batch), 256, 0, (cudaStream_t)stream>>>(
        K, N, W, stride_k, stride_n, w_batch_stride, reinterpret_cast<uint8_t *>(packed), packed_batch_stride);
    return finish_launch();
}

extern "C" int dispu_dense_pack(int K, int N, const float *W, void *packed, dispu_stream_t stream) {
    return dispu_dense_pack_strided(K, N, W, N, 1, 1, 0, packed, 0, stream);
}

extern "C" int dispu_dense_tf32x3_ex(const dispu_dense_args *a, dispu_stream_t stream) {
    DISPU_REQUIRE(a, DISPU_ENULL);
    const long long M = a->M;
    const int K = a->K, N = a->N;
    DISPU_REQUIRE(M >= 0 && K > 0 && N > 0, DISPU_EINVAL);
    DISPU_REQUIRE(N <= 256, DISPU_EUNSUPPORTED);
    if (M == 0) return DISPU_OK;
    DISPU_REQUIRE(a->X && a->packed && a->Y, DISPU_ENULL);
    const long long ldx = a->ldx > 0 ? a->ldx : K, ldy = a->ldy > 0 ? a->ldy : N;
    DISPU_REQUIRE(ldx >= K && ldy >= N, DISPU_EINVAL);
    if (a->gather_idx) {
        DISPU_REQUIRE(a->gather_n > 0 && a->gather_ns > 0 && a->gather_rows_per_cloud > 0, DISPU_EINVAL);
        DISPU_REQUIRE(!a->Q || (a->ldq > 0 ? a->ldq : K) >= K, DISPU_EINVAL);
    }
    const bool batched = a->rows_per_batch > 0;
    if (batched) {
        DISPU_REQUIRE(a->rows_per_batch % 128 == 0 && M % a->rows_per_batch == 0, DISPU_EUNSUPPORTED);
        DISPU_REQUIRE(a->packed_batch_stride % 16 == 0, DISPU_EINVAL);
    }
    const long long n_tiles_ll = ceil_div64(M, 128);
    DISPU_REQUIRE(n_tiles_ll <= 0x7fffffffLL, DISPU_EUNSUPPORTED);
    const int n_tiles = (int)n_tiles_ll;
    const int npad = dn_npad(N);
    const size_t a_stage = 2 * 128 * kDnKC * 4, b_chunk = 2 * (size_t)npad * kDnKC * 4;
    const size_t w_all = (size_t)dn_chunks(K) * b_chunk;
    const bool resident = !batched && w_all + 2 * a_stage <= (size_t)kDnSmemBudget;
    int stages;
    size_t smem;
    if (resident) {
        stages = (int)(((size_t)kDnSmemBudget - w_all) / a_stage);
        if (stages > kDnMaxStages) stages = kDnMaxStages;
        smem = stages * a_stage + w_all;
    } else {
        stages = (int)((size_t)kDnSmemBudget / (a_stage + b_chunk));
        if (stages > kDnMaxStages) stages = kDnMaxStages;
        smem = stages * (a_stage + b_chunk);
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < sms ? n_tiles : sms;
    DenseParams p;
    p.M = M; p.K = K; p.N = N;
    p.X = a->X; p.ldx = ldx;
    p.packed = reinterpret_cast<const uint8_t *>(a->packed);
    p.bias = a->bias; p.relu = a->relu;
    p.Y = a->Y; p.ldy = ldy;
    p.gidx = a->gather_idx; p.g_n = a->gather_n; p.g_ns = a->gather_ns > 0 ? a->gather_ns : 1;
    p.g_rows_per_cloud = a->gather_rows_per_cloud > 0 ? a->gather_rows_per_cloud : 1;
    p.Q = a->gather_idx ? a->Q : nullptr; p.ldq = a->ldq > 0 ? a->ldq : K; p.sign = a->sign;
    p.R1 = a->R1; p.ldr1 = a->ldr1 > 0 ? a->ldr1 : N;
    p.R2 = a->R2; p.ldr2 = a->ldr2 > 0 ? a->ldr2 : N;
    p.rows_per_batch = batched ? a->rows_per_batch : 0;
    p.packed_batch_stride = a->packed_batch_stride;
    p.n_tiles = n_tiles; p.n_stages = stages;
    cudaError_t e;
    count_launch();
    if (resident) {
        e = cudaFuncSetAttribute(dense_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dense_tf32x3_kernel<true><<<grid, kDnThreads, smem, (cudaStream_t)stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(dense_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dense_tf32x3_kernel<false><<<grid, kDnThreads, smem, (cudaStream_t)stream>>>(p);
    }
    return finish_launch();
}

extern "C" int dispu_dense_tf32x3(long long M, int K, int N, const float *X, const void *packed, const float *bias,
                                  int relu, float *Y, dispu_stream_t stream) {
    dispu_dense_args a = {};
    a.M = M; a.K = K; a.N = N; a.X = X; a.packed = packed; a.bias = bias; a.relu = relu; a.Y = Y;
    return dispu_dense_tf32x3_ex(&a, stream);
}

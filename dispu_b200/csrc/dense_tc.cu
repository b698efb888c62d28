// Dense layer on the 5th-generation tensor cores with FP32-level accuracy:
//     Y[M,N] = act(X[M,K] . W[K,N] + bias)        act = ReLU or identity, N <= 256
// tcgen05.mma kind::tf32 with the 3xTF32 split (x = hi + lo, hi*w_hi + lo*w_hi + hi*w_lo), FP32 accumulators in
// TMEM.  This is the unit the grouped point-feature MLPs of the live generator reduce to once their first
// layer is hoisted (Common/ops.py:1055-1072: conv1 on (b,4N,16,128) rows and the [1,128]-conv = a 2048->256
// contraction per point; tf_util.conv2d, Common/tf_util.py:120-185), with bias and ReLU fused in the epilogue
// instead of separate bias_add / relu passes.
//
// Warp-specialised persistent kernel, one CTA per SM, 17 warps (TS form: the A operand lives in tensor memory):
//   warps 0-7   producers: fetch the operand rows, apply the optional gather front-end (relu(X[idx] + sign*Q)) or softmax
//               transform, split into TF32 hi/lo and write both planes straight into the TMEM A ring (4 slots x 64 columns)
//               with tcgen05.st.  Three loops (template parameter kLean):
//                 2  lean, cp.async landing tiles: two 4-warp groups on alternate K-chunks, thread = row; the rows are
//                    copied global -> shared 16 bytes at a time with the eight lanes of a quarter-warp sharing a row (four
//                    128-byte lines per instruction instead of 32), into XOR-swizzled per-warp tiles read back row per
//                    thread — the gather front-end, and streamed operands when the tiles fit next to the weights;
//                 1  lean, register ring: same two-group structure, one own chunk in registers with immediate refill and
//                    prefetch.global.L2 ahead — streamed operands next to 64 KB weight stages (after_conv, aggregation);
//                 0  general: thread (row, half chunk), scalar fallbacks for any alignment / K, softmax operand mode,
//                    component masks and clock64 probes.
//               The lean loops hold no state in local memory and carry no fallbacks: see the comments at their heads.
//   one producer thread per group also issues the TMA bulk copy (cp.async.bulk + mbarrier transaction bytes) of the
//               pre-packed weight chunk when the weight does not stay resident in shared memory;
//   warps 8-15  epilogue, two per TMEM lane quarter on alternating 16-column blocks: tcgen05.ld -> bias / ReLU /
//               row scale -> shared-memory staging -> 8 rows x 64 B per store instruction (+ residuals), or the
//               weight-net contraction on warp-level f16 MMAs (kEpi == 1);
//   warp  16    one thread issues up to 4 K-steps x 3 tcgen05.mma per chunk (A from TMEM, B from shared memory, M=128,
//               N=N_pad) and releases ring slots with tcgen05.commit.
// full/empty mbarriers on the ring slots and on the TMEM accumulators (two of <= 128 columns, or one of 256) keep loads,
// MMAs and the epilogue of consecutive tiles overlapped.  Development probes (general loop only): DISPU_DENSE_DBG masks
// pipeline components, DISPU_DENSE_TRACE records clock64 traces of the three roles (tools/dense_dbg.py, tools/dense_trace.py);
// DISPU_DENSE_GENERIC=1 forces the general loop, DISPU_DENSE_REGRING=1 the register ring for streamed operands.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kDnKC = 32;        // K columns per pipeline chunk
constexpr int kDnMaxStages = 4;
constexpr int kDnProducerWarps = 8;
constexpr int kDnEpiWarps = 8;      // two per TMEM lane quarter, alternating 16-column blocks
constexpr int kDnThreads = (kDnProducerWarps + kDnEpiWarps + 1) * 32;  // producers | epilogue | MMA issuer
constexpr int kDnXposeBytes = 32 * 128 + 2 * 128;  // lean gather producer, per warp and buffer: 32 operand rows + 2 centre rows of one K-chunk
constexpr int kDnSmemBudget = 192 * 1024;  // + <= 29 KB static (epilogue staging, weight-net rows, bias, barriers) <= 227 KB

__host__ __device__ __forceinline__ int dn_npad(int N) { return (N + 15) & ~15; }
__host__ __device__ __forceinline__ int dn_chunks(int K) { return (K + kDnKC - 1) / kDnKC; }
// bytes of one packed weight chunk: hi plane then lo plane, each N_pad rows x 32 k (K-major UMMA layout)
__host__ __device__ __forceinline__ size_t dn_chunk_bytes(int N) { return (size_t)2 * dn_npad(N) * kDnKC * 4; }

// W[k][n] = Wsrc[k*sk + n*sn] (row-major (K,N): sk=N, sn=1) -> per K-chunk [hi | lo] planes in UMMA K-major
// layout, zero padded in K and N.  blockIdx.y = batch (source and destination advance by their batch strides).
__global__ void __launch_bounds__(256)
dense_pack_kernel(int K, int N, const float *__restrict__ W, long long sk, long long sn, long long w_batch_stride,
                  uint8_t *__restrict__ packed, long long packed_batch_stride, int parts) {
    const int npad = dn_npad(N);
    const int per_chunk = npad * kDnKC;
    const int chunk = blockIdx.x / parts;  // `parts` CTAs share one K-chunk
    const int part = blockIdx.x - chunk * parts;
    W += (long long)blockIdx.y * w_batch_stride;
    uint8_t *hi = packed + (long long)blockIdx.y * packed_batch_stride + (size_t)chunk * dn_chunk_bytes(N);
    uint8_t *lo = hi + (size_t)npad * kDnKC * 4;
    const bool k_fast = sk < sn;  // walk the contiguous source axis with consecutive threads
    for (int e = part * 256 + threadIdx.x; e < per_chunk; e += 256 * parts) {
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
    // column blocks: tile t -> row tile t / n_blocks, block j = t % n_blocks; block j multiplies the operand at
    // (+ j*packed_block_stride bytes) and writes its N columns at column offset j*N of Y
    int n_blocks;
    long long packed_block_stride;
    // weight-net epilogue (kEpi == 1): Y[row/16][ch*16 + t] = sum_s act(acc)[16*(row/16)+s][ch] * wn[s][t],
    // wn[s][t] = relu(((xyz[cloud, gidx[row]] - xyz[row/16]) . wn_w[:,t] + wn_b[t]) * wn_scale)
    const float *wn_xyz, *wn_w, *wn_b;
    float wn_scale;
    // softmax over a row that spans several column blocks / launches, without a softmax pass:
    //   stat_out: the epilogue also writes per (row, column block, warp half) the maximum of the values it stores
    //             -> stat_out[(row * n_blocks + block) * 2 + half];
    //   stat_in : the producers turn every operand element x of row r into exp(x - M_r), M_r = max of the n_stat partial
    //             maxima stat_in[r * n_stat ...], and sum what they produce; the epilogue divides the row by that sum.
    float *stat_out;
    const float *stat_in;
    int n_stat;
    // split output: column blocks j >= y2_block are written to Y2 (row stride ldy2) at columns (j - y2_block)*N —
    // layers that share their input but feed different consumers keep separate (compact) output tensors
    float *Y2;
    long long ldy2;
    int y2_block;
    int n_tiles, n_stages;
    uint32_t xpose_off;  // lean gather producer: byte offset of the per-warp transposition buffers in dynamic shared memory
    long long *dbg_buf;  // development only: clock64 samples of CTA 0 (3 roles x 64 iterations x 8 slots)
    int dbg;  // development only (DISPU_DENSE_DBG): 1 = no A loads, 2 = no weight TMA, 4 = no MMA issue; results are garbage
};

// 64-bit / 64-bit division is a ~100-instruction routine on the GPU; row and tile counts almost always fit 31 bits
__device__ __forceinline__ long long dn_div(long long a, long long b) {
    return (((a | b) >> 31) == 0) ? (long long)((unsigned)a / (unsigned)b) : a / b;
}

// four consecutive floats k..k+3 of a row (zero beyond K).  kWide: the vector path asks L2 for the whole 256-byte
// block — a streamed row's next K-chunk sits in the same block, so it is already L2-resident when the ring reaches it
// (not used for gathered rows, whose table lives in L2 anyway).
template <bool kWide = false>
__device__ __forceinline__ float4 dn_load4(const float *__restrict__ row, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) {
        if (kWide) {
            asm("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(row + k));
        } else {
            v = __ldg(reinterpret_cast<const float4 *>(row + k));
        }
    } else {
        if (k < K) v.x = __ldg(row + k);
        if (k + 1 < K) v.y = __ldg(row + k + 1);
        if (k + 2 < K) v.z = __ldg(row + k + 2);
        if (k + 3 < K) v.w = __ldg(row + k + 3);
    }
    return v;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (UBLKCP), completion reported to an mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// D (16 x 8, fp32) += A (16 x 16, f16, row) . B (16 x 8, f16, col): the whole 16-sample contraction of the weight-net
// epilogue in ONE warp-level instruction per operand pair (a tf32 m16n8k8 form needs two k = 8 steps)
__device__ __forceinline__ void mma_sync_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x0, x1 -> packed f16 pair hi = rn(x) and the SCALED remainder lo = rn((x - hi) * 2^11): x = hi + lo * 2^-11 to ~2^-22 |x|
// (f16 has TF32's 11-bit significand; the scaling keeps the remainder out of the f16 subnormal range, and the products
// that carry it are accumulated separately and weighted by 2^-11 at the end).  |x| must stay below 65504.
__device__ __forceinline__ void f16_split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    const float r0 = __fmul_rn(__fsub_rn(x0, h0), 2048.f), r1 = __fmul_rn(__fsub_rn(x1, h1), 2048.f);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

constexpr int kDnEpiStride = 20;  // floats per staged epilogue row (16 + 4: 16-byte aligned, conflict-free stores)

// kResident: the whole packed weight stays in shared memory for the CTA's lifetime (small K*N) and the pipeline
// stages hold only X chunks; otherwise every stage carries its X chunk and the matching weight chunk.  Weights
// always arrive by TMA bulk copies (one elected thread, mbarrier transaction bytes): the pre-packed planes are
// contiguous, so a chunk is ONE copy and no thread touches the bytes.
// kLean: 0 = general producer loop, 1 = lean loop with a register ring (streamed operand), 2 = lean loop with cp.async
// landing tiles (gather front-end, or a streamed operand when the tiles fit next to the weights)
template <bool kResident, bool kGather, int kEpi, int kLean = 0>
__global__ void __launch_bounds__(kDnThreads, 1)
dense_tf32x3_kernel(const __grid_constant__ DenseParams prm) {
    const long long M = prm.M;
    const int K = prm.K, N = prm.N;
    const float *__restrict__ X = prm.X;
    const uint8_t *__restrict__ packed = prm.packed;
    const float *__restrict__ bias = prm.bias;
    const int relu = prm.relu;
    const int n_tiles = prm.n_tiles, n_stages = prm.n_stages;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[kDnMaxStages], s_empty[kDnMaxStages], s_tfull[2], s_tempty[2], s_wfull;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_bias[1024];  // bias of every column block (n_blocks * N <= 1024 when a bias is given)
    __shared__ __align__(16) float s_epi[kDnEpiWarps][32 * kDnEpiStride];
    __shared__ float s_wnw[64];                                                          // wn_w (3x16) | wn_b (16)
    __shared__ float s_rsum[2][128][2];  // softmax operand mode: row sums of the two K-halves, per tile parity

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npad = dn_npad(N);
    const int nchunks = dn_chunks(K);
    const uint32_t b_plane = (uint32_t)npad * kDnKC * 4u;
    const uint32_t stage_bytes = kResident ? 0u : 2 * b_plane;  // [B_hi | B_lo] of one K-chunk; the A operand lives in TMEM
    const uint32_t w_base = (uint32_t)n_stages * stage_bytes;   // resident weights live here
    // TMEM columns: accumulators at [0, 256) (two buffers of <= 128 columns, or one of 256), A-operand ring at
    // [256, 512): stage s holds the chunk's 32 K-columns as TF32 hi at 256 + 64 s and lo at 256 + 64 s + 32
    const int acc_bufs = npad <= 128 ? 2 : 1;
    const uint32_t smem_base = smem_u32(smem);
    constexpr int kMmaWarp = kDnProducerWarps + kDnEpiWarps;

    if (tid == 0) {
        for (int s = 0; s < kDnMaxStages; ++s) {
            // arrivals per stage: the producer warps that fill it (one 4-warp group in the lean loop, all 8 otherwise) + the TMA issuer's expect_tx
            mbar_init(&s_full[s], (kLean ? kDnProducerWarps / 2 : kDnProducerWarps) + (kResident ? 0 : 1));
            mbar_init(&s_empty[s], 1);                                      // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&s_tfull[a], 1);   // tcgen05.commit
            mbar_init(&s_tempty[a], kDnEpiWarps);  // one arrival per epilogue warp
        }
        mbar_init(&s_wfull, 1);
        fence_mbar_init();
        if (kResident) {  // whole weight, one bulk copy per K-chunk, all on one barrier
            mbar_arrive_expect_tx(&s_wfull, (uint32_t)nchunks * 2u * b_plane);
            for (int c = 0; c < nchunks; ++c)
                bulk_g2s(smem_base + w_base + (uint32_t)c * 2u * b_plane, packed + (size_t)c * 2u * b_plane, 2u * b_plane, &s_wfull);
        }
    }
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 512);
    for (int e = tid; e < 1024; e += kDnThreads) {  // block j's bias at s_bias[j*npad + col]
        const int j = e / npad, col = e - j * npad;
        s_bias[e] = (bias && j < prm.n_blocks && col < N) ? __ldg(bias + (size_t)j * N + col) : 0.f;
    }
    if (kEpi == 1 && tid < 64) s_wnw[tid] = tid < 48 ? __ldg(prm.wn_w + tid) : (prm.wn_b ? __ldg(prm.wn_b + tid - 48) : 0.f);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem;

    if (warp < kDnProducerWarps && kLean == 2) {
        // ================= producers, gather front-end, lean loop =================
        // Operand row r of a tile = relu(X[cloud(r), gidx[r], :] + sign * Q[r / g_ns, :]).  Same two-group structure as the
        // streamed lean loop below, plus COALESCED, REGISTER-FREE gathers: a thread fetching its own row touches one
        // 128-byte line per lane, 32 L1 wavefronts per load instruction, and the SM's single L1 wavefront queue (~1-2 cycles
        // per line) then bounds the kernel (conv1: 2.7 k cycles per chunk against 1.4 k of MMAs, unchanged with the epilogue
        // masked out).  Here the eight lanes of a quarter-warp share a row — copy m of lane (a, j) moves columns 4j..4j+3 of
        // warp row 8a + m, four lines per instruction, row addresses handed around with one shuffle each — and the bytes go
        // straight to shared memory with cp.async (zero-filled past the end), one own chunk ahead, into a double-buffered
        // per-warp tile whose 16-byte chunks are XOR-swizzled by the row so that both the copies and the row-per-thread
        // reads (thread r = TMEM lane r needs row r) are conflict-free.  The centre rows (at most two per warp:
        // g_ns % 16 == 0) ride the same way and are read back as shared-memory broadcasts.
        const int r = tid & 127, g = tid >> 7;
        const bool leader = !kResident && r == 0;
        const uint32_t tm_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
        uint8_t *xbuf = smem + prm.xpose_off + (size_t)warp * (2 * kDnXposeBytes);
        const uint32_t xbuf_s = smem_u32(xbuf);
        const int la = lane >> 3, lj = lane & 7;
        const int nb = prm.n_blocks;
        const int G = (int)gridDim.x;
        const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
        const int total = my_tiles * nchunks;
        int left = total > g ? (total - g + 1) / 2 : 0;  // own chunks still to convert
        const int *__restrict__ gidx = prm.gidx;
        auto cp16 = [&](uint32_t dst, const float *src, bool ok) {  // 16 bytes global -> shared, zeros when !ok
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? src : X), "r"(ok ? 16 : 0) : "memory");
        };
        // gathered table row of my operand row in `t` (-1: past the end -> zeros)
        // The index is kept as (table offset of the cloud, raw gathered index): the raw value is a global load whose first
        // consumer must not sit right behind it (an in-order warp would stall for the whole L2 round trip) — the two are
        // added only when the cursor reaches that tile, two iterations later.
        auto fetch_ridx = [&](int t, int &base) -> int {  // (a streamed operand is the identity gather)
            const long long row = (long long)(nb == 1 ? t : t / nb) * 128 + r;
            base = 0;
            if (t >= n_tiles || row >= M) return -1;
            if (!kGather) return (int)row;
            base = (int)(dn_div(row, prm.g_rows_per_cloud) * prm.g_n);
            return __ldg(gidx + row);
        };
        // centre row my lane copies for tile t (lanes 0-7: the warp's rows 0-15, lanes 8-15: rows 16-31; -1: none)
        auto fetch_qrow = [&](int t) -> int {
            const long long row = (long long)(nb == 1 ? t : t / nb) * 128 + (warp & 3) * 32 + (la & 1) * 16;
            if (!kGather || !prm.Q || lane >= 16 || t >= n_tiles || row >= M) return -1;
            return (int)dn_div(row, prm.g_ns);
        };
        int tile = blockIdx.x, c = g;
        while (c >= nchunks) { c -= nchunks; tile += G; }
        int ridx = -1, ridx_nxt = -1, base_nxt = 0;
        if (left > 0) {
            int b0;
            const int g0 = fetch_ridx(tile, b0);
            ridx = g0 < 0 ? -1 : b0 + g0;
            ridx_nxt = fetch_ridx(tile + G, base_nxt);
        }
        int qrow = left > 0 ? fetch_qrow(tile) : -1;
        auto issue_copies = [&](int cc, uint32_t buf) {  // chunk cc of the tile `ridx` / `qrow` describe -> buffer `buf`
            const int k = cc * kDnKC + lj * 4;
            const uint32_t dst0 = xbuf_s + buf * kDnXposeBytes + (uint32_t)(8 * la) * 128u;
            int src[8];  // all eight row handles first: the copies then issue back to back instead of one per shuffle round trip
#pragma unroll
            for (int m = 0; m < 8; ++m) src[m] = __shfl_sync(kFull, ridx, (lane & 24) + m);
#pragma unroll
            for (int m = 0; m < 8; ++m)
                cp16(dst0 + (uint32_t)m * 128u + (uint32_t)((lj ^ m) << 4), X + (long long)src[m] * prm.ldx + k, src[m] >= 0 && k < K);
            if (kGather && lane < 16)
                cp16(xbuf_s + buf * kDnXposeBytes + 4096u + (uint32_t)la * 128u + (uint32_t)lj * 16u, prm.Q + (long long)qrow * prm.ldq + k,
                     qrow >= 0 && k < K);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint32_t buf = 0;
        if (left > 0) issue_copies(c, 0);
        uint32_t s = g;       // stage of flat chunk g (n_stages >= 2 > g)
        uint32_t ph = 1;      // parity to wait for on the stage's empty barrier (first pass: passes immediately)
        const uint8_t *wsrc = packed;  // leader only: packed weight chunk of (tile, c)
        bool wnew = true;
        const float sign = prm.sign;
        for (; left > 0; --left) {
            if (leader && wnew) {
                const int mt = nb == 1 ? tile : tile / nb;
                wsrc = packed + (long long)(tile - mt * nb) * prm.packed_block_stride;
                if (prm.rows_per_batch > 0) wsrc += dn_div((long long)mt * 128, prm.rows_per_batch) * prm.packed_batch_stride;
            }
            const uint32_t c_cur = (uint32_t)c;
            asm volatile("cp.async.wait_group 0;" ::: "memory");  // my copies of this chunk have landed ...
            __syncwarp();                                          // ... and so have everyone else's; the other buffer is free
            // point the cursor at the next own chunk and request it
            c += 2;
            wnew = false;
            if (c >= nchunks) {  // my next chunk lies in a later tile of this CTA
                int skipped = 0;
                do { c -= nchunks; tile += G; ++skipped; } while (c >= nchunks);
                if (left > 1) {
                    if (skipped == 1) {
                        ridx = ridx_nxt < 0 ? -1 : base_nxt + ridx_nxt;
                    } else {
                        int b0;
                        const int g0 = fetch_ridx(tile, b0);
                        ridx = g0 < 0 ? -1 : b0 + g0;
                    }
                    ridx_nxt = fetch_ridx(tile + G, base_nxt);
                } else {
                    ridx = -1;
                    ridx_nxt = -1;
                }
                qrow = (left > 1) ? fetch_qrow(tile) : -1;
                wnew = true;
            }
            if (left > 1) issue_copies(c, buf ^ 1u);
            mbar_wait(&s_empty[s], ph);
            tc_fence_after_sync();
            if (leader) {  // weight chunk: one TMA bulk copy of the pre-packed [hi | lo] planes
                mbar_arrive_expect_tx(&s_full[s], 2u * b_plane);
                bulk_g2s(smem_base + s * stage_bytes, wsrc + (size_t)c_cur * 2u * b_plane, 2u * b_plane, &s_full[s]);
            }
            const uint32_t tm = tm_row + s * 64u;
            const uint8_t *myx = xbuf + buf * kDnXposeBytes + lane * 128;
            const uint8_t *myq = xbuf + buf * kDnXposeBytes + 4096 + (lane >> 4) * 128;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float vv[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // relu(x + sign*q); q is zero when there is no Q (x + sign*0 == x exactly)
                    const int jj = hf * 4 + q;
                    const float4 v = *reinterpret_cast<const float4 *>(myx + ((jj ^ (lane & 7)) << 4));
                    if (kGather) {
                        const float4 qq = *reinterpret_cast<const float4 *>(myq + (jj << 4));
                        vv[4 * q] = fmaxf(__fmaf_rn(sign, qq.x, v.x), 0.f);
                        vv[4 * q + 1] = fmaxf(__fmaf_rn(sign, qq.y, v.y), 0.f);
                        vv[4 * q + 2] = fmaxf(__fmaf_rn(sign, qq.z, v.z), 0.f);
                        vv[4 * q + 3] = fmaxf(__fmaf_rn(sign, qq.w, v.w), 0.f);
                    } else {
                        vv[4 * q] = v.x; vv[4 * q + 1] = v.y; vv[4 * q + 2] = v.z; vv[4 * q + 3] = v.w;
                    }
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) hi[u] = __float_as_uint(tf32_round_finite(vv[u]));
                tmem_st16(tm + (uint32_t)hf * 16u, hi);
#pragma unroll
                for (int u = 0; u < 16; ++u)  // lo is rounded (not left to the tensor core's truncation): halves the error of the lo*w_hi term
                    lo[u] = __float_as_uint(tf32_round_finite(__fsub_rn(vv[u], __uint_as_float(hi[u]))));
                tmem_st16(tm + 32u + (uint32_t)hf * 16u, lo);
            }
            tmem_st_wait();  // my part of the A operand has landed in tensor memory
            tc_fence_before_sync();
            __syncwarp();    // (also: every lane has read its row before the buffer is refilled two iterations on)
            if (lane == 0) mbar_arrive(&s_full[s]);
            buf ^= 1u;
            s += 2;
            if (s >= (uint32_t)n_stages) { s -= (uint32_t)n_stages; ph ^= 1; }
        }
    } else if (warp < kDnProducerWarps && kLean == 1) {
        // ================= producers, streamed operand, lean loop =================
        // (kLean: no gather front-end, no softmax transform, 16-byte aligned rows, K % 4 == 0 — every plain layer of the step.)
        // Two groups of four warps take ALTERNATE K-chunks: thread (r, g) owns row r = tid & 127 (= TMEM lane r) and converts
        // the whole 32-column chunk of its row for chunks g, g + 2, ... of the CTA's flat (tile, chunk) sequence, so the
        // per-chunk fixed costs (empty-barrier round trip, tcgen05 fences, cursor arithmetic, the arrive) are paid once per
        // 32 columns and a group has two chunk-times per iteration.  The loop is kept SMALL on purpose: the general
        // producer below carries scalar-load fallbacks, bounds checks, the gather / softmax transforms and probes, a
        // body of ~1500 instructions that the four warps of a scheduler stream through the instruction caches once per
        // chunk — its trivial cursor arithmetic alone took ~450 cycles per chunk in the clock64 trace.
        // Register ring of ONE own chunk with immediate refill: as soon as a 16-column half has been read out of its
        // registers the same half of the NEXT own chunk is requested, so every load has a full group period to land;
        // prefetch.global.L2 runs kPf chunks ahead of that for the HBM latency.
        // Nothing of this loop may live in local memory: a spilled cursor variable costs an L1 round trip queued BEHIND the
        // eight 16-byte global loads of the refill (ncu: 20 % of all warp samples sat on LDL consumers), hence the
        // minimal state, the leader-only weight cursor and the absence of probes here.
        const int r = tid & 127, g = tid >> 7;
        const bool leader = !kResident && r == 0;
        const uint32_t tm_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
        const int nb = prm.n_blocks;
        const int G = (int)gridDim.x;
        const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
        const int total = my_tiles * nchunks;
        int left = total > g ? (total - g + 1) / 2 : 0;  // own chunks still to convert
        constexpr int kPf = 8;
        auto prefetch_l2 = [&](const float *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        auto row_of = [&](int t) -> long long { return (long long)(nb == 1 ? t : t / nb) * 128 + r; };
        auto load4 = [&](float4 &a, const float *p, bool ok) {  // K % 4 == 0: a float4 is inside the row or wholly past it
            a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
        };
        // cursor of the chunk whose values sit in xb: tile, chunk c, row pointer xr (nullptr: row past M -> zeros)
        int tile = blockIdx.x, c = g;
        while (c >= nchunks) { c -= nchunks; tile += G; }
        const float *xr = nullptr;
        if (left > 0 && tile < n_tiles && row_of(tile) < M) xr = X + row_of(tile) * prm.ldx;
        float4 xb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) load4(xb[q], xr + c * kDnKC + q * 4, xr != nullptr && c * kDnKC + q * 4 < K);
        if (xr) {  // the L2 lines my later chunks of this row will ask for, and the head of my row in the CTA's next tile
            for (int cc = c + 2; cc < nchunks && cc < c + 2 + kPf; cc += 2) prefetch_l2(xr + cc * kDnKC);
            if (tile + G < n_tiles && row_of(tile + G) < M)
                for (int cc = g; cc < nchunks && cc < kPf; cc += 2) prefetch_l2(X + row_of(tile + G) * prm.ldx + cc * kDnKC);
        }
        uint32_t s = g;       // stage of flat chunk g (n_stages >= 3 > g)
        uint32_t ph = 1;      // parity to wait for on the stage's empty barrier (first pass: passes immediately)
        const uint8_t *wsrc = packed;  // leader only: packed weight chunk of (tile, c)
        bool wnew = true;
        for (; left > 0; --left) {
            if (leader) {
                if (wnew) {
                    const int mt = nb == 1 ? tile : tile / nb;
                    wsrc = packed + (long long)(tile - mt * nb) * prm.packed_block_stride;
                    if (prm.rows_per_batch > 0) wsrc += dn_div((long long)mt * 128, prm.rows_per_batch) * prm.packed_batch_stride;
                }
            }
            const uint32_t c_cur = (uint32_t)c;
            // the next own chunk (its loads are issued while this one is converted)
            c += 2;
            wnew = false;
            if (c >= nchunks) {  // rare path: my next chunk lies in a later tile of this CTA
                do { c -= nchunks; tile += G; } while (c >= nchunks);
                xr = nullptr;
                if (left > 1 && tile < n_tiles && row_of(tile) < M) {
                    xr = X + row_of(tile) * prm.ldx;
                    if (tile + G < n_tiles && row_of(tile + G) < M)  // head of my row in the tile after that one
                        for (int cc = g; cc < nchunks && cc < kPf; cc += 2) prefetch_l2(X + row_of(tile + G) * prm.ldx + cc * kDnKC);
                }
                wnew = true;
            }
            const float *np = (left > 1 && xr) ? xr + c * kDnKC : nullptr;
            const int krem = K - c * kDnKC;  // columns of the next chunk that exist
            mbar_wait(&s_empty[s], ph);
            tc_fence_after_sync();
            if (leader) {  // weight chunk: one TMA bulk copy of the pre-packed [hi | lo] planes
                mbar_arrive_expect_tx(&s_full[s], 2u * b_plane);
                bulk_g2s(smem_base + s * stage_bytes, wsrc + (size_t)c_cur * 2u * b_plane, 2u * b_plane, &s_full[s]);
            }
            const uint32_t tm = tm_row + s * 64u;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = xb[4 * hf + q];
                    hi[4 * q] = __float_as_uint(tf32_round_finite(v.x)); hi[4 * q + 1] = __float_as_uint(tf32_round_finite(v.y));
                    hi[4 * q + 2] = __float_as_uint(tf32_round_finite(v.z)); hi[4 * q + 3] = __float_as_uint(tf32_round_finite(v.w));
                }
                tmem_st16(tm + (uint32_t)hf * 16u, hi);
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // lo is rounded (not left to the tensor core's truncation): halves the error of the lo*w_hi term
                    const float4 v = xb[4 * hf + q];
                    lo[4 * q] = __float_as_uint(tf32_round_finite(__fsub_rn(v.x, __uint_as_float(hi[4 * q]))));
                    lo[4 * q + 1] = __float_as_uint(tf32_round_finite(__fsub_rn(v.y, __uint_as_float(hi[4 * q + 1]))));
                    lo[4 * q + 2] = __float_as_uint(tf32_round_finite(__fsub_rn(v.z, __uint_as_float(hi[4 * q + 2]))));
                    lo[4 * q + 3] = __float_as_uint(tf32_round_finite(__fsub_rn(v.w, __uint_as_float(hi[4 * q + 3]))));
                }
                tmem_st16(tm + 32u + (uint32_t)hf * 16u, lo);
#pragma unroll
                for (int q = 0; q < 4; ++q)  // refill this half: the loads have a whole group period (two chunk-times) to land
                    load4(xb[4 * hf + q], np + hf * 16 + q * 4, np != nullptr && hf * 16 + q * 4 < krem);
            }
            tmem_st_wait();  // my part of the A operand has landed in tensor memory
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
            if (np != nullptr && c + kPf < nchunks) prefetch_l2(np + kPf * kDnKC);
            s += 2;
            if (s >= (uint32_t)n_stages) { s -= (uint32_t)n_stages; ph ^= 1; }
        }
    } else if (warp < kDnProducerWarps) {
        // ================= producers, gather front-end =================
        // Thread (r, h): row r = tid & 127 of the tile (= TMEM lane r: tcgen05.st can only write the issuing warp's
        // lane quarter, warp % 4) and K-columns 16h..16h+15 of the chunk.  The A operand is written STRAIGHT INTO TENSOR
        // MEMORY (TS-form tcgen05.mma): no shared-memory A planes, no generic->async proxy fence, and the tensor core
        // reads only the B operand from shared memory — the SS form spent most of the SM's shared-memory bandwidth
        // re-reading A three times per K-step.
        // Flat (tile, chunk) iteration over a register ring of kDepth chunks: the raw global loads of chunk i+kDepth
        // are issued as soon as slot i has been consumed; the gather front-end's transform runs at CONSUME time so it
        // never waits on a load it has just issued, and the gather index of the next tile is fetched a whole tile
        // ahead of the row pointer that depends on it.
        constexpr int kDepth = 2;
        const int r = tid & 127, h = tid >> 7;
        const uint32_t tm_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)h * 16u;
        const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (prm.ldx % 4 == 0);
        const bool qvec = ((reinterpret_cast<uintptr_t>(prm.Q) & 15u) == 0) && (prm.ldq % 4 == 0);
        const int *__restrict__ gidx = prm.gidx;
        const int nb = prm.n_blocks;
        const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        const int total = my_tiles * nchunks;
        int ld_tile = blockIdx.x, ld_c = 0;
        const float *xr = X;
        bool rvalid = false;
        int idx_next = 0;
        auto fetch_idx = [&](int tile) -> int {  // gather index of my row in `tile` (0 when out of range)
            const long long row = (long long)(nb == 1 ? tile : tile / nb) * 128 + r;
            return (kGather && tile < n_tiles && row < M) ? __ldg(gidx + row) : 0;
        };
        auto enter_tile = [&](int idx_cur) {  // row pointers of ld_tile; prefetches the index of the tile after it
            const long long row = (long long)(nb == 1 ? ld_tile : ld_tile / nb) * 128 + r;
            rvalid = ld_tile < n_tiles && row < M;
            const long long rowc = rvalid ? row : 0;
            if (kGather) {
                const long long cloud = dn_div(rowc, prm.g_rows_per_cloud);
                xr = X + (cloud * prm.g_n + idx_cur) * prm.ldx;
                idx_next = fetch_idx(ld_tile + gridDim.x);
            } else {
                xr = X + rowc * prm.ldx;
            }
        };
        float4 xb[kDepth][4];
        float4 qb[4];  // the centre term of the gather front-end: ONE chunk ahead (its rows are shared by g_ns
                       // consecutive operand rows, so these are L1/L2 hits and need no deeper ring)
        // The register ring keeps only 2 chunks (32 KB per SM) in flight — too little to cover HBM latency — so the
        // lines the ring will ask for LATER are pulled into L2 ahead of time with prefetch.global.L2 (no registers, no
        // shared memory): kPf chunks ahead inside a row, and the head of my row in the next tile when a tile starts.
        constexpr int kPf = 8;
        auto prefetch_l2 = [&](const float *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        auto prefetch_row_head = [&](const float *rowp, int from) {  // lines from..kPf-1 of a row, split between the two halves
            for (int cc = from + h; cc < nchunks && cc < kPf; cc += 2) prefetch_l2(rowp + cc * kDnKC);
        };
        const float *q_row = prm.Q;  // centre-term row of the tile being consumed (one per tile: no division per chunk)
        bool q_ok = false;
        auto load_q = [&](int tile_q, int c_q) {  // centre-term loads of chunk (tile_q, c_q)
            if (c_q == 0) {
                const long long row = (long long)(nb == 1 ? tile_q : tile_q / nb) * 128 + r;
                q_ok = kGather && prm.Q && tile_q < n_tiles && row < M;
                q_row = q_ok ? prm.Q + dn_div(row, prm.g_ns) * prm.ldq : prm.Q;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                qb[j] = q_ok ? dn_load4(q_row, c_q * kDnKC + h * 16 + j * 4, K, qvec) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto load_chunk = [&](float4 (&xv)[4]) {  // raw loads of chunk (ld_tile, ld_c); advances the cursor
            const int k0 = ld_c * kDnKC + h * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + j * 4;
                xv[j] = (rvalid && !(prm.dbg & 1)) ? dn_load4<!kGather>(xr, k, K, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!kGather && rvalid && ld_c + kPf < nchunks && ((ld_c + kPf) & 1) == h) prefetch_l2(xr + (ld_c + kPf) * kDnKC);
            if (++ld_c == nchunks) {
                ld_c = 0;
                ld_tile += gridDim.x;
                enter_tile(idx_next);
                if (rvalid && !kGather) {  // (a gather source is a small table that lives in L2 anyway)
                    const int nt = ld_tile + (int)gridDim.x;
                    const long long nrow = (long long)(nb == 1 ? nt : nt / nb) * 128 + r;
                    if (nt < n_tiles && nrow < M) prefetch_row_head(X + nrow * prm.ldx, 0);
                }
            }
        };
        enter_tile(fetch_idx(ld_tile));
        if (!kGather) {  // the rest of my first row and my row of this CTA's second tile
            if (rvalid) prefetch_row_head(xr, kDepth);
            const int nt = ld_tile + (int)gridDim.x;
            const long long nrow = (long long)(nb == 1 ? nt : nt / nb) * 128 + r;
            if (nt < n_tiles && nrow < M) prefetch_row_head(X + nrow * prm.ldx, 0);
        }
#pragma unroll
        for (int d = 0; d < kDepth; ++d)
            if (d < total) load_chunk(xb[d]);
        if (kGather) load_q(blockIdx.x, 0);
        uint32_t it = 0;  // global chunk counter -> stage / phase
        int tile = blockIdx.x, c = 0;
        float sm_ml2 = 0.f, sm_sum = 0.f;  // softmax operand mode: max * log2(e) of the row, running sum of my half of the row
        bool sm_ok = false;
        uint32_t ptile = 0;                // tiles this CTA has produced (parity selects the row-sum buffer)
        const uint8_t *packed_t = packed;
        for (int base = 0; base < total; base += kDepth) {
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                if (base + d < total) {
                    const int s = it % n_stages;
                    const uint32_t ph = (it / n_stages) & 1;
                    if (!kResident && c == 0) {
                        const int mt = nb == 1 ? tile : tile / nb;
                        packed_t = packed + (long long)(tile - mt * nb) * prm.packed_block_stride;
                        if (prm.rows_per_batch > 0) packed_t += dn_div((long long)mt * 128, prm.rows_per_batch) * prm.packed_batch_stride;
                    }
                    if (prm.stat_in && c == 0) {  // row maximum (times log2 e) of my row of the tile consumed next
                        const long long row = (long long)(nb == 1 ? tile : tile / nb) * 128 + r;
                        float mx = 0.f;
                        sm_ok = row < M;
                        if (sm_ok) {
                            const float *st = prm.stat_in + row * prm.n_stat;
                            mx = __ldg(st);
                            for (int i = 1; i < prm.n_stat; ++i) mx = fmaxf(mx, __ldg(st + i));
                        }
                        sm_ml2 = __fmul_rn(mx, 1.4426950408889634f);
                        sm_sum = 0.f;
                    }
                    const bool ptrace = prm.dbg_buf && blockIdx.x == 0 && tid == 0 && it < 64;
                    long long tq0 = 0, tq1 = 0, tq2 = 0;
                    if (ptrace) tq0 = clock64();
                    mbar_wait(&s_empty[s], ph ^ 1);  // stage free (first pass: passes immediately)
                    tc_fence_after_sync();
                    if (ptrace) tq1 = clock64();
                    if (!kResident && tid == 0 && (prm.dbg & 2)) mbar_arrive(&s_full[s]);
                    if (!kResident && tid == 0 && !(prm.dbg & 2)) {  // weight chunk: one TMA bulk copy of the pre-packed [hi | lo] planes
                        mbar_arrive_expect_tx(&s_full[s], 2u * b_plane);
                        bulk_g2s(smem_base + (uint32_t)s * stage_bytes, packed_t + (size_t)c * 2u * b_plane, 2u * b_plane, &s_full[s]);
                    }
                    if (!(prm.dbg & 8)) {
#pragma unroll
                        for (int j2 = 0; j2 < 2; ++j2) {  // 8 columns at a time
                            uint32_t hi8[8], lo8[8];
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj) {
                                float4 v = xb[d][2 * j2 + jj];
                                if (kGather) {  // relu(x + sign*q); q is zero when there is no Q (x + sign*0 == x exactly)
                                    const float4 q = qb[2 * j2 + jj];
                                    v.x = fmaxf(__fmaf_rn(prm.sign, q.x, v.x), 0.f);
                                    v.y = fmaxf(__fmaf_rn(prm.sign, q.y, v.y), 0.f);
                                    v.z = fmaxf(__fmaf_rn(prm.sign, q.z, v.z), 0.f);
                                    v.w = fmaxf(__fmaf_rn(prm.sign, q.w, v.w), 0.f);
                                }
                                float vv[4] = {v.x, v.y, v.z, v.w};
                                if (!kGather && prm.stat_in) {  // exp(x - max) = 2^(x log2e - max log2e); rows past M stay zero
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        float e2;
                                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(__fmaf_rn(vv[u], 1.4426950408889634f, -sm_ml2)));
                                        vv[u] = sm_ok ? e2 : 0.f;
                                        sm_sum = __fadd_rn(sm_sum, vv[u]);
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float hh = tf32_round_finite(vv[u]);
                                    hi8[4 * jj + u] = __float_as_uint(hh);
                                    // lo is rounded (not left to the tensor core's truncation): halves the error of the lo*w_hi term
                                    lo8[4 * jj + u] = __float_as_uint(tf32_round_finite(__fsub_rn(vv[u], hh)));
                                }
                            }
                            tmem_st8(tm_row + (uint32_t)s * 64u + (uint32_t)j2 * 8u, hi8);
                            tmem_st8(tm_row + (uint32_t)s * 64u + 32u + (uint32_t)j2 * 8u, lo8);
                        }
                        if (ptrace) tq2 = clock64();
                        tmem_st_wait();  // my part of the A operand has landed in tensor memory
                    }
                    if (prm.stat_in && c == nchunks - 1) s_rsum[ptile & 1][r][h] = sm_sum;  // published by the arrive below
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_full[s]);
                    if (ptrace) {
                        long long *b = prm.dbg_buf + (size_t)it * 8;
                        b[0] = tq0; b[1] = tq1; b[2] = tq2; b[3] = clock64();
                    }
                    if (base + d + kDepth < total) load_chunk(xb[d]);  // refill the slot just consumed
                    ++it;
                    if (++c == nchunks) { c = 0; tile += gridDim.x; ++ptile; }
                    if (kGather) load_q(tile, c);  // centre term of the chunk consumed next
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, npad);
            const uint32_t lboB = (uint32_t)(npad / 8) * 128u;
            uint32_t it = 0, tcount = 0;
            if (kResident) {
                mbar_wait(&s_wfull, 0);
                tc_fence_after_sync();
            }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const int a = acc_bufs == 2 ? (tcount & 1) : 0;
                const uint32_t aph = acc_bufs == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
                mbar_wait(&s_tempty[a], aph ^ 1);  // accumulator drained by the epilogue (first use: immediate)
                tc_fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)a * 128u;
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (it / n_stages) & 1;
                    mbar_wait(&s_full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t ta = tmem_base + 256u + (uint32_t)s * 64u;  // A hi columns of this stage (lo at +32)
                    const uint32_t bh = kResident ? smem_base + w_base + (uint32_t)c * 2 * b_plane : smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t bl = bh + b_plane;
                    // the last chunk of a K that is not a multiple of 32 holds zero columns: issue only the K-steps that carry data
                    // (K = 132 / 136, the [xyz | features] layers: 17 instead of 20 K-steps per tile)
                    const int nks = c == nchunks - 1 ? (K - c * kDnKC + 7) >> 3 : kDnKC / 8;
                    if (!(prm.dbg & 4))
#pragma unroll
                    for (int ks = 0; ks < kDnKC / 8; ++ks) {
                        if (ks >= nks) break;
                        const uint64_t dbh = make_smem_desc(bh + ks * 2 * lboB, lboB, 128u);
                        const uint64_t dbl = make_smem_desc(bl + ks * 2 * lboB, lboB, 128u);
                        mma_tf32_ts(d, ta + 32u + ks * 8u, dbh, idesc, (c | ks) != 0);  // a_lo * b_hi
                        mma_tf32_ts(d, ta + ks * 8u, dbl, idesc, 1);                    // a_hi * b_lo
                        mma_tf32_ts(d, ta + ks * 8u, dbh, idesc, 1);                    // a_hi * b_hi
                    }
                    mma_commit(&s_empty[s]);  // stage reusable once these MMAs have read it
                }
                mma_commit(&s_tfull[a]);  // accumulator complete
            }
        }
    } else {
        // ================= epilogue (8 warps: two per TMEM lane quarter, alternating 16-column blocks) =================
        const int ew = warp - kDnProducerWarps;
        const int q = ew & 3;      // == warp % 4: the TMEM lane quarter this warp may read
        const int half = ew >> 2;  // which of the quarter's two warps
        uint32_t tcount = 0;
        const bool yvec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(prm.Y) & 15u) == 0) && (prm.ldy % 4 == 0);
        const bool r1vec = ((reinterpret_cast<uintptr_t>(prm.R1) & 15u) == 0) && (prm.ldr1 % 4 == 0) && (N % 4 == 0);
        const bool r2vec = ((reinterpret_cast<uintptr_t>(prm.R2) & 15u) == 0) && (prm.ldr2 % 4 == 0) && (N % 4 == 0);
        float *stg = s_epi[ew];
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int a = acc_bufs == 2 ? (tcount & 1) : 0;
            const uint32_t aph = acc_bufs == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
            mbar_wait(&s_tfull[a], aph);
            tc_fence_after_sync();
            const int mt = prm.n_blocks == 1 ? tile : tile / prm.n_blocks;
            const long long row0 = (long long)mt * 128 + q * 32;  // first row of this warp's quarter
            const long long row = row0 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * 128u;
            const int blk = tile - mt * prm.n_blocks;  // column block of this tile
            const bool to_y2 = prm.Y2 != nullptr && blk >= prm.y2_block;
            float *__restrict__ Y = to_y2 ? prm.Y2 + (long long)(blk - prm.y2_block) * N : prm.Y + (long long)blk * N;
            const long long ldy_t = to_y2 ? prm.ldy2 : prm.ldy;
            const float *sb = s_bias + (tile - mt * prm.n_blocks) * npad;
            const float *__restrict__ R1 = prm.R1 ? prm.R1 + (long long)(tile - mt * prm.n_blocks) * N : nullptr;  // residuals
            const float *__restrict__ R2 = prm.R2 ? prm.R2 + (long long)(tile - mt * prm.n_blocks) * N : nullptr;  // follow Y's block
            float rscale = 1.f;  // softmax operand mode: 1 / (sum of my row's exponentials), applied to the raw accumulator
            if (prm.stat_in) {
                const float t = __fadd_rn(s_rsum[tcount & 1][q * 32 + lane][0], s_rsum[tcount & 1][q * 32 + lane][1]);
                rscale = t > 0.f ? __fdividef(1.f, t) : 0.f;
            }
            if (prm.dbg & 64) {
            } else if (kEpi == 1) {
                // ---- weight-net epilogue: this warp's 32 rows are two points x 16 samples.  Lane (p, s) holds its row's
                // neighbour offset; per 16-channel block the raw accumulators are staged in shared memory and each point's
                // 16 x 16 (channel x hidden unit) result is contracted over the 16 samples and written 64 B per channel.
                float gx, gy, gz;  // my row's neighbour offset (lane = (point, sample))
                {
                    const long long rowc = row < M ? row : 0;
                    const long long cloud = dn_div(rowc, prm.g_rows_per_cloud);
                    const float *ctr = prm.wn_xyz + (rowc >> 4) * 3;
                    const float *nbp = prm.wn_xyz + (cloud * prm.g_n + __ldg(prm.gidx + rowc)) * 3;
                    gx = __fsub_rn(__ldg(nbp), __ldg(ctr));
                    gy = __fsub_rn(__ldg(nbp + 1), __ldg(ctr + 1));
                    gz = __fsub_rn(__ldg(nbp + 2), __ldg(ctr + 2));
                }
                // The 16-sample contraction runs on the WARP-level tensor-core path (mma.sync): per point and 16-channel block,
                // H^T[c][t] = sum_s G^T[c][s] . wn[s][t] is a 16 x 16 x 16 product whose A fragments are read straight from the
                // staged activations (8 conflict-free LDS.32) and whose B fragments (wn) stay in registers for the whole tile.
                // A CUDA-core form (4 channels x 4 hidden units per lane on FFMA2, wn rows in shared memory) measures 174 us
                // against 162 us for this one at c2.
                const int fg = lane >> 2, ft = lane & 3;  // mma fragment coordinates: group id, thread in group
                // One m16n8k16 f16 instruction spans all 16 samples (k = s).  3-term split like the main product, on f16 pairs
                // (hi, scaled lo): hi*hi goes to the main accumulator, lo*hi + hi*lo to a correction accumulator weighted by
                // 2^-11 at the end — half the warp-level MMAs of the k = 8 tf32 form, whose issue rate (~26 cycles per
                // instruction and scheduler, ncu: stall_math) bounded this epilogue, and the B fragments fit 16 registers.
                // With 20-float rows the rows a fragment load touches (s = 2 ft [+1] [+8]) start 8 banks apart: the float4
                // staging stores and the scalar fragment loads are conflict-free.
                // B fragments = wn[s][t] = relu(((g_s . Ww[:, t]) + bw[t]) * scale) for my (s, t) pairs, built once per tile
                // from the owning lanes' offsets (shuffles): no shared-memory copy of wn at all.
                uint32_t bh[2][2][2], bl[2][2][2];  // [point][n-block][b0 | b1]: wn[s = 2 ft (+1) | 2 ft + 8 (+1)][t = 8 nb + fg]
                {
                    float w0[2], w1[2], w2[2], wb[2];
#pragma unroll
                    for (int nbk = 0; nbk < 2; ++nbk) {
                        const int t = nbk * 8 + fg;
                        w0[nbk] = s_wnw[t]; w1[nbk] = s_wnw[16 + t]; w2[nbk] = s_wnw[32 + t]; wb[nbk] = s_wnw[48 + t];
                    }
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int kh = 0; kh < 2; ++kh) {
                            float wv[2][2];  // [u][nbk]
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int src = pp * 16 + kh * 8 + 2 * ft + u;
                                const float sx = __shfl_sync(kFull, gx, src), sy = __shfl_sync(kFull, gy, src), sz = __shfl_sync(kFull, gz, src);
#pragma unroll
                                for (int nbk = 0; nbk < 2; ++nbk) {
                                    const float w = __fmaf_rn(sz, w2[nbk], __fmaf_rn(sy, w1[nbk], __fmul_rn(sx, w0[nbk])));
                                    wv[u][nbk] = fmaxf(__fmul_rn(__fadd_rn(w, wb[nbk]), prm.wn_scale), 0.f);
                                }
                            }
#pragma unroll
                            for (int nbk = 0; nbk < 2; ++nbk) f16_split2(wv[0][nbk], wv[1][nbk], bh[pp][nbk][kh], bl[pp][nbk][kh]);
                        }
                }
                for (int c0 = half * 16; c0 < npad; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld16(trow + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 4)  // raw accumulators; bias + activation are applied to the fragments (2 bias loads per block)
                        *reinterpret_cast<uint4 *>(stg + lane * kDnEpiStride + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    const float bias0 = sb[c0 + fg], bias1 = sb[c0 + fg + 8];
                    __syncwarp();
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        // A = G^T (16 channels x 16 samples): a0 (c = fg; s = 2 ft, 2 ft + 1), a1 (fg + 8; same), a2 (fg; s + 8), a3 (fg + 8; s + 8)
                        uint32_t ah[4], al[4];
#pragma unroll
                        for (int kh = 0; kh < 2; ++kh) {
                            const float *gp = stg + (pp * 16 + kh * 8 + 2 * ft) * kDnEpiStride + fg;
                            float x00 = __fadd_rn(gp[0], bias0), x01 = __fadd_rn(gp[kDnEpiStride], bias0);
                            float x10 = __fadd_rn(gp[8], bias1), x11 = __fadd_rn(gp[kDnEpiStride + 8], bias1);
                            if (relu) { x00 = fmaxf(x00, 0.f); x01 = fmaxf(x01, 0.f); x10 = fmaxf(x10, 0.f); x11 = fmaxf(x11, 0.f); }
                            f16_split2(x00, x01, ah[2 * kh], al[2 * kh]);
                            f16_split2(x10, x11, ah[2 * kh + 1], al[2 * kh + 1]);
                        }
                        float acc[2][4], cor[2][4];
#pragma unroll
                        for (int nbk = 0; nbk < 2; ++nbk) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) { acc[nbk][i] = 0.f; cor[nbk][i] = 0.f; }
                            mma_sync_f16(acc[nbk], ah, bh[pp][nbk]);
                            mma_sync_f16(cor[nbk], al, bh[pp][nbk]);
                            mma_sync_f16(cor[nbk], ah, bl[pp][nbk]);
                        }
                        // C = H^T (16 channels x 8 hidden units per n-block): c0,c1 (c = fg, t = 2 ft, 2 ft + 1), c2,c3 (c = fg + 8)
                        const long long pt = (row0 >> 4) + pp;
                        if (row0 + pp * 16 < M) {
                            float *o = prm.Y + pt * prm.ldy + (long long)(c0 + fg) * 16 + 2 * ft;
                            constexpr float kInv = 1.f / 2048.f;
#pragma unroll
                            for (int nbk = 0; nbk < 2; ++nbk) {
                                if (c0 + fg < N)
                                    *reinterpret_cast<float2 *>(o + nbk * 8) =
                                        make_float2(__fmaf_rn(cor[nbk][0], kInv, acc[nbk][0]), __fmaf_rn(cor[nbk][1], kInv, acc[nbk][1]));
                                if (c0 + fg + 8 < N)
                                    *reinterpret_cast<float2 *>(o + 8 * 16 + nbk * 8) =
                                        make_float2(__fmaf_rn(cor[nbk][2], kInv, acc[nbk][2]), __fmaf_rn(cor[nbk][3], kInv, acc[nbk][3]));
                            }
                        }
                    }
                    __syncwarp();
                }
            } else if (yvec) {
                // 16-column blocks: accumulator row -> registers -> (+bias, act) -> staged in shared memory -> written
                // back 8 rows x 64 B per store instruction (full sectors); residuals are read with the same pattern
                const int cq = lane & 3;  // float4 column inside the block
                float run_m = -__int_as_float(0x7f800000);  // row maximum of what this warp stores (stat_out)
                for (int c0 = half * 16; c0 < npad; c0 += 32) {
                    uint32_t v[32];
                    const int eit = (int)tcount * (npad / 32) + c0 / 32;
                    const bool trace = prm.dbg_buf && blockIdx.x == 0 && ew == 0 && lane == 0 && eit < 64;
                    long long te0 = 0, te1 = 0, te2 = 0;
                    if (trace) te0 = clock64();
                    tmem_ld16(trow + c0, v);
                    tmem_ld_wait();
                    if (trace) te1 = clock64();
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4 *>(sb + c0 + j);
                        float4 o = make_float4(__fmaf_rn(__uint_as_float(v[j]), rscale, bv.x), __fmaf_rn(__uint_as_float(v[j + 1]), rscale, bv.y),
                                               __fmaf_rn(__uint_as_float(v[j + 2]), rscale, bv.z), __fmaf_rn(__uint_as_float(v[j + 3]), rscale, bv.w));
                        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4 *>(stg + lane * kDnEpiStride + j) = o;
                        if (prm.stat_out && c0 + j < N)  // running maximum of the columns this warp stores (N % 4 == 0 here)
                            run_m = fmaxf(run_m, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
                    }
                    __syncwarp();
                    if (trace) te2 = clock64();
                    const int col = c0 + cq * 4;
                    if (col < N) {
                        // all residual loads of the block first (independent, in flight together), then add + store
                        float4 r1v[4], r2v[4];
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const long long grow = row0 + pass * 8 + (lane >> 2);
                            r1v[pass] = (R1 && grow < M) ? dn_load4(R1 + grow * prm.ldr1, col, N, r1vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                            r2v[pass] = (R2 && grow < M) ? dn_load4(R2 + grow * prm.ldr2, col, N, r2vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int rr = pass * 8 + (lane >> 2);
                            const long long grow = row0 + rr;
                            if (grow < M) {
                                float4 o = *reinterpret_cast<const float4 *>(stg + rr * kDnEpiStride + cq * 4);
                                if (R1) { o.x = __fadd_rn(o.x, r1v[pass].x); o.y = __fadd_rn(o.y, r1v[pass].y); o.z = __fadd_rn(o.z, r1v[pass].z); o.w = __fadd_rn(o.w, r1v[pass].w); }
                                if (R2) { o.x = __fadd_rn(o.x, r2v[pass].x); o.y = __fadd_rn(o.y, r2v[pass].y); o.z = __fadd_rn(o.z, r2v[pass].z); o.w = __fadd_rn(o.w, r2v[pass].w); }
                                if (!(prm.dbg & 32)) *reinterpret_cast<float4 *>(Y + grow * ldy_t + col) = o;
                            }
                        }
                    }
                    __syncwarp();
                    if (trace) {
                        long long *b = prm.dbg_buf + 1024 + (size_t)eit * 8;
                        b[0] = te0; b[1] = te1; b[2] = te2; b[3] = clock64();
                    }
                }
                if (prm.stat_out && row < M)
                    prm.stat_out[(row * prm.n_blocks + (tile - mt * prm.n_blocks)) * 2 + half] = run_m;
            } else {
                // narrow / unaligned outputs (e.g. the 3-channel coordinate layers): one row per thread, scalar stores
                float *yr = Y + row * ldy_t;
                const float *r1 = R1 ? R1 + row * prm.ldr1 : nullptr;
                const float *r2 = R2 ? R2 + row * prm.ldr2 : nullptr;
                for (int c0 = half * 16; c0 < npad; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld16(trow + c0, v);
                    tmem_ld_wait();
                    if (row < M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = c0 + j;
                            if (col < N) {
                                float x = __fmaf_rn(__uint_as_float(v[j]), rscale, sb[col]);
                                x = relu ? fmaxf(x, 0.f) : x;
                                if (r1) x = __fadd_rn(x, __ldg(r1 + col));
                                if (r2) x = __fadd_rn(x, __ldg(r2 + col));
                                yr[col] = x;
                            }
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

extern "C" size_t dispu_dense_args_size(void) { return sizeof(dispu_dense_args); }

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
    // a few CTAs per chunk so even a two-chunk operand fills the machine
    int parts = (dn_npad(N) * kDnKC + 1023) / 1024;
    if (parts > 8) parts = 8;
    dense_pack_kernel<<<dim3(dn_chunks(K) * parts, batch), 256, 0, (cudaStream_t)stream>>>(
        K, N, W, stride_k, stride_n, w_batch_stride, reinterpret_cast<uint8_t *>(packed), packed_batch_stride, parts);
    return finish_launch();
}

extern "C" int dispu_dense_pack(int K, int N, const float *W, void *packed, dispu_stream_t stream) {
    return dispu_dense_pack_strided(K, N, W, N, 1, 1, 0, packed, 0, stream);
}

static long long *g_dense_dbg_buf = nullptr;  // development only (DISPU_DENSE_TRACE)

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
    const int n_blocks = a->n_blocks > 0 ? a->n_blocks : 1;
    const bool wnet = a->wn_xyz != nullptr;
    if (n_blocks > 1)
        DISPU_REQUIRE(a->packed_block_stride % 16 == 0 && ldy >= (long long)(a->Y2 ? a->y2_block : n_blocks) * N, DISPU_EINVAL);
    if (a->bias) DISPU_REQUIRE((long long)n_blocks * dn_npad(N) <= 1024, DISPU_EUNSUPPORTED);
    if (wnet) {
        DISPU_REQUIRE(a->gather_idx && a->wn_w, DISPU_ENULL);
        DISPU_REQUIRE(a->gather_ns == 16 && N % 32 == 0 && M % 128 == 0 && n_blocks == 1, DISPU_EUNSUPPORTED);
        DISPU_REQUIRE((a->ldy > 0 ? a->ldy : (long long)N * 16) >= (long long)N * 16 && (a->ldy % 4) == 0 &&
                          (reinterpret_cast<uintptr_t>(a->Y) & 15u) == 0, DISPU_EINVAL);
    }
    if (a->Y2) {
        DISPU_REQUIRE(n_blocks > 1 && a->y2_block > 0 && a->y2_block < n_blocks && !wnet && !a->stat_out && !a->R1 && !a->R2, DISPU_EINVAL);
        const long long ldy2 = a->ldy2 > 0 ? a->ldy2 : (long long)(n_blocks - a->y2_block) * N;
        DISPU_REQUIRE(ldy2 >= (long long)(n_blocks - a->y2_block) * N, DISPU_EINVAL);
        DISPU_REQUIRE(N % 4 == 0 && ldy % 4 == 0 && ldy2 % 4 == 0 &&
                          ((reinterpret_cast<uintptr_t>(a->Y) | reinterpret_cast<uintptr_t>(a->Y2)) & 15u) == 0, DISPU_EUNSUPPORTED);
    }
    if (a->stat_out) DISPU_REQUIRE(!wnet && N % 4 == 0 && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(a->Y) & 15u) == 0, DISPU_EUNSUPPORTED);
    if (a->stat_in) DISPU_REQUIRE(!a->gather_idx && a->n_stat > 0 && K % 32 == 0, DISPU_EUNSUPPORTED);
    const long long n_tiles_ll = ceil_div64(M, 128) * n_blocks;
    DISPU_REQUIRE(n_tiles_ll <= 0x7fffffffLL, DISPU_EUNSUPPORTED);
    const int n_tiles = (int)n_tiles_ll;
    const int npad = dn_npad(N);
    const size_t b_chunk = 2 * (size_t)npad * kDnKC * 4;
    const size_t w_all = (size_t)dn_chunks(K) * b_chunk;
    // the A operand is staged in tensor memory (4 ring slots of 64 columns); shared memory holds only weights
    const bool resident = !batched && n_blocks == 1 && w_all <= (size_t)kDnSmemBudget;
    int stages = kDnMaxStages;
    size_t smem = w_all;
    if (!resident) {
        stages = (int)((size_t)kDnSmemBudget / b_chunk);
        if (stages > kDnMaxStages) stages = kDnMaxStages;
        smem = stages * b_chunk;
    }
    static const bool no_lean = getenv("DISPU_DENSE_GENERIC") != nullptr;
    // lean gather producer (coalesced gathers transposed through per-warp shared-memory buffers): aligned rows, K % 4 == 0,
    // centre rows shared by multiples of 16 operand rows
    const bool lean_ok = !no_lean && !a->stat_in && (reinterpret_cast<uintptr_t>(a->X) & 15u) == 0 && ldx % 4 == 0 && K % 4 == 0;
    const bool lean_gather = a->gather_idx && lean_ok && (a->gather_ns > 0 ? a->gather_ns : 1) % 16 == 0 &&
                             (!a->Q || ((reinterpret_cast<uintptr_t>(a->Q) & 15u) == 0 && (a->ldq > 0 ? a->ldq : K) % 4 == 0));
    // a streamed operand takes the same cp.async landing tiles when they leave the weight ring at least three stages
    // (256-column layers with streamed weights keep the register ring: their 64 KB stages need the room)
    const size_t xp = (size_t)kDnProducerWarps * 2 * kDnXposeBytes;
    const bool lean_stream_tiles = !a->gather_idx && lean_ok && !getenv("DISPU_DENSE_REGRING") &&
                                   (resident ? w_all + xp <= (size_t)kDnSmemBudget + 8 * 1024 : ((size_t)kDnSmemBudget - xp) / b_chunk >= 3);
    if (lean_gather || lean_stream_tiles) {
        if (!resident) {
            stages = (int)(((size_t)kDnSmemBudget - xp) / b_chunk);
            if (stages > kDnMaxStages) stages = kDnMaxStages;
            smem = stages * b_chunk;
        }
    }
    const bool lean_gather_ok = (lean_gather || lean_stream_tiles) && (resident ? w_all + xp <= (size_t)kDnSmemBudget + 8 * 1024 : stages >= 2);
    const uint32_t xpose_off = (uint32_t)((smem + 15) & ~(size_t)15);
    if (lean_gather_ok) smem = xpose_off + (size_t)kDnProducerWarps * 2 * kDnXposeBytes;
    if (smem < 1024) smem = 1024;
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
    p.n_blocks = n_blocks; p.packed_block_stride = a->packed_block_stride;
    p.wn_xyz = a->wn_xyz; p.wn_w = a->wn_w; p.wn_b = a->wn_b; p.wn_scale = a->wn_scale;
    p.stat_out = reinterpret_cast<float *>(a->stat_out);
    p.stat_in = reinterpret_cast<const float *>(a->stat_in);
    p.n_stat = a->n_stat;
    p.Y2 = a->Y2; p.y2_block = a->y2_block;
    p.ldy2 = a->ldy2 > 0 ? a->ldy2 : (long long)(n_blocks - a->y2_block) * N;
    if (wnet) p.ldy = a->ldy > 0 ? a->ldy : (long long)N * 16;
    p.n_tiles = n_tiles; p.n_stages = stages;
    p.xpose_off = xpose_off;
    {
        static const int dbg = [] { const char *e = getenv("DISPU_DENSE_DBG"); return e ? atoi(e) : 0; }();
        p.dbg = dbg;
        if (getenv("DISPU_DENSE_TRACE") && !g_dense_dbg_buf) cudaMalloc(&g_dense_dbg_buf, 1536 * sizeof(long long));
        p.dbg_buf = g_dense_dbg_buf;
        if (g_dense_dbg_buf) cudaMemsetAsync(g_dense_dbg_buf, 0, 1536 * sizeof(long long), (cudaStream_t)stream);
    }
    count_launch();
    auto launch = [&](auto kern) -> int {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kDnThreads, smem, (cudaStream_t)stream>>>(p);
        return finish_launch();
    };
    if (lean_gather_ok && !p.dbg) {
        if (!p.gidx) return resident ? launch(dense_tf32x3_kernel<true, false, 0, 2>) : launch(dense_tf32x3_kernel<false, false, 0, 2>);
        if (wnet) return resident ? launch(dense_tf32x3_kernel<true, true, 1, 2>) : launch(dense_tf32x3_kernel<false, true, 1, 2>);
        return resident ? launch(dense_tf32x3_kernel<true, true, 0, 2>) : launch(dense_tf32x3_kernel<false, true, 0, 2>);
    }
    if (wnet) return resident ? launch(dense_tf32x3_kernel<true, true, 1>) : launch(dense_tf32x3_kernel<false, true, 1>);
    const bool lean = !p.gidx && !p.stat_in && !p.dbg && (reinterpret_cast<uintptr_t>(p.X) & 15u) == 0 && ldx % 4 == 0 && K % 4 == 0 && !no_lean;
    if (lean) return resident ? launch(dense_tf32x3_kernel<true, false, 0, 1>) : launch(dense_tf32x3_kernel<false, false, 0, 1>);
    if (resident) return p.gidx ? launch(dense_tf32x3_kernel<true, true, 0>) : launch(dense_tf32x3_kernel<true, false, 0>);
    return p.gidx ? launch(dense_tf32x3_kernel<false, true, 0>) : launch(dense_tf32x3_kernel<false, false, 0>);
}

// development only: copy the clock64 trace of the last launch to the host (1536 int64)
extern "C" int dispu_dense_debug_trace(long long *host_out) {
    if (!g_dense_dbg_buf || !host_out) return DISPU_EUNSUPPORTED;
    return (int)cudaMemcpy(host_out, g_dense_dbg_buf, 1536 * sizeof(long long), cudaMemcpyDeviceToHost);
}

extern "C" int dispu_dense_tf32x3(long long M, int K, int N, const float *X, const void *packed, const float *bias,
                                  int relu, float *Y, dispu_stream_t stream) {
    dispu_dense_args a = {};
    a.M = M; a.K = K; a.N = N; a.X = X; a.packed = packed; a.bias = bias; a.relu = relu; a.Y = Y;
    return dispu_dense_tf32x3_ex(&a, stream);
}

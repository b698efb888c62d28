// EdgeConv dense block on the tensor cores (tcgen05, 3xTF32, FP32-level accuracy) — the per-edge part of dense_conv
// (Common/ops.py:1897-1915) after the hoisting of edgeconv.cu:
//     y0 = relu(pn[j] + pc[i])                      (gather, 24 channels)
//     y1 = relu(y0 . W1a + t1[i])                   (24 x 24)
//     y2 =      y1 . W2a + y0 . W2b + t2[i]         (48 x 24, no activation)
//     out[i] = [max_s y2 | max_s y1 | max_s y0 | f_i]          s = the k = 16 neighbours of point i
// with proj[i] = [pn | pc | t1 | t2] (96 floats per point) from the projection kernel / the feature-kNN tail.
// The round-1 kernel ran these products as packed FFMA2 chains at 254 registers per thread (12 % occupancy, tensor
// pipe 0 %); here a 128-row tile = 8 points x 16 edges lives in tensor memory for the whole block, exactly like the
// grouped MLP of the set abstraction (sa_mlp_tc.cu):
//   thread r of a warpgroup owns TMEM lane r = edge (point r/16, neighbour r%16): it gathers y0, writes it as the A
//   operand (TF32 hi | lo) with tcgen05.st; one thread issues  D[:, 0:32 | 32:64] = y0 . [W2b | W1a]  (K = 24, N = 64:
//   both products that consume y0 in ONE MMA group); the epilogue reads the y1 half, adds the per-point term, applies
//   ReLU, writes y1 back as the next A operand; the second MMA group accumulates  D[:, 0:32] += y1 . W2a  on top of
//   y0 . W2b; the last epilogue adds t2.  Every max over the 16 edges of a point is one redux.sync per column over the
//   16 lanes of the half-warp.  Four warpgroups per CTA, each with its own tile stream, 128 TMEM columns and mbarrier,
//   so one warpgroup's gathers / epilogues overlap the others' tensor-core phases.  Weights ([W2b | W1a] and W2a as
//   K-major UMMA operands, hi / lo planes) are staged once per CTA.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kEtG = 24;             // growth rate (output channels of every layer)
constexpr int kEtProjW = 4 * kEtG;   // pn | pc | t1 | t2
constexpr int kEtK = 16;             // neighbours per point (rows of a group)
constexpr int kEtWg = 4;             // warpgroups (tile streams) per CTA
constexpr int kEtThreads = kEtWg * 128;
constexpr int kEtWgCols = 128;       // TMEM columns per warpgroup: A hi [0,32) | A lo [32,64) | D [64,128)
constexpr uint32_t kEtB1Plane = kEtG * 64 * 4;  // [W2b | W1a]: K = 24, N = 64
constexpr uint32_t kEtB2Plane = kEtG * 32 * 4;  // W2a: K = 24, N = 32
constexpr uint32_t kEtSmem = 2 * kEtB1Plane + 2 * kEtB2Plane;

__device__ __forceinline__ void et_wg_barrier(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// 24 values of my row -> A operand (hi at [0,24), lo at [32,56) of the warpgroup's region)
__device__ __forceinline__ void et_store_a(uint32_t tmem_row, const float (&x)[kEtG]) {
#pragma unroll
    for (int k0 = 0; k0 < kEtG; k0 += 8) {
        uint32_t h[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float hh = tf32_round_finite(x[k0 + u]);
            h[u] = __float_as_uint(hh);
            lo[u] = __float_as_uint(tf32_round_finite(__fsub_rn(x[k0 + u], hh)));
        }
        tmem_st8(tmem_row + (uint32_t)k0, h);
        tmem_st8(tmem_row + 32u + (uint32_t)k0, lo);
    }
}

// max over the 16 rows (lanes of my half-warp) of 24 columns; lane lm ends up with column lm (keep0) and, for lm < 8,
// column 16 + lm (keep1).  Recursive halving with full-warp xor shuffles: at every step a lane hands the half of its
// columns that the partner lane will own to that partner and max-combines what it receives — 15 + 8 shuffles for the 24
// columns.  (A redux.sync per column with one member mask per half-warp is NOT an option: the result lands in a uniform
// register, so the compiler serialises the two halves behind divergence code — measured 5.5k cycles per pool.)
__device__ __forceinline__ void et_pool24(const float (&x)[kEtG], int lm, float &keep0, float &keep1) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = x[i];
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {  // w columns survive this step; partner = lane ^ w
        const bool up = (lm & w) != 0;  // I will own the upper half of my current columns
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? v[i] : v[i + w];
            const float keep = up ? v[i + w] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(kFull, send, w));
        }
    }
    keep0 = v[0];
    float u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = x[16 + i];
#pragma unroll
    for (int w = 4; w >= 1; w >>= 1) {
        const bool up = (lm & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? u[i] : u[i + w];
            const float keep = up ? u[i + w] : u[i];
            u[i] = fmaxf(keep, __shfl_xor_sync(kFull, send, w));
        }
    }
    keep1 = fmaxf(u[0], __shfl_xor_sync(kFull, u[0], 8));  // the two 8-lane halves of the point
}

__device__ __forceinline__ void et_load24(const float *__restrict__ p, float (&v)[kEtG]) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int q = 0; q < kEtG / 4; ++q) {
        const float4 t = __ldg(p4 + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

__global__ void __launch_bounds__(kEtThreads, 1)
edge_mlp_tc_kernel(long long rows /* points */, int n, int c, const float *__restrict__ proj, const int *__restrict__ idx,
                   const float *__restrict__ w1, const float *__restrict__ w2, const float *__restrict__ feat, int ldf,
                   float *__restrict__ out, int ldo, int n_tiles, long long *__restrict__ dbg /* development: clock64 samples */) {
    extern __shared__ __align__(1024) uint8_t et_smem[];  // B1 hi | B1 lo | B2 hi | B2 lo
    __shared__ __align__(8) uint64_t s_bar[kEtWg];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x;
    const int wg = tid >> 7, t = tid & 127, lane = tid & 31, q = (tid >> 5) & 3;
    const uint32_t smem_base = smem_u32(et_smem);

    if (tid == 0) {
        for (int i = 0; i < kEtWg; ++i) mbar_init(&s_bar[i], 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&s_tmem, 512);
    // B1[n][k]: n < 32 -> W2b column n (w2 rows [g, 2g)), n >= 32 -> W1a column n - 32 (w1 rows [0, g)); columns >= 24 of each half are zero
    for (int e = tid; e < kEtG * 64; e += kEtThreads) {
        const int k = e / 64, nn = e - k * 64;
        const int col = nn & 31;
        float w = 0.f;
        if (col < kEtG) w = nn < 32 ? __ldg(w2 + (size_t)(kEtG + k) * kEtG + col) : __ldg(w1 + (size_t)k * kEtG + col);
        const float hi = tf32_round(w);
        const uint32_t o = op_offset(64, nn, k);
        *reinterpret_cast<float *>(et_smem + o) = hi;
        *reinterpret_cast<float *>(et_smem + kEtB1Plane + o) = tf32_round(__fsub_rn(w, hi));
    }
    // B2[n][k] = W2a (w2 rows [0, g)) column n
    for (int e = tid; e < kEtG * 32; e += kEtThreads) {
        const int k = e / 32, nn = e - k * 32;
        const float w = nn < kEtG ? __ldg(w2 + (size_t)k * kEtG + nn) : 0.f;
        const float hi = tf32_round(w);
        const uint32_t o = op_offset(32, nn, k);
        *reinterpret_cast<float *>(et_smem + 2 * kEtB1Plane + o) = hi;
        *reinterpret_cast<float *>(et_smem + 2 * kEtB1Plane + kEtB2Plane + o) = tf32_round(__fsub_rn(w, hi));
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    const uint32_t tmem_wg = s_tmem + (uint32_t)(wg * kEtWgCols);
    const uint32_t tmem_row = tmem_wg + ((uint32_t)(q * 32) << 16);
    const uint32_t d_col = 64;
    uint32_t phase = 0;
    const int lm = lane & 15;
    const int tile_stride = gridDim.x * kEtWg;

    // neighbour index of my edge in the NEXT tile, fetched a tile ahead (the gather's row address depends on it)
    int j_next = 0;
    {
        const int tile0 = blockIdx.x * kEtWg + wg;
        const long long p0 = (long long)tile0 * 8 + (t >> 4);
        if (tile0 < n_tiles) j_next = __ldg(idx + (p0 < rows ? p0 : rows - 1) * kEtK + (t & 15));
    }
    int dbg_i = 0;
    const bool tr = dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define ET_MARK() do { if (tr && dbg_i < 60) dbg[dbg_i++] = clock64(); } while (0)
    ET_MARK();
    for (int tile = blockIdx.x * kEtWg + wg; tile < n_tiles; tile += tile_stride) {
        ET_MARK();
        const long long p = (long long)tile * 8 + (t >> 4);  // my point
        const bool pvalid = p < rows;
        const long long pc_ = pvalid ? p : rows - 1;
        const long long cloud = pc_ / n;
        const float *prow = proj + pc_ * kEtProjW;
        float *orow = out + pc_ * (long long)ldo;
        // operands of the LAST epilogue, requested now so that their latency hides behind the whole tile: my two columns of
        // the per-point term t2 and my share of the centre point's features (the tail of the output row)
        const float t2a = __ldg(prow + 3 * kEtG + lm), t2b = lm < 8 ? __ldg(prow + 3 * kEtG + 16 + lm) : 0.f;
        float fcp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) fcp[u] = (lm + 16 * u) < c ? __ldg(feat + pc_ * (long long)ldf + lm + 16 * u) : 0.f;
        // ---- y0 = relu(pn[j] + pc[i]) -> A operand; max over the point's edges -> out[48:72)
        float y[kEtG];
        {
            const int j = j_next;
            if (tile + tile_stride < n_tiles) {
                const long long pn_ = (long long)(tile + tile_stride) * 8 + (t >> 4);
                j_next = __ldg(idx + (pn_ < rows ? pn_ : rows - 1) * kEtK + (t & 15));
            }
            float pcv[kEtG];
            et_load24(proj + (cloud * n + j) * kEtProjW, y);
            et_load24(prow + kEtG, pcv);
#pragma unroll
            for (int o = 0; o < kEtG; ++o) y[o] = fmaxf(__fadd_rn(y[o], pcv[o]), 0.f);
        }
        ET_MARK();
        et_store_a(tmem_row, y);
        {
            float k0, k1;
            et_pool24(y, lm, k0, k1);
            if (pvalid) {
                orow[2 * kEtG + lm] = k0;
                if (lm < 8) orow[2 * kEtG + 16 + lm] = k1;
            }
        }
        ET_MARK();
        tmem_st_wait();
        tc_fence_before_sync();
        et_wg_barrier(wg);
        ET_MARK();
        // ---- D[:, 0:32 | 32:64] = y0 . [W2b | W1a]
        if (t == 0) {
            tc_fence_after_sync();
            const uint32_t idesc = make_idesc_tf32(128, 64);
            const uint32_t lbo = (64 / 8) * 128u;
            const uint32_t bh0 = smem_base, bl0 = smem_base + kEtB1Plane;
#pragma unroll
            for (int ks = 0; ks < kEtG / 8; ++ks) {
                const uint64_t bh = make_smem_desc(bh0 + ks * 2 * lbo, lbo, 128), bl = make_smem_desc(bl0 + ks * 2 * lbo, lbo, 128);
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + 32u + ks * 8, bh, idesc, ks > 0);  // a_lo * b_hi
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + ks * 8, bl, idesc, 1);             // a_hi * b_lo
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + ks * 8, bh, idesc, 1);             // a_hi * b_hi
            }
            mma_commit(&s_bar[wg]);
        }
        ET_MARK();
        if (t < 32) mbar_wait(&s_bar[wg], phase);
        phase ^= 1;
        ET_MARK();
        et_wg_barrier(wg);
        tc_fence_after_sync();
        ET_MARK();
        // ---- y1 = relu(D[:, 32:56] + t1[i]) -> A operand; max -> out[24:48)
        {
            uint32_t v[32];
            tmem_ld32(tmem_row + d_col + 32u, v);
            float t1[kEtG];
            et_load24(prow + 2 * kEtG, t1);
            tmem_ld_wait();
#pragma unroll
            for (int o = 0; o < kEtG; ++o) y[o] = fmaxf(__fadd_rn(__uint_as_float(v[o]), t1[o]), 0.f);
        }
        et_store_a(tmem_row, y);
        {
            float k0, k1;
            et_pool24(y, lm, k0, k1);
            if (pvalid) {
                orow[kEtG + lm] = k0;
                if (lm < 8) orow[kEtG + 16 + lm] = k1;
            }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        et_wg_barrier(wg);
        ET_MARK();
        // ---- D[:, 0:32] += y1 . W2a
        if (t == 0) {
            tc_fence_after_sync();
            const uint32_t idesc = make_idesc_tf32(128, 32);
            const uint32_t lbo = (32 / 8) * 128u;
            const uint32_t bh0 = smem_base + 2 * kEtB1Plane, bl0 = bh0 + kEtB2Plane;
#pragma unroll
            for (int ks = 0; ks < kEtG / 8; ++ks) {
                const uint64_t bh = make_smem_desc(bh0 + ks * 2 * lbo, lbo, 128), bl = make_smem_desc(bl0 + ks * 2 * lbo, lbo, 128);
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + 32u + ks * 8, bh, idesc, 1);
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + ks * 8, bl, idesc, 1);
                mma_tf32_ts(tmem_wg + d_col, tmem_wg + ks * 8, bh, idesc, 1);
            }
            mma_commit(&s_bar[wg]);
        }
        if (t < 32) mbar_wait(&s_bar[wg], phase);
        phase ^= 1;
        et_wg_barrier(wg);
        tc_fence_after_sync();
        ET_MARK();
        // ---- y2 = D[:, 0:24] + t2[i] (no activation); max -> out[0:24): the per-point term commutes with the max
        {
            uint32_t v[32];
            tmem_ld32(tmem_row + d_col, v);
            tmem_ld_wait();
#pragma unroll
            for (int o = 0; o < kEtG; ++o) y[o] = __uint_as_float(v[o]);
            float k0, k1;
            et_pool24(y, lm, k0, k1);
            if (pvalid) {
                orow[lm] = __fadd_rn(k0, t2a);
                if (lm < 8) orow[16 + lm] = __fadd_rn(k1, t2b);
                // the centre point's own features close the concat (Common/ops.py:1905-1913)
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (lm + 16 * u < c) orow[3 * kEtG + lm + 16 * u] = fcp[u];
            }
        }
        ET_MARK();
        // the next tile's A-operand stores touch only my own lanes, after my own tcgen05.wait::ld; its first MMA group
        // overwrites D only after the warpgroup barrier that follows those stores
    }
#undef ET_MARK
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(s_tmem, 512);
}

static long long *g_edge_dbg = nullptr;

// rows = b*n points, k = 16 neighbours, growth 24.  Returns false when the shape is not the one this kernel is built for.
bool edge_mlp_tc_launch(long long rows, int n, int c, int k, int growth, const float *proj, const int *idx, const float *w1,
                        const float *w2, const float *feat, int ldf, float *out, int ldo, cudaStream_t stream, int *rc) {
    static const bool off = getenv("DISPU_EDGE_SIMT") != nullptr;  // development switch (A/B timing against the FFMA2 kernel)
    if (off || k != kEtK || growth != kEtG) return false;
    const long long n_tiles_ll = ceil_div64(rows, 8);
    if (n_tiles_ll > 0x7fffffffLL) return false;
    const int n_tiles = (int)n_tiles_ll;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = ceil_div64(n_tiles, kEtWg);
    if (grid > sms) grid = sms;
    const cudaError_t e = cudaFuncSetAttribute(edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEtSmem);
    if (e != cudaSuccess) { *rc = (int)e; return true; }
    static long long *dbg = nullptr;  // development only (DISPU_EDGE_TRACE): clock64 samples of CTA 0, thread 0
    if (getenv("DISPU_EDGE_TRACE") && !dbg) { cudaMalloc(&dbg, 64 * sizeof(long long)); g_edge_dbg = dbg; }
    if (dbg) cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), stream);
    count_launch();
    edge_mlp_tc_kernel<<<(unsigned)grid, kEtThreads, kEtSmem, stream>>>(rows, n, c, proj, idx, w1, w2, feat, ldf, out, ldo, n_tiles, dbg);
    *rc = finish_launch();
    return true;
}

}  // namespace dispu

// development only: copy the clock64 samples of the last traced launch to the host (64 int64)
extern "C" int dispu__edge_debug_trace(long long *host_out) {
    if (!dispu::g_edge_dbg || !host_out) return DISPU_EUNSUPPORTED;
    return (int)cudaMemcpy(host_out, dispu::g_edge_dbg, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
}

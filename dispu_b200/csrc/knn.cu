// Brute-force k-nearest-neighbour kernels in 3-D for sm_100a.
//
//  * dispu_knn      — kNN grouping: replaces the nanoflann KD-tree round trip the live generator
//                     makes every step (libs/nearest_neighbors/knn_.cxx:104-135 through
//                     tf.py_func, Common/ops.py:110-118,165: device -> host -> KD-tree build per
//                     cloud -> host -> device) and the pure-TF knn_point
//                     (tf_ops/grouping/tf_grouping.py:116-141), which materialises a (b,m,n)
//                     distance tensor before a full top_k.
//  * dispu_three_nn — ThreeNN: replaces the single-threaded host loop threenn_cpu
//                     (tf_ops/interpolation/tf_interpolate.cpp:60-103); the reference registers
//                     no GPU kernel for it.
//
// Both: the dataset cloud is staged in shared memory as float4 tiles (one broadcast LDS.128 per
// candidate per warp), one thread owns one query and keeps its running top-k sorted in registers
// (statically indexed, strict '<' insertion so equal distances keep the lower index first — the
// order of tf.nn.top_k and of the reference's cascades).  Distances use the unfused host-loop
// contraction (d2_ref_cpu) so they are bit-identical to nanoflann's / threenn_cpu's floats.
#include <stdlib.h>

#include "common.cuh"
#include "tile_stage.cuh"

namespace dispu {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 2048;  // float4 points per tile (32 KB)

template <int KT, bool kTma>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
           float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKnnTile];
    extern __shared__ __align__(16) float s_raw[];  // kTma: one or two 24 KB landing buffers (cp.async.bulk)
    __shared__ __align__(8) uint64_t s_bar[2];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kKnnThreads + threadIdx.x;
    const float *pts = data + (size_t)cloud * n * 3;
    const bool active = j < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *q = query + ((size_t)cloud * m + j) * 3;
        qx = __ldg(q); qy = __ldg(q + 1); qz = __ldg(q + 2);
    }
    float bd[KT];
    int bi[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        bd[i] = __int_as_float(0x7f800000);  // +inf
        bi[i] = 0;
    }
    const int ntiles = ceil_div(n, kKnnTile);
    ts::stage_begin<kTma>(s_raw, s_bar, pts, n, kKnnTile);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int t0 = ti * kKnnTile;
        const int tn = ts::stage_tile<kTma>(s_tile, s_raw, s_bar, pts, n, kKnnTile, ti, ntiles);
        if (active) {
            for (int e = 0; e < tn; ++e) {
                const float4 p = s_tile[e];
                // diff = query - data, accumulated dimension by dimension (nanoflann.hpp:323-348)
                const float d = d2_ref_cpu(__fsub_rn(qx, p.x), __fsub_rn(qy, p.y), __fsub_rn(qz, p.z));
                if (d < bd[KT - 1]) {
                    bd[KT - 1] = d;
                    bi[KT - 1] = t0 + e;
#pragma unroll
                    for (int i = KT - 1; i > 0; --i) {
                        if (bd[i] < bd[i - 1]) {
                            const float tf = bd[i]; bd[i] = bd[i - 1]; bd[i - 1] = tf;
                            const int ti = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (active) {
        const size_t o = ((size_t)cloud * m + j) * k;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            if (i < k) {
                if (dist) dist[o + i] = bd[i];
                idx[o + i] = bi[i];
            }
        }
    }
}


// ---- selection kernel (k <= 32): no insertion in the common case -------------------------------------------
// A warp owns one query and takes the dataset in chunks of 1024 candidates, 32 per lane, whose squared
// distances stay in registers.  Threshold by LANE MINIMA: sort the 32 per-lane minima across the warp
// (bitonic, shuffles); the k-th smallest of them, M, bounds the k-th smallest distance of the chunk from
// above (k different lanes each hold a value <= M).  Every candidate with d <= tau = min(M, carried k-th
// best) is compacted into a small shared-memory buffer by ballot (about 32(H32-H16) ~ 22 survivors for
// k = 16), the carried list joins it, and the survivors are ordered by rank-counting on the (d, idx) key —
// ties to the lower index, the order of tf.nn.top_k.  A query whose survivors overflow the
// buffer (masses of exactly equal distances) is redone with the ballot-ranked insertion scan.
constexpr int kKsWarps = 8;
constexpr int kKsChunk = 1024;
constexpr int kKsQueriesPerWarp = 4;
constexpr int kKsCap = 128;

__device__ __forceinline__ float warp_sort32_asc(float v, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(kFull, v, j);
            const bool up = (lane & k2) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}

// Slow exact path for one query: serial definition with a lane-distributed sorted list (32 entries).
__device__ __noinline__ void knn_insertion_scan(int n, int k, const float *__restrict__ pts, float qx, float qy, float qz,
                                                int lane, float &ld, int &li) {
    const float inf = __int_as_float(0x7f800000);
    ld = inf; li = 0;
    float tau = inf;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + lane;
        float d = inf;
        if (c < n) d = d2_ref_cpu(__fsub_rn(qx, __ldg(pts + (size_t)c * 3)), __fsub_rn(qy, __ldg(pts + (size_t)c * 3 + 1)),
                                  __fsub_rn(qz, __ldg(pts + (size_t)c * 3 + 2)));
        unsigned vote = __ballot_sync(kFull, c < n && d < tau);
        while (vote) {
            const int src = __ffs(vote) - 1;
            vote &= vote - 1;
            const float dn = __shfl_sync(kFull, d, src);
            if (!(dn < tau)) continue;
            const int pos = __popc(__ballot_sync(kFull, ld <= dn));
            const float pd = __shfl_up_sync(kFull, ld, 1);
            const int pi = __shfl_up_sync(kFull, li, 1);
            if (lane > pos) { ld = pd; li = pi; }
            if (lane == pos) { ld = dn; li = c0 + src; }
            tau = __shfl_sync(kFull, ld, k - 1);
        }
    }
}

// __launch_bounds__(.., 4): 64 registers per thread (the compiler settles at 128 otherwise) -> four CTAs = 32 warps per SM;
// the kernel is a chain of dependent shuffles and shared-memory round trips, so it needs the extra warps to hide them.
template <bool kTma>
__global__ void __launch_bounds__(kKsWarps * 32, 4)
knn_select_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
                  float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKsChunk];
    extern __shared__ __align__(16) float s_raw[];  // kTma: one or two 12 KB landing buffers (cp.async.bulk)
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float s_bd[kKsWarps][kKsCap];
    __shared__ int s_bi[kKsWarps][kKsCap];
    __shared__ float s_od[kKsWarps][32];
    __shared__ int s_oi[kKsWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    const float *pts = data + (size_t)cloud * n * 3;
    const int q0 = (blockIdx.x * kKsWarps + warp) * kKsQueriesPerWarp;
    const float inf = __int_as_float(0x7f800000);
    const unsigned lt_mask = (1u << lane) - 1u;
    float *bd = s_bd[warp];
    int *bi = s_bi[warp];

    float ld[kKsQueriesPerWarp];  // carried list: lane j holds the j-th best so far
    int li[kKsQueriesPerWarp];
    bool overflow[kKsQueriesPerWarp];
    float qx[kKsQueriesPerWarp], qy[kKsQueriesPerWarp], qz[kKsQueriesPerWarp];
#pragma unroll
    for (int u = 0; u < kKsQueriesPerWarp; ++u) {
        ld[u] = inf; li[u] = 0; overflow[u] = false;
        const int j = min(q0 + u, m - 1);
        const float *q = query + ((size_t)cloud * m + j) * 3;
        qx[u] = __ldg(q); qy[u] = __ldg(q + 1); qz[u] = __ldg(q + 2);
    }
    const int ntiles = ceil_div(n, kKsChunk);
    ts::stage_begin<kTma>(s_raw, s_bar, pts, n, kKsChunk);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int t0 = ti * kKsChunk;
        const int tn = ts::stage_tile<kTma>(s_tile, s_raw, s_bar, pts, n, kKsChunk, ti, ntiles);
        if (q0 >= m) continue;
#pragma unroll 1
        for (int u = 0; u < kKsQueriesPerWarp; ++u) {
            if (overflow[u]) continue;
            if (t0 > 0) {
                // later chunks: the carried list is already tight (expected k*ln(n/1024) improvements in total),
                // so a threshold-gated step per 32 candidates with ballot-ranked insertion is cheaper than
                // another selection round
                float tau_u = __shfl_sync(kFull, ld[u], k - 1);
                for (int c0 = 0; c0 < tn; c0 += 32) {
                    const int c = c0 + lane;
                    float dv = inf;
                    if (c < tn) {
                        const float4 p = s_tile[c];
                        dv = d2_ref_cpu(__fsub_rn(qx[u], p.x), __fsub_rn(qy[u], p.y), __fsub_rn(qz[u], p.z));
                    }
                    unsigned vote = __ballot_sync(kFull, dv < tau_u);
                    while (vote) {
                        const int src = __ffs(vote) - 1;
                        vote &= vote - 1;
                        const float dn = __shfl_sync(kFull, dv, src);
                        if (!(dn < tau_u)) continue;
                        const int pos = __popc(__ballot_sync(kFull, ld[u] <= dn));
                        const float pd = __shfl_up_sync(kFull, ld[u], 1);
                        const int pi = __shfl_up_sync(kFull, li[u], 1);
                        if (lane > pos) { ld[u] = pd; li[u] = pi; }
                        if (lane == pos) { ld[u] = dn; li[u] = t0 + c0 + src; }
                        tau_u = __shfl_sync(kFull, ld[u], k - 1);
                    }
                }
                continue;
            }
            // 32 candidates per lane, distances in registers.  Lane `lane` takes candidate c = s*32 + (lane ^ s) of
            // slot s: the permutation by s spreads index-periodic structure over the lanes (the generator's coarse cloud is
            // 4 near-copies of each input point at indices i, i+n/4, ...: unrotated, all copies of a point fall into ONE
            // lane, the lane minima stop being a useful threshold and the survivor list balloons: 147 vs 84 us at c2);
            // still one conflict-free LDS.128 per slot (a permutation of the 32 consecutive float4s)
            float d[32];
            float lmin = inf;
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) {
                const int c = sI * 32 + (lane ^ sI);
                float v = inf;
                if (c < tn) {
                    const float4 p = s_tile[c];
                    // diff = query - data, accumulated dimension by dimension (nanoflann.hpp:323-348)
                    v = d2_ref_cpu(__fsub_rn(qx[u], p.x), __fsub_rn(qy[u], p.y), __fsub_rn(qz[u], p.z));
                }
                d[sI] = v;
                lmin = fminf(lmin, v);
            }
            const float sorted_min = warp_sort32_asc(lmin, lane);
            const float M = __shfl_sync(kFull, sorted_min, k - 1);
            const float tau_carry = __shfl_sync(kFull, ld[u], k - 1);
            const float tau = fminf(M, tau_carry);
            // survivors: carried entries first, then this chunk's candidates, all with d <= tau
            int S = 0;
            {
                const bool keep = ld[u] <= tau && ld[u] < inf;
                const unsigned mask = __ballot_sync(kFull, keep);
                if (keep) { const int pos = __popc(mask & lt_mask); bd[pos] = ld[u]; bi[pos] = li[u]; }
                S = __popc(mask);
            }
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) {
                const bool keep = d[sI] <= tau && d[sI] < inf;
                const unsigned mask = __ballot_sync(kFull, keep);
                if (mask) {
                    const int pos = S + __popc(mask & lt_mask);
                    if (keep && pos < kKsCap) { bd[pos] = d[sI]; bi[pos] = t0 + sI * 32 + (lane ^ sI); }
                    S += __popc(mask);
                }
            }
            __syncwarp();
            if (S > kKsCap) { overflow[u] = true; continue; }
            // rank by counting on the (d, idx) key; ranks < 32 become the new carried list
            for (int e0 = 0; e0 < S; e0 += 32) {
                const int e = e0 + lane;
                const float de = e < S ? bd[e] : inf;
                const int ie = e < S ? bi[e] : 0x7fffffff;
                int rank = 0;
                for (int f = 0; f < S; ++f) {
                    const float df = bd[f];
                    const int jf = bi[f];
                    rank += (df < de) || (df == de && jf < ie);
                }
                if (e < S && rank < 32) {  // ranks are a permutation of 0..S-1: hand element e to lane `rank`
                    s_od[warp][rank] = de;
                    s_oi[warp][rank] = ie;
                }
            }
            __syncwarp();
            const int have = min(S, 32);
            ld[u] = lane < have ? s_od[warp][lane] : inf;
            li[u] = lane < have ? s_oi[warp][lane] : 0;
            __syncwarp();
        }
    }
#pragma unroll 1
    for (int u = 0; u < kKsQueriesPerWarp; ++u) {
        const int j = q0 + u;
        if (j >= m) continue;
        if (overflow[u]) knn_insertion_scan(n, k, pts, qx[u], qy[u], qz[u], lane, ld[u], li[u]);
        if (lane < k) {
            const size_t o = ((size_t)cloud * m + j) * k + lane;
            if (dist) dist[o] = ld[u];
            idx[o] = li[u];
        }
    }
}

// ---- single-chunk form (n <= 1024, the generator's clouds): everything of a query lives in registers -------------------
// The generic kernel above carries per-query state across chunks in arrays indexed by the query loop (local memory at
// 64 registers) and compacts survivors with one ballot per candidate slot; an instruction-level profile (ncu source page,
// profiles/r02_step_others_ncu.md: issue-bound at 76 %, 1618 warp instructions per query) put 40 % of the work into the
// distance loop's index / guard / local-memory overhead and 27 % into the 32 ballots.  With the whole cloud in ONE
// chunk there is no carried list: the query is three scalars, a lane marks its survivors in a 32-bit mask (one compare
// per slot), a warp prefix sum over the 32 counts gives every lane its write offset, and the ranked result goes straight
// to global memory.  Same selection rule and the same fallback as the generic kernel.
// packed FP32 (one issue slot, two IEEE results: bit-identical to the scalar forms)
typedef unsigned long long ks_f32x2;
__device__ __forceinline__ ks_f32x2 ks_pack(float lo, float hi) {
    ks_f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ks_unpack(ks_f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ ks_f32x2 ks_sub2(ks_f32x2 a, ks_f32x2 b) {
    ks_f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ ks_f32x2 ks_mul2(ks_f32x2 a, ks_f32x2 b) {
    ks_f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <bool kTma>
__global__ void __launch_bounds__(kKsWarps * 32, 4)
knn_select1_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
                   float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKsChunk];
    extern __shared__ __align__(16) float s_raw[];  // kTma: one 12 KB landing buffer (cp.async.bulk)
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ unsigned long long s_bk[kKsWarps][kKsCap];  // survivors: (distance bits : index) keys
    __shared__ float s_od[kKsWarps][32];
    __shared__ int s_oi[kKsWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    const float *pts = data + (size_t)cloud * n * 3;
    const int q0 = (blockIdx.x * kKsWarps + warp) * kKsQueriesPerWarp;
    const float inf = __int_as_float(0x7f800000);
    unsigned long long *bk = s_bk[warp];
    ts::stage_begin<kTma>(s_raw, s_bar, pts, n, kKsChunk);
    const int tn = ts::stage_tile<kTma>(s_tile, s_raw, s_bar, pts, n, kKsChunk, 0, 1);
    // Full chunk: a second layout for the distance loop — slot pair (2h, 2h + 1) of lane l as three packed pairs
    // (x, x'), (y, y'), (z, z'), so that TWO candidates cost three 8-byte loads, six packed FP32 instructions (sub / mul
    // .f32x2, IEEE per element: the same bits as the scalar form) and four scalar additions instead of two 16-byte loads
    // and sixteen scalar instructions.  Candidate of slot s for lane l is c = s*32 + (l ^ s), as below.
    __shared__ float2 s_px[16 * 32], s_py[16 * 32], s_pz[16 * 32];
    if (tn == kKsChunk) {
        for (int e = threadIdx.x; e < 16 * 32; e += kKsWarps * 32) {
            const int h = e >> 5, l = e & 31;
            const float4 a = s_tile[(2 * h) * 32 + (l ^ (2 * h))], b = s_tile[(2 * h + 1) * 32 + (l ^ (2 * h + 1))];
            s_px[e] = make_float2(a.x, b.x);
            s_py[e] = make_float2(a.y, b.y);
            s_pz[e] = make_float2(a.z, b.z);
        }
        __syncthreads();
    }
    if (q0 >= m) return;
#pragma unroll 1
    for (int u = 0; u < kKsQueriesPerWarp; ++u) {
        const int j = q0 + u;
        if (j >= m) break;
        const float *q = query + ((size_t)cloud * m + j) * 3;
        const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
        // candidate of slot s for this lane: c = s*32 + (lane ^ s) (see the generic kernel: spreads index-periodic clouds)
        float d[32];
        float lmin = inf;
        if (tn == kKsChunk) {
            const ks_f32x2 qx2 = ks_pack(qx, qx), qy2 = ks_pack(qy, qy), qz2 = ks_pack(qz, qz);
#pragma unroll
            for (int h = 0; h < 16; ++h) {
                // diff = query - data, ((dx*dx + dy*dy) + dz*dz) dimension by dimension (nanoflann.hpp:323-348), two candidates at once
                const ks_f32x2 dx = ks_sub2(qx2, *reinterpret_cast<const ks_f32x2 *>(&s_px[h * 32 + lane]));
                const ks_f32x2 dy = ks_sub2(qy2, *reinterpret_cast<const ks_f32x2 *>(&s_py[h * 32 + lane]));
                const ks_f32x2 dz = ks_sub2(qz2, *reinterpret_cast<const ks_f32x2 *>(&s_pz[h * 32 + lane]));
                // (the additions stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which would change the bits)
                float xx0, xx1, yy0, yy1, zz0, zz1;
                ks_unpack(ks_mul2(dx, dx), xx0, xx1);
                ks_unpack(ks_mul2(dy, dy), yy0, yy1);
                ks_unpack(ks_mul2(dz, dz), zz0, zz1);
                d[2 * h] = __fadd_rn(__fadd_rn(xx0, yy0), zz0);
                d[2 * h + 1] = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
                lmin = fminf(lmin, fminf(d[2 * h], d[2 * h + 1]));
            }
        } else {
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) {
                const int c = sI * 32 + (lane ^ sI);
                const float4 p = s_tile[c < tn ? c : 0];
                const float v = d2_ref_cpu(__fsub_rn(qx, p.x), __fsub_rn(qy, p.y), __fsub_rn(qz, p.z));
                d[sI] = c < tn ? v : inf;
                lmin = fminf(lmin, d[sI]);
            }
        }
        const float tau = __shfl_sync(kFull, warp_sort32_asc(lmin, lane), k - 1);
        // my survivors as a bit mask, then a warp prefix sum of the counts
        unsigned mine = 0;
        if (tn == kKsChunk) {  // no padding slots: every distance is finite (or NaN, which fails the comparison)
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) mine |= (d[sI] <= tau) ? (1u << sI) : 0u;
        } else {
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) mine |= (d[sI] <= tau && d[sI] < inf) ? (1u << sI) : 0u;
        }
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int S = __shfl_sync(kFull, incl, 31);
        float rd = inf;
        int ri = 0;
        if (S > kKsCap) {  // masses of exactly equal distances: the serial definition
            knn_insertion_scan(n, k, pts, qx, qy, qz, lane, rd, ri);
        } else {
            // survivors as ONE 64-bit key (distance bits : index): squared distances are >= +0, so their bit patterns order
            // like the values and a single unsigned comparison is the (d, idx) lexicographic order — ties to the lower index
            int pos = incl - cnt;
#pragma unroll
            for (int sI = 0; sI < 32; ++sI) {
                if (mine & (1u << sI)) {
                    bk[pos] = ((unsigned long long)__float_as_uint(d[sI]) << 32) | (unsigned)(sI * 32 + (lane ^ sI));
                    ++pos;
                }
            }
            __syncwarp();
            // rank by counting; ranks < 32 are the answer, ascending
            for (int e0 = 0; e0 < S; e0 += 32) {
                const int e = e0 + lane;
                const unsigned long long ke = e < S ? bk[e] : ~0ull;
                int rank = 0;
#pragma unroll 4
                for (int f = 0; f < S; ++f) rank += bk[f] < ke;
                if (e < S && rank < 32) {
                    s_od[warp][rank] = __uint_as_float((unsigned)(ke >> 32));
                    s_oi[warp][rank] = (int)(unsigned)ke;
                }
            }
            __syncwarp();
            const int have = min(S, 32);
            rd = lane < have ? s_od[warp][lane] : inf;
            ri = lane < have ? s_oi[warp][lane] : 0;
            __syncwarp();
        }
        if (lane < k) {
            const size_t o = ((size_t)cloud * m + j) * k + lane;
            if (dist) dist[o] = rd;
            idx[o] = ri;
        }
    }
}

// ---- warp-per-query kernel (k <= 32) -----------------------------------------------------------------
// The thread-per-query kernel above pays its 75-instruction sorted insertion on every candidate for which
// ANY of a warp's 32 queries improves — i.e. almost always.  Here a warp owns ONE query: each lane
// tests one candidate per step against the warp-uniform threshold tau (= current k-th best), so the
// common step is 13 instructions and an insertion happens only for a genuine improvement
// (~k(1+ln(n/k)) times per query).  The running list is distributed over the lanes (lane j holds the
// j-th best (d, idx), ascending, 32 entries), an insertion is a ballot-ranked shift by one lane.
// Candidates are taken in ascending index inside a step (lowest lane first) and only on strict
// improvement, so equal distances keep the lower index first — the order of tf.nn.top_k.
constexpr int kKwWarps = 8;
constexpr int kKwTile = 2048;        // float4 candidates per shared-memory tile
constexpr int kKwQueriesPerWarp = 4; // queries a warp processes against each tile load

__global__ void __launch_bounds__(kKwWarps * 32)
knn_warp_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
                float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKwTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    const float *pts = data + (size_t)cloud * n * 3;
    const int q0 = (blockIdx.x * kKwWarps + warp) * kKwQueriesPerWarp;
    const float inf = __int_as_float(0x7f800000);

    float ld[kKwQueriesPerWarp];
    int li[kKwQueriesPerWarp];
    float tau[kKwQueriesPerWarp], qx[kKwQueriesPerWarp], qy[kKwQueriesPerWarp], qz[kKwQueriesPerWarp];
#pragma unroll
    for (int u = 0; u < kKwQueriesPerWarp; ++u) {
        ld[u] = inf; li[u] = 0; tau[u] = inf;
        const int j = min(q0 + u, m - 1);
        const float *q = query + ((size_t)cloud * m + j) * 3;
        qx[u] = __ldg(q); qy[u] = __ldg(q + 1); qz[u] = __ldg(q + 2);
    }
    for (int t0 = 0; t0 < n; t0 += kKwTile) {
        const int tn = min(kKwTile, n - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < tn; e += kKwWarps * 32) {
            const float *p = pts + (size_t)(t0 + e) * 3;
            s_tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (q0 < m) {
            for (int c0 = 0; c0 < tn; c0 += 32) {
                const int c = c0 + lane;
                const bool cvalid = c < tn;
                const float4 p = s_tile[cvalid ? c : 0];
#pragma unroll
                for (int u = 0; u < kKwQueriesPerWarp; ++u) {
                    // diff = query - data, accumulated dimension by dimension (nanoflann.hpp:323-348)
                    const float d = d2_ref_cpu(__fsub_rn(qx[u], p.x), __fsub_rn(qy[u], p.y), __fsub_rn(qz[u], p.z));
                    unsigned vote = __ballot_sync(kFull, cvalid && d < tau[u]);
                    while (vote) {
                        const int src = __ffs(vote) - 1;
                        vote &= vote - 1;
                        const float dn = __shfl_sync(kFull, d, src);
                        if (!(dn < tau[u])) continue;  // tau shrank since the vote
                        const int in = t0 + c0 + src;
                        const int pos = __popc(__ballot_sync(kFull, ld[u] <= dn));  // equal distances: earlier index stays first
                        const float pd = __shfl_up_sync(kFull, ld[u], 1);
                        const int pi = __shfl_up_sync(kFull, li[u], 1);
                        if (lane > pos) { ld[u] = pd; li[u] = pi; }
                        if (lane == pos) { ld[u] = dn; li[u] = in; }
                        tau[u] = __shfl_sync(kFull, ld[u], k - 1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kKwQueriesPerWarp; ++u) {
        const int j = q0 + u;
        if (j < m && lane < k) {
            const size_t o = ((size_t)cloud * m + j) * k + lane;
            if (dist) dist[o] = ld[u];
            idx[o] = li[u];
        }
    }
}

// The KT-slot list holds the KT smallest; the first k of them are the k smallest, so running
// with KT >= k is exact.
// dynamic shared memory of a kTma launch: two landing buffers when the cloud spans several tiles, else one
static size_t landing_bytes(int n, int tile_pts) { return (size_t)(n > tile_pts ? 2 : 1) * tile_pts * 12; }

static bool knn_use_tma(const float *data, int n) {
    static const bool no_tma = getenv("DISPU_NO_TMA") != nullptr;  // development switch (A/B timing)
    return !no_tma && ts::tma_ok(data, n);
}

template <int KT>
static int launch_knn(int b, int n, int m, int k, const float *data, const float *query, float *dist, int *idx,
                      cudaStream_t stream) {
    const dim3 grid(ceil_div(m, kKnnThreads), b);
    if (knn_use_tma(data, n)) {
        const size_t smem = landing_bytes(n, kKnnTile);
        const cudaError_t e = cudaFuncSetAttribute(knn_kernel<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kKnnTile * 12);
        if (e != cudaSuccess) return (int)e;
        count_launch();
        knn_kernel<KT, true><<<grid, kKnnThreads, smem, stream>>>(n, m, k, data, query, dist, idx);
    } else {
        count_launch();
        knn_kernel<KT, false><<<grid, kKnnThreads, 0, stream>>>(n, m, k, data, query, dist, idx);
    }
    return finish_launch();
}

// ThreeNN: (x2-x1) with x2 the known point, strict '<' cascades (tf_interpolate.cpp:77-91).
// The reference compares in double against 1e40; for float inputs that is the float order with a
// +inf start (a finite float is always < 1e40, inf/NaN never are), and float(1e40) = +inf is what
// it stores when fewer than three known points exist.
template <bool kTma>
__global__ void __launch_bounds__(kKnnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKnnTile];
    extern __shared__ __align__(16) float s_raw[];  // kTma: one or two 24 KB landing buffers (cp.async.bulk)
    __shared__ __align__(8) uint64_t s_bar[2];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kKnnThreads + threadIdx.x;
    const float *known = xyz2 + (size_t)cloud * m * 3;
    const bool active = j < n;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float *u = xyz1 + ((size_t)cloud * n + j) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    const float inf = __int_as_float(0x7f800000);
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = 0, i2 = 0, i3 = 0;
    const int ntiles = ceil_div(m, kKnnTile);
    ts::stage_begin<kTma>(s_raw, s_bar, known, m, kKnnTile);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int t0 = ti * kKnnTile;
        const int tn = ts::stage_tile<kTma>(s_tile, s_raw, s_bar, known, m, kKnnTile, ti, ntiles);
        if (active) {
            for (int e = 0; e < tn; ++e) {
                const float4 p = s_tile[e];
                const float d = d2_ref_cpu(__fsub_rn(p.x, ux), __fsub_rn(p.y, uy), __fsub_rn(p.z, uz));
                const int k = t0 + e;
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else if (d < b3) {
                    b3 = d; i3 = k;
                }
            }
        }
    }
    if (active) {
        const size_t o = ((size_t)cloud * n + j) * 3;
        dist[o] = b1; dist[o + 1] = b2; dist[o + 2] = b3;
        idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *dist, int *idx,
                         dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(k >= 1 && k <= n, DISPU_EINVAL);
    DISPU_REQUIRE(k <= 64, DISPU_EUNSUPPORTED);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && idx, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    if (k <= 32) {
        const dim3 grid(ceil_div(m, kKsWarps * kKsQueriesPerWarp), b);
        static const bool generic = getenv("DISPU_KNN_GENERIC") != nullptr;  // development switch (A/B timing)
        if (n <= kKsChunk && !generic) {  // whole cloud in one chunk: the register-resident form
            count_launch();
            if (knn_use_tma(xyz1, n)) {  // 39 KB static + the 12 KB landing buffer: above the 48 KB a kernel gets without opting in
                const cudaError_t e = cudaFuncSetAttribute(knn_select1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kKsChunk * 12);
                if (e != cudaSuccess) return (int)e;
                knn_select1_kernel<true><<<grid, kKsWarps * 32, (size_t)kKsChunk * 12, stream>>>(n, m, k, xyz1, xyz2, dist, idx);
            } else
                knn_select1_kernel<false><<<grid, kKsWarps * 32, 0, stream>>>(n, m, k, xyz1, xyz2, dist, idx);
            return finish_launch();
        }
        if (knn_use_tma(xyz1, n)) {
            const cudaError_t e = cudaFuncSetAttribute(knn_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kKsChunk * 12);
            if (e != cudaSuccess) return (int)e;
            count_launch();
            knn_select_kernel<true><<<grid, kKsWarps * 32, landing_bytes(n, kKsChunk), stream>>>(n, m, k, xyz1, xyz2, dist, idx);
        } else {
            count_launch();
            knn_select_kernel<false><<<grid, kKsWarps * 32, 0, stream>>>(n, m, k, xyz1, xyz2, dist, idx);
        }
        return finish_launch();
    }
    if (k <= 4) return launch_knn<4>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 8) return launch_knn<8>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 16) return launch_knn<16>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 24) return launch_knn<24>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 32) return launch_knn<32>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    return launch_knn<64>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
}

extern "C" int dispu_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                              dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n >= 0 && m > 0, DISPU_EINVAL);
    if (b == 0 || n == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && dist && idx, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    count_launch();
    if (knn_use_tma(xyz2, m)) {
        const cudaError_t e = cudaFuncSetAttribute(three_nn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kKnnTile * 12);
        if (e != cudaSuccess) return (int)e;
        three_nn_kernel<true><<<dim3(ceil_div(n, kKnnThreads), b), kKnnThreads, landing_bytes(m, kKnnTile), stream>>>(n, m, xyz1, xyz2, dist, idx);
    } else
        three_nn_kernel<false><<<dim3(ceil_div(n, kKnnThreads), b), kKnnThreads, 0, stream>>>(n, m, xyz1, xyz2, dist, idx);
    return finish_launch();
}

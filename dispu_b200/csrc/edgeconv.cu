// Feature-space kNN + EdgeConv dense block + weight-net aggregation for sm_100a.
//
//  * dispu_feature_knn — knn_point_2 (tf_ops/grouping/tf_grouping.py:95-114): k nearest neighbours in a
//    C-dimensional FEATURE space on the matmul-form distance matrix (batch_distance_matrix_general, :61-66)
//    D = (r_q - 2 q.p) + r_p, ascending, ties to the lower index (tf.nn.top_k).  The reference materialises the
//    (b,n,n) matrix, runs a full top_k and builds (b,n,k,2) gather_nd pairs; here a warp owns two queries, the
//    cloud's features sit transposed in shared memory (conflict-free column reads), distances stay in
//    registers and the k winners are selected by a lane-minima threshold + ballot compaction + rank counting (k rounds
//    of warp arg-min — redux.sync on an order-preserving key, then on the index — when exact ties overflow that path).  Arithmetic is pinned (sequential fmaf over the channels), so results are reproducible bit for bit.
//  * dispu_dense_conv — dense_conv (Common/ops.py:1897-1915) over get_edge_feature (:1856-1877): the three 1x1
//    convs on concat[central, neighbour-central | y | central] and the max over the k neighbours in ONE pass.
//    Everything that depends only on the centre point is hoisted to a per-point product (exact reassociation):
//        y0 = relu(pn[j] + pc[i])            pn = f.W0[c:2c]        pc = f.(W0[:c]-W0[c:]) + b0
//        y1 = relu(y0.W1[:g] + t1[i])        t1 = f.W1[g:] + b1
//        y2 =      y1.W2[:g] + y0.W2[g:2g] + t2[i]                  t2 = f.W2[2g:] + b2
//    so no (b,n,k,*) tensor exists anywhere; the per-edge 24x24 / 48x24 products run as FP32 FMA chains with the
//    weights broadcast from shared memory (two edges per thread share every weight load).
//  * dispu_weightnet_aggregate — PointShuffle2's weight net + transposed product (Common/ops.py:1062-1067,
//    weight_net_hidden :181-191): H[i] = transpose(G[i]) . relu(bn(gx.Ww + bw)) without the (b,m,ns,3),
//    (b,m,ns,ns) and transposed (b,m,C,ns) intermediates.
#include <stdlib.h>

#include "common.cuh"

namespace dispu {

constexpr int kEcG = 24;            // growth rate of the live generator (DisPU/generator.py:36)
constexpr int kEcProjW = 4 * kEcG;  // pn | pc | t1 | t2

// Optional tail of the feature-kNN kernel: the EdgeConv projection of the CTA's own 32 query points (see edge_proj_kernel).
struct EcProjArgs {
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    float *proj;  // (b*n, 96); nullptr = no projection
    int raw_async;  // weights are 16-byte aligned: their raw rows are fetched with cp.async at kernel start (off the critical path)
    long long *dbg;  // development only: clock64 samples of CTA (0,0), thread 0
};

// ------------------------------------------------------------------------------------------------------------------
// feature kNN
constexpr int kFkWarps = 8;
constexpr int kFkQueriesPerWarp = 4;  // processed two at a time

__device__ __forceinline__ unsigned ord_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float ord_key_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// ascending bitonic sort of one value per lane (15 shuffle steps)
__device__ __forceinline__ float fk_warp_sort32(float v, int lane) {
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

constexpr int kFkCap = 96;  // survivors of the threshold selection a warp can rank in shared memory

// P = candidates per lane per chunk (chunk = 32*P points staged transposed in shared memory).
template <int P>
__global__ void __launch_bounds__(kFkWarps * 32)
feature_knn_kernel(int n, int m, int c, int k, int skip, const float *__restrict__ points, int ldp,
                   const float *__restrict__ queries, int ldq, float *__restrict__ dist, int *__restrict__ idx,
                   const EcProjArgs pj) {
    extern __shared__ __align__(16) float fk_smem[];
    constexpr int kChunk = 32 * P;
    constexpr int kStride = kChunk + 1;  // odd stride: the transposing stores are conflict-free too
    constexpr int kQ = P <= 8 ? 4 : 2;   // queries whose distances are in registers at once
    float *s_ft = fk_smem;                           // [c][kStride]
    float *s_r = s_ft + (size_t)c * kStride;         // [kChunk]
    float *s_q = s_r + kChunk;                       // [warps][kFkQueriesPerWarp][c]
    float *s_sd = s_q + (size_t)kFkWarps * kFkQueriesPerWarp * c;                 // [warps][kFkCap] survivor keys (with s_si: 64-bit)
    s_sd += (reinterpret_cast<uintptr_t>(s_sd) >> 2) & 1;                         // 8-byte aligned for any c (2 spare floats allocated)
    int *s_si = reinterpret_cast<int *>(s_sd + kFkWarps * kFkCap);                // [warps][kFkCap] survivor indices
    float *s_od = reinterpret_cast<float *>(s_si + kFkWarps * kFkCap);            // [warps][32] ranked output
    int *s_oi = reinterpret_cast<int *>(s_od + kFkWarps * 32);
    float *s_raw = reinterpret_cast<float *>(s_oi + kFkWarps * 32);               // projection tail: raw W0 | W1[g:] | W2[2g:] rows (96 c floats)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    if (pj.proj != nullptr && pj.raw_async) {
        // the projection's weight rows travel to shared memory while the neighbour search runs: 16-byte cp.async, no registers
        const int n4 = 24 * c;  // float4s: W0 (2c x 24) | W1 rows g.. (c x 24) | W2 rows 2g.. (c x 24)
        for (int e = threadIdx.x; e < n4; e += kFkWarps * 32) {
            const float *src = e < 12 * c ? pj.w0 + 4 * e
                               : e < 18 * c ? pj.w1 + (size_t)kEcG * kEcG + 4 * (e - 12 * c)
                                            : pj.w2 + (size_t)2 * kEcG * kEcG + 4 * (e - 18 * c);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_raw + 4 * e)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const float *pts = points + (size_t)cloud * n * ldp;
    const float *qry = queries + (size_t)cloud * m * ldq;
    const int q0 = (blockIdx.x * kFkWarps + warp) * kFkQueriesPerWarp;
    const float inf = __int_as_float(0x7f800000);

    // carried lists: lane j holds the j-th best (d, idx) of query u so far
    float bd[kFkQueriesPerWarp];
    int bi[kFkQueriesPerWarp];
#pragma unroll
    for (int u = 0; u < kFkQueriesPerWarp; ++u) { bd[u] = inf; bi[u] = 0x7fffffff; }

    int dbg_i = 0;
    const bool tr = pj.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
#define FK_MARK() do { if (tr && dbg_i < 60) pj.dbg[dbg_i++] = clock64(); } while (0)
    FK_MARK();
    for (int t0 = 0; t0 < n; t0 += kChunk) {
        const int tn = min(kChunk, n - t0);
        __syncthreads();
        if (((c | ldp) & 3) == 0 && (reinterpret_cast<uintptr_t>(pts) & 15u) == 0) {
            // 16-byte loads, four in flight per thread; the transposing stores see at most 2-way conflicts
            const int c4 = c >> 2;
#pragma unroll 4
            for (int e = threadIdx.x; e < tn * c4; e += kFkWarps * 32) {
                const int p = e / c4, l4 = e - p * c4;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(pts + (size_t)(t0 + p) * ldp) + l4);
                float *dst = s_ft + (size_t)(4 * l4) * kStride + p;
                dst[0] = v.x; dst[kStride] = v.y; dst[2 * kStride] = v.z; dst[3 * kStride] = v.w;
            }
        } else {
            for (int e = threadIdx.x; e < tn * c; e += kFkWarps * 32) {
                const int p = e / c, l = e - p * c;
                s_ft[(size_t)l * kStride + p] = __ldg(pts + (size_t)(t0 + p) * ldp + l);
            }
        }
        __syncthreads();
        FK_MARK();
        for (int p = threadIdx.x; p < kChunk; p += kFkWarps * 32) {
            float r = 0.f;
            if (p < tn)
                for (int l = 0; l < c; ++l) { const float v = s_ft[(size_t)l * kStride + p]; r = __fmaf_rn(v, v, r); }
            s_r[p] = r;
        }
        __syncthreads();
        FK_MARK();
        if (q0 >= m) continue;
#pragma unroll
        for (int u0 = 0; u0 < kFkQueriesPerWarp; u0 += kQ) {
            if (q0 + u0 >= m) break;
            // stage the kQ query vectors (rows past m clamp to a valid row; their results are discarded)
            float *sq = s_q + (size_t)warp * kFkQueriesPerWarp * c;
            __syncwarp();
            // queries that ARE points of the staged chunk (the EdgeConv case) come from the transposed tile, not from global memory
            const bool self_q = qry == pts && ldq == ldp && t0 == 0 && q0 + kFkQueriesPerWarp <= tn;
            if (self_q) {
#pragma unroll
                for (int w = 0; w < kQ; ++w)
                    for (int l = lane; l < c; l += 32) sq[w * c + l] = s_ft[(size_t)l * kStride + q0 + u0 + w];
            } else {
                for (int e = lane; e < kQ * c; e += 32) {
                    const int w = e / c, l = e - w * c;
                    sq[e] = __ldg(qry + (size_t)min(q0 + u0 + w, m - 1) * ldq + l);
                }
            }
            __syncwarp();
            float rq[kQ];
            if (self_q) {  // the point norms were accumulated in exactly this order
#pragma unroll
                for (int w = 0; w < kQ; ++w) rq[w] = s_r[q0 + u0 + w];
            } else {
#pragma unroll
                for (int w = 0; w < kQ; ++w) rq[w] = 0.f;
                for (int l = 0; l < c; ++l) {
#pragma unroll
                    for (int w = 0; w < kQ; ++w) rq[w] = __fmaf_rn(sq[w * c + l], sq[w * c + l], rq[w]);
                }
            }
            FK_MARK();
            float d[kQ][P];
#pragma unroll
            for (int w = 0; w < kQ; ++w)
#pragma unroll
                for (int t = 0; t < P; ++t) d[w][t] = 0.f;
            for (int l = 0; l < c; ++l) {
                float a[kQ];
#pragma unroll
                for (int w = 0; w < kQ; ++w) a[w] = sq[w * c + l];
                const float *col = s_ft + (size_t)l * kStride + lane;
#pragma unroll
                for (int t = 0; t < P; ++t) {
                    const float v = col[t * 32];
#pragma unroll
                    for (int w = 0; w < kQ; ++w) d[w][t] = __fmaf_rn(a[w], v, d[w][t]);
                }
            }
#pragma unroll
            for (int t = 0; t < P; ++t) {
                const int p = t * 32 + lane;
                const float rp = s_r[p];
#pragma unroll
                for (int w = 0; w < kQ; ++w) d[w][t] = p < tn ? __fadd_rn(__fmaf_rn(-2.f, d[w][t], rq[w]), rp) : inf;
            }
            FK_MARK();
            float cd[kQ], nd[kQ];
            int ci[kQ], ni[kQ];
#pragma unroll
            for (int w = 0; w < kQ; ++w) { cd[w] = bd[u0 + w]; ci[w] = bi[u0 + w]; nd[w] = inf; ni[w] = 0x7fffffff; }
            // ---- selection, fast path: threshold by LANE MINIMA.  The k-th smallest of the 32 per-lane minima bounds the
            // k-th smallest value from above (k different lanes hold something <= it); everything <= that bound (about
            // k + a few elements) is compacted into shared memory by ballot and ordered by rank counting on (d, idx) —
            // ties to the lower index.  A query whose survivors overflow the buffer (masses of equal distances) takes the
            // round-by-round extraction below instead.
            unsigned slow = 0;  // bit w: query w needs the fallback
            // survivors as ONE 64-bit key (order-preserving distance key : index): a single unsigned comparison is the
            // (d, idx) lexicographic order — ties to the lower index — and a survivor is one 8-byte store
            unsigned long long *sk = reinterpret_cast<unsigned long long *>(s_sd) + warp * kFkCap;  // the s_sd | s_si region: 2 x warps x kFkCap words
            const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
            for (int w = 0; w < kQ; ++w) {
                float lmin = cd[w];
#pragma unroll
                for (int t = 0; t < P; ++t) lmin = fminf(lmin, d[w][t]);
                const float bound = __shfl_sync(kFull, fk_warp_sort32(lmin, lane), k - 1);
                int S = 0;
                {
                    const bool keep = cd[w] <= bound && cd[w] < inf;
                    const unsigned mask = __ballot_sync(kFull, keep);
                    if (keep) { const int pos = __popc(mask & lt_mask); sk[pos] = ((unsigned long long)ord_key(cd[w]) << 32) | (unsigned)ci[w]; }
                    S = __popc(mask);
                }
#pragma unroll
                for (int t = 0; t < P; ++t) {
                    const bool keep = d[w][t] <= bound && d[w][t] < inf;
                    const unsigned mask = __ballot_sync(kFull, keep);
                    if (mask) {
                        const int pos = S + __popc(mask & lt_mask);
                        if (keep && pos < kFkCap) sk[pos] = ((unsigned long long)ord_key(d[w][t]) << 32) | (unsigned)(t0 + t * 32 + lane);
                        S += __popc(mask);
                    }
                }
                __syncwarp();
                if (S > kFkCap || S < k) {  // (S < k only with NaN/inf distances)
                    slow |= 1u << w;
                } else {
                    for (int e0 = 0; e0 < S; e0 += 32) {
                        const int e = e0 + lane;
                        const unsigned long long ke = e < S ? sk[e] : ~0ull;
                        int rank = 0;
#pragma unroll 4
                        for (int f = 0; f < S; ++f) rank += sk[f] < ke;
                        if (e < S && rank < 32) { s_od[warp * 32 + rank] = ord_key_inv((unsigned)(ke >> 32)); s_oi[warp * 32 + rank] = (int)(unsigned)ke; }
                    }
                    __syncwarp();
                    nd[w] = lane < k ? s_od[warp * 32 + lane] : inf;
                    ni[w] = lane < k ? s_oi[warp * 32 + lane] : 0x7fffffff;
                }
                __syncwarp();
            }
            // ---- fallback: k rounds of warp arg-min
            for (int r = 0; r < (slow ? k : 0); ++r) {
#pragma unroll
                for (int w = 0; w < kQ; ++w) {
                    if (!(slow & (1u << w))) continue;
                    // lane-local minimum; strict '<' keeps the carried entry (lower index) and the lower slot on ties
                    float lm = cd[w];
                    int lt = -1;
#pragma unroll
                    for (int t = 0; t < P; ++t)
                        if (d[w][t] < lm) { lm = d[w][t]; lt = t; }
                    const int gi = lt < 0 ? ci[w] : t0 + lt * 32 + lane;
                    const unsigned key = ord_key(lm);
                    const unsigned kmin = __reduce_min_sync(kFull, key);
                    const unsigned imin = __reduce_min_sync(kFull, key == kmin ? (unsigned)gi : 0xffffffffu);
                    const bool win = key == kmin && (unsigned)gi == imin;
                    const int wl = __ffs(__ballot_sync(kFull, win)) - 1;
                    const float dmin = __shfl_sync(kFull, lm, wl);
                    if (lane == r) { nd[w] = dmin; ni[w] = (int)imin; }
                    if (win && lane == wl) {
                        if (lt < 0) {
                            cd[w] = inf;
                        } else {
#pragma unroll
                            for (int t = 0; t < P; ++t)
                                if (t == lt) d[w][t] = inf;
                        }
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < kQ; ++w) { bd[u0 + w] = nd[w]; bi[u0 + w] = ni[w]; }
            FK_MARK();
        }
    }
#pragma unroll
    for (int u = 0; u < kFkQueriesPerWarp; ++u) {
        const int j = q0 + u;
        if (j < m && lane >= skip && lane < k) {
            const size_t o = ((size_t)cloud * m + j) * (k - skip) + (lane - skip);
            if (dist) dist[o] = bd[u];
            idx[o] = bi[u];
        }
    }
    FK_MARK();
    if (pj.proj == nullptr) return;
    // ---- EdgeConv projection of this CTA's 32 query points (queries == points here): saves the separate projection
    // launch of dispu_dense_conv.  Same derived weights and the same accumulation order as edge_proj_kernel.
    __syncthreads();  // every warp is done with the staged features: reuse the region for [c][96] weights + 96 biases
    float *s_w = fk_smem, *s_b = fk_smem + (size_t)c * kEcProjW;
    const int tid = threadIdx.x;
    if (pj.raw_async) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float *r0 = s_raw, *r1 = s_raw + 48 * c, *r2 = s_raw + 72 * c;
        if (tid < 2 * kEcProjW) {  // thread = (output column o, parity of the input row): no index arithmetic in the loop
            const int o = tid % kEcProjW, l0 = tid / kEcProjW;
            const int part = o / kEcG, oc = o - part * kEcG;
            const float *src = part == 0 ? r0 + c * kEcG : part == 1 ? r0 : part == 2 ? r1 : r2;
            if (part == 1) {
#pragma unroll 4
                for (int l = l0; l < c; l += 2) s_w[l * kEcProjW + o] = __fsub_rn(src[l * kEcG + oc], src[(c + l) * kEcG + oc]);
            } else {
#pragma unroll 8
                for (int l = l0; l < c; l += 2) s_w[l * kEcProjW + o] = src[l * kEcG + oc];
            }
            if (l0 == 0) {
                const float *bsrc = part == 1 ? pj.b0 : part == 2 ? pj.b1 : part == 3 ? pj.b2 : nullptr;
                s_b[o] = bsrc ? __ldg(bsrc + oc) : 0.f;
            }
        }
    } else if (tid < 2 * kEcProjW) {
        const int o = tid % kEcProjW, l0 = tid / kEcProjW;
        const int part = o / kEcG, oc = o - part * kEcG;
        const float *src = part == 0 ? pj.w0 + (size_t)c * kEcG : part == 1 ? pj.w0 : part == 2 ? pj.w1 + (size_t)kEcG * kEcG
                                                                                               : pj.w2 + (size_t)2 * kEcG * kEcG;
        if (part == 1) {
#pragma unroll 4
            for (int l = l0; l < c; l += 2)
                s_w[l * kEcProjW + o] = __fsub_rn(__ldg(src + (size_t)l * kEcG + oc), __ldg(src + (size_t)(c + l) * kEcG + oc));
        } else {
#pragma unroll 8
            for (int l = l0; l < c; l += 2) s_w[l * kEcProjW + o] = __ldg(src + (size_t)l * kEcG + oc);
        }
        if (l0 == 0) {
            const float *bsrc = part == 1 ? pj.b0 : part == 2 ? pj.b1 : part == 3 ? pj.b2 : nullptr;
            s_b[o] = bsrc ? __ldg(bsrc + oc) : 0.f;
        }
    }
    __syncthreads();
    FK_MARK();
    {   // thread -> (row = tid / 8 of the CTA's 32 queries, 12 of the 96 columns)
        const int rq = tid >> 3, cg = tid & 7;
        const int j = blockIdx.x * (kFkWarps * kFkQueriesPerWarp) + rq;
        if (j < m) {
            const float *f = qry + (size_t)j * ldq;
            float acc[12];
#pragma unroll
            for (int o = 0; o < 12; ++o) acc[o] = s_b[cg * 12 + o];
            const float4 *w4 = reinterpret_cast<const float4 *>(s_w) + cg * 3;
            const bool fvec = ((c | ldq) & 3) == 0 && (reinterpret_cast<uintptr_t>(qry) & 15u) == 0;
#pragma unroll 4
            for (int l0 = 0; l0 < c; l0 += 4) {  // (unrolled: four 16-byte loads of the row in flight)
                float fv4[4];
                if (fvec) {
                    const float4 t4 = __ldg(reinterpret_cast<const float4 *>(f + l0));
                    fv4[0] = t4.x; fv4[1] = t4.y; fv4[2] = t4.z; fv4[3] = t4.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) fv4[i] = l0 + i < c ? __ldg(f + l0 + i) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int l = l0 + i;
                    if (l < c) {
                        const float fv = fv4[i];
#pragma unroll
                        for (int q4 = 0; q4 < 3; ++q4) {
                            const float4 w = w4[(size_t)l * (kEcProjW / 4) + q4];
                            acc[4 * q4 + 0] = __fmaf_rn(fv, w.x, acc[4 * q4 + 0]);
                            acc[4 * q4 + 1] = __fmaf_rn(fv, w.y, acc[4 * q4 + 1]);
                            acc[4 * q4 + 2] = __fmaf_rn(fv, w.z, acc[4 * q4 + 2]);
                            acc[4 * q4 + 3] = __fmaf_rn(fv, w.w, acc[4 * q4 + 3]);
                        }
                    }
                }
            }
            float4 *o4 = reinterpret_cast<float4 *>(pj.proj + ((size_t)cloud * m + j) * kEcProjW + cg * 12);
#pragma unroll
            for (int q4 = 0; q4 < 3; ++q4) o4[q4] = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
        }
    }
    FK_MARK();
#undef FK_MARK
}

// ------------------------------------------------------------------------------------------------------------------
// EdgeConv dense block
constexpr int kEcThreads = 256;
constexpr int kEcProjRows = 64;

// proj[row] = [f.W0e | f.(W0c-W0e)+b0 | f.W1c+b1 | f.W2c+b2]
__global__ void __launch_bounds__(kEcThreads)
edge_proj_kernel(long long rows, int c, const float *__restrict__ feat, int ldf, const float *__restrict__ w0,
                 const float *__restrict__ b0, const float *__restrict__ w1, const float *__restrict__ b1,
                 const float *__restrict__ w2, const float *__restrict__ b2, float *__restrict__ proj) {
    extern __shared__ __align__(16) float ep_smem[];
    float *s_w = ep_smem;                          // [c][96]
    float *s_b = s_w + (size_t)c * kEcProjW;       // [96]
    float *s_f = s_b + kEcProjW;                   // [64][c+1]
    const int tid = threadIdx.x;
    const int fs = c + 1;
    if (tid < 2 * kEcProjW) {  // thread -> fixed output column o (so a fixed source matrix), rows l0, l0+2, ...: plain strided loads
        const int o = tid % kEcProjW, l0 = tid / kEcProjW;
        const int part = o / kEcG, oc = o - part * kEcG;
        const float *src = part == 0 ? w0 + (size_t)c * kEcG : part == 1 ? w0 : part == 2 ? w1 + (size_t)kEcG * kEcG
                                                                                         : w2 + (size_t)2 * kEcG * kEcG;
        if (part == 1) {
#pragma unroll 4
            for (int l = l0; l < c; l += 2)
                s_w[l * kEcProjW + o] = __fsub_rn(__ldg(src + (size_t)l * kEcG + oc), __ldg(src + (size_t)(c + l) * kEcG + oc));
        } else {
#pragma unroll 8
            for (int l = l0; l < c; l += 2) s_w[l * kEcProjW + o] = __ldg(src + (size_t)l * kEcG + oc);
        }
    }
    for (int o = tid; o < kEcProjW; o += kEcThreads) {
        const int part = o / kEcG, oc = o - part * kEcG;
        float v = 0.f;
        if (part == 1 && b0) v = __ldg(b0 + oc);
        if (part == 2 && b1) v = __ldg(b1 + oc);
        if (part == 3 && b2) v = __ldg(b2 + oc);
        s_b[o] = v;
    }
    const long long row0 = (long long)blockIdx.x * kEcProjRows;
    for (int e = tid; e < kEcProjRows * c; e += kEcThreads) {
        const int r = e / c, l = e - r * c;
        const long long row = row0 + r;
        s_f[r * fs + l] = row < rows ? __ldg(feat + row * (long long)ldf + l) : 0.f;
    }
    __syncthreads();
    const int r = tid >> 2, part = tid & 3;
    float acc[kEcG];
#pragma unroll
    for (int o = 0; o < kEcG; ++o) acc[o] = s_b[part * kEcG + o];
    const float4 *w4 = reinterpret_cast<const float4 *>(s_w) + part * (kEcG / 4);
    for (int l = 0; l < c; ++l) {
        const float f = s_f[r * fs + l];
#pragma unroll
        for (int q = 0; q < kEcG / 4; ++q) {
            const float4 w = w4[(size_t)l * (kEcProjW / 4) + q];
            acc[4 * q + 0] = __fmaf_rn(f, w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = __fmaf_rn(f, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = __fmaf_rn(f, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = __fmaf_rn(f, w.w, acc[4 * q + 3]);
        }
    }
    const long long row = row0 + r;
    if (row < rows) {
        float4 *o4 = reinterpret_cast<float4 *>(proj + row * kEcProjW + part * kEcG);
#pragma unroll
        for (int q = 0; q < kEcG / 4; ++q) o4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
}

// Packed FP32 pairs (FFMA2: one issue slot, two IEEE fma.rn results — bit-identical to two __fmaf_rn).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// y += x . W for two edges at once (x: 24 values each, W: 24x24 row-major in shared memory read as 16-byte broadcasts;
// y kept as 12 packed column pairs per edge)
__device__ __forceinline__ void ec_matvec2(const ulonglong2 *__restrict__ w2, const float (&xa)[kEcG], const float (&xb)[kEcG],
                                           f32x2 (&ya)[kEcG / 2], f32x2 (&yb)[kEcG / 2]) {
#pragma unroll
    for (int i = 0; i < kEcG; ++i) {
        const f32x2 xa2 = pack2(xa[i], xa[i]), xb2 = pack2(xb[i], xb[i]);
#pragma unroll
        for (int q = 0; q < kEcG / 4; ++q) {
            const ulonglong2 w = w2[i * (kEcG / 4) + q];  // columns 4q..4q+3 of row i
            ya[2 * q] = fma2(xa2, w.x, ya[2 * q]);
            ya[2 * q + 1] = fma2(xa2, w.y, ya[2 * q + 1]);
            yb[2 * q] = fma2(xb2, w.x, yb[2 * q]);
            yb[2 * q + 1] = fma2(xb2, w.y, yb[2 * q + 1]);
        }
    }
}

__device__ __forceinline__ void ec_load24(const float *__restrict__ p, float (&v)[kEcG]) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int q = 0; q < kEcG / 4; ++q) {
        const float4 t = __ldg(p4 + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
__device__ __forceinline__ void ec_load24x2(const float *__restrict__ p, f32x2 (&a)[kEcG / 2], f32x2 (&b)[kEcG / 2]) {
    const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(p);
#pragma unroll
    for (int q = 0; q < kEcG / 4; ++q) {
        const ulonglong2 t = __ldg(p2 + q);
        a[2 * q] = t.x; a[2 * q + 1] = t.y;
        b[2 * q] = t.x; b[2 * q + 1] = t.y;
    }
}

// 8 lanes per point, two edges per lane per pass (k <= 16: a single pass, kSingle, with no running maximum held in
// registers).  The per-edge 24x24 products run on packed FFMA2.  out row = [max y2 | max y1 | max y0 | f_i].
constexpr int kEcLanes = 8;
template <bool kSingle>
__global__ void __launch_bounds__(kEcThreads, 1)
edge_mlp_max_kernel(long long rows, int n, int c, int k, const float *__restrict__ proj, const int *__restrict__ idx,
                    const float *__restrict__ w1, const float *__restrict__ w2, const float *__restrict__ feat, int ldf,
                    float *__restrict__ out, int ldo) {
    __shared__ __align__(16) float s_w1a[kEcG * kEcG], s_w2a[kEcG * kEcG], s_w2b[kEcG * kEcG];
    const int tid = threadIdx.x;
    for (int e = tid; e < kEcG * kEcG; e += kEcThreads) {
        s_w1a[e] = __ldg(w1 + e);                  // W1[:g]
        s_w2a[e] = __ldg(w2 + e);                  // W2[:g]     (multiplies y1)
        s_w2b[e] = __ldg(w2 + kEcG * kEcG + e);    // W2[g:2g]   (multiplies y0)
    }
    __syncthreads();
    const int t = tid & (kEcLanes - 1);
    const long long row = (long long)blockIdx.x * (kEcThreads / kEcLanes) + (tid / kEcLanes);
    const bool valid = row < rows;
    const long long rowc = valid ? row : rows - 1;  // keep every lane in the shuffles
    const long long cloud = rowc / n;
    const float *prow = proj + rowc * kEcProjW;
    const int *irow = idx + rowc * (long long)k;
    const float neg_inf = __int_as_float(0xff800000);
    float m0[kEcG], m1[kEcG], m2[kEcG];
    if (!kSingle) {
#pragma unroll
        for (int o = 0; o < kEcG; ++o) { m0[o] = neg_inf; m1[o] = neg_inf; m2[o] = neg_inf; }
    }
    for (int s0 = t; s0 < (kSingle ? t + 1 : k); s0 += 2 * kEcLanes) {
        // second edge of the pair; past the end it repeats a valid edge (the maximum is unchanged)
        const int sa = kSingle ? min(s0, k - 1) : s0;
        const int sb = min(s0 + kEcLanes, k - 1);
        const int ja = __ldg(irow + sa), jb = __ldg(irow + sb);
        float ya0[kEcG], yb0[kEcG];
        {
            float pc[kEcG];
            ec_load24(prow + kEcG, pc);
            ec_load24(proj + (cloud * n + ja) * kEcProjW, ya0);
            ec_load24(proj + (cloud * n + jb) * kEcProjW, yb0);
#pragma unroll
            for (int o = 0; o < kEcG; ++o) {
                ya0[o] = fmaxf(__fadd_rn(ya0[o], pc[o]), 0.f);
                yb0[o] = fmaxf(__fadd_rn(yb0[o], pc[o]), 0.f);
            }
        }
        f32x2 ya2[kEcG / 2], yb2[kEcG / 2], ya1[kEcG / 2], yb1[kEcG / 2];
        ec_load24x2(prow + 3 * kEcG, ya2, yb2);
        ec_matvec2(reinterpret_cast<const ulonglong2 *>(s_w2b), ya0, yb0, ya2, yb2);
        ec_load24x2(prow + 2 * kEcG, ya1, yb1);
        ec_matvec2(reinterpret_cast<const ulonglong2 *>(s_w1a), ya0, yb0, ya1, yb1);
        float xa1[kEcG], xb1[kEcG];
#pragma unroll
        for (int o = 0; o < kEcG / 2; ++o) {
            unpack2(ya1[o], xa1[2 * o], xa1[2 * o + 1]);
            unpack2(yb1[o], xb1[2 * o], xb1[2 * o + 1]);
        }
#pragma unroll
        for (int o = 0; o < kEcG; ++o) {
            xa1[o] = fmaxf(xa1[o], 0.f);
            xb1[o] = fmaxf(xb1[o], 0.f);
            const float v0 = fmaxf(ya0[o], yb0[o]), v1 = fmaxf(xa1[o], xb1[o]);
            m0[o] = kSingle ? v0 : fmaxf(m0[o], v0);
            m1[o] = kSingle ? v1 : fmaxf(m1[o], v1);
        }
        ec_matvec2(reinterpret_cast<const ulonglong2 *>(s_w2a), xa1, xb1, ya2, yb2);
#pragma unroll
        for (int o = 0; o < kEcG / 2; ++o) {
            float a0, a1, b0, b1;
            unpack2(ya2[o], a0, a1);
            unpack2(yb2[o], b0, b1);
            const float v0 = fmaxf(a0, b0), v1 = fmaxf(a1, b1);
            m2[2 * o] = kSingle ? v0 : fmaxf(m2[2 * o], v0);
            m2[2 * o + 1] = kSingle ? v1 : fmaxf(m2[2 * o + 1], v1);
        }
    }
    // max over the 8 lanes of the point
#pragma unroll
    for (int off = 1; off < kEcLanes; off <<= 1) {
#pragma unroll
        for (int o = 0; o < kEcG; ++o) {
            m0[o] = fmaxf(m0[o], __shfl_xor_sync(kFull, m0[o], off));
            m1[o] = fmaxf(m1[o], __shfl_xor_sync(kFull, m1[o], off));
            m2[o] = fmaxf(m2[o], __shfl_xor_sync(kFull, m2[o], off));
        }
    }
    if (!valid) return;
    float4 *o4 = reinterpret_cast<float4 *>(out + row * (long long)ldo);
    // 18 float4 of maxima: chunk q goes to lane q % 8 (static register indices, predicated stores)
#pragma unroll
    for (int q = 0; q < kEcG / 4; ++q) {
        if ((q & 7) == t) o4[q] = make_float4(m2[4 * q], m2[4 * q + 1], m2[4 * q + 2], m2[4 * q + 3]);
        if (((q + 6) & 7) == t) o4[q + 6] = make_float4(m1[4 * q], m1[4 * q + 1], m1[4 * q + 2], m1[4 * q + 3]);
        if (((q + 12) & 7) == t) o4[q + 12] = make_float4(m0[4 * q], m0[4 * q + 1], m0[4 * q + 2], m0[4 * q + 3]);
    }
    const float4 *f4 = reinterpret_cast<const float4 *>(feat + row * (long long)ldf);
    for (int q = t; q < c / 4; q += kEcLanes) o4[3 * kEcG / 4 + q] = __ldg(f4 + q);
}

// ------------------------------------------------------------------------------------------------------------------
// weight net + transposed product:  H[i][ch][t] = sum_s G[i][s][ch] * relu(((xyz[idx[i][s]]-xyz[i]).Ww[:,t] + bw[t]) * scale)
constexpr int kWaThreads = 256;
constexpr int kWaPts = 4;   // points per CTA
constexpr int kWaNs = 16;   // nsample
constexpr int kWaT = 16;    // hidden units of the weight net
constexpr int kWaC = 128;   // channels of G

__global__ void __launch_bounds__(kWaThreads)
weightnet_aggregate_kernel(long long pts_total, int m, const float *__restrict__ xyz, const int *__restrict__ idx,
                           const float *__restrict__ G, const float *__restrict__ Ww, const float *__restrict__ bw,
                           float bn_scale, float *__restrict__ out) {
    __shared__ __align__(16) float s_g[kWaPts][kWaNs][kWaC];   // 32 KB
    __shared__ __align__(16) float s_wn[kWaPts][kWaNs][kWaT];  // 4 KB
    const int tid = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * kWaPts;
    const int np = (int)min((long long)kWaPts, pts_total - p0);
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(G + p0 * (long long)(kWaNs * kWaC));
        float4 *s4 = reinterpret_cast<float4 *>(&s_g[0][0][0]);
        const int n4 = np * kWaNs * kWaC / 4;
        for (int e = tid; e < n4; e += kWaThreads) s4[e] = __ldg(g4 + e);
    }
    for (int e = tid; e < np * kWaNs * kWaT; e += kWaThreads) {
        const int p = e / (kWaNs * kWaT), rem = e - p * (kWaNs * kWaT);
        const int s = rem / kWaT, t = rem - s * kWaT;
        const long long pt = p0 + p;
        const long long cloud = pt / m;
        const float *ctr = xyz + pt * 3;
        const float *q = xyz + (cloud * m + __ldg(idx + pt * kWaNs + s)) * 3;
        const float gx = __fsub_rn(__ldg(q), __ldg(ctr)), gy = __fsub_rn(__ldg(q + 1), __ldg(ctr + 1)),
                    gz = __fsub_rn(__ldg(q + 2), __ldg(ctr + 2));
        float a = __fmaf_rn(gz, __ldg(Ww + 2 * kWaT + t), __fmaf_rn(gy, __ldg(Ww + kWaT + t), __fmul_rn(gx, __ldg(Ww + t))));
        a = __fmul_rn(__fadd_rn(a, bw ? __ldg(bw + t) : 0.f), bn_scale);
        s_wn[p][s][t] = fmaxf(a, 0.f);
    }
    __syncthreads();
    // thread -> (point pair slot, 4 channels, 4 hidden units): 16 accumulators, 16 FMA per two LDS.128
    const int half = tid >> 7;         // which of two points per pass
    const int chq = (tid & 127) >> 2;  // channels 4*chq .. 4*chq+3
    const int tq = tid & 3;            // hidden units 4*tq .. 4*tq+3
    for (int pp = half; pp < np; pp += 2) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = 0.f;
#pragma unroll
        for (int s = 0; s < kWaNs; ++s) {
            const float4 g = *reinterpret_cast<const float4 *>(&s_g[pp][s][4 * chq]);
            const float4 w = *reinterpret_cast<const float4 *>(&s_wn[pp][s][4 * tq]);
            const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                acc[a][0] = __fmaf_rn(gv[a], w.x, acc[a][0]);
                acc[a][1] = __fmaf_rn(gv[a], w.y, acc[a][1]);
                acc[a][2] = __fmaf_rn(gv[a], w.z, acc[a][2]);
                acc[a][3] = __fmaf_rn(gv[a], w.w, acc[a][3]);
            }
        }
        float *o = out + (p0 + pp) * (long long)(kWaC * kWaT);
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<float4 *>(o + (4 * chq + a) * kWaT + 4 * tq) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
}

static inline bool ec_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace dispu

using namespace dispu;

template <int P>
static int launch_feature_knn(int b, int n, int m, int c, int k, int skip, const float *points, int ldp, const float *queries,
                              int ldq, float *dist, int *idx, cudaStream_t stream, const EcProjArgs &pj = EcProjArgs{}) {
    size_t smem = ((size_t)c * (32 * P + 1) + 32 * P + (size_t)kFkWarps * kFkQueriesPerWarp * c + 2 * kFkWarps * (kFkCap + 32) + 2) * sizeof(float);
    const size_t smem_proj = ((size_t)c * kEcProjW + kEcProjW) * sizeof(float);
    if (pj.proj && smem < smem_proj) smem = smem_proj;
    EcProjArgs pja = pj;
    if (pj.proj && pj.raw_async) {  // raw weight rows fetched with cp.async — when they fit next to the staged cloud
        if (smem + (size_t)96 * c * sizeof(float) <= 220 * 1024) smem += (size_t)96 * c * sizeof(float);
        else pja.raw_async = 0;
    }
    if (smem > 220 * 1024) return DISPU_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(feature_knn_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    dim3 grid(ceil_div(m, kFkWarps * kFkQueriesPerWarp), b);
    feature_knn_kernel<P><<<grid, kFkWarps * 32, smem, stream>>>(n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, pja);
    return finish_launch();
}

static int feature_knn_dispatch(int b, int n, int m, int c, int k, int skip, const float *points, int ldp, const float *queries,
                                int ldq, float *dist, int *idx, cudaStream_t stream, const EcProjArgs &pj = EcProjArgs{}) {
    // largest chunk (32*P candidates) that covers the cloud and whose transposed features fit in shared memory
    auto fits = [&](int P) {
        return ((size_t)c * (32 * P + 1) + 32 * P + (size_t)kFkWarps * kFkQueriesPerWarp * c + 2 * kFkWarps * (kFkCap + 32) + 2) * sizeof(float) <= 220 * 1024;
    };
    static const int force_p = [] { const char *e = getenv("DISPU_FK_P"); return e ? atoi(e) : 0; }();  // development switch (A/B timing)
    if (force_p == 8) return launch_feature_knn<8>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
    if (force_p == 16 && fits(16)) return launch_feature_knn<16>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
    // Clouds of more than 512 points: measured at b = 32, n = 1024 (c4's second stage), one 1024-candidate chunk (one CTA of
    // 251 registers per SM) loses to several smaller chunks with the carried list (EdgeConv block, c = 48: 398 us vs 351 us
    // with 256-candidate chunks; c = 24: 281 vs 238 us with 512-candidate chunks)
    if (n > 512 && force_p == 0) {
        if (c <= 32 && fits(16)) return launch_feature_knn<16>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
        return launch_feature_knn<8>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
    }
    if (n > 512 && fits(32)) return launch_feature_knn<32>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
    if (n > 256 && fits(16)) return launch_feature_knn<16>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
    return launch_feature_knn<8>(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream, pj);
}

extern "C" int dispu_feature_knn(int b, int n, int m, int c, int k, int skip, const float *points, int ldp,
                                 const float *queries, int ldq, float *dist, int *idx, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0 && c > 0 && k > 0 && skip >= 0 && skip < k, DISPU_EINVAL);
    DISPU_REQUIRE(k <= n && ldp >= c && ldq >= c, DISPU_EINVAL);
    DISPU_REQUIRE(k <= 32 && c <= 128, DISPU_EUNSUPPORTED);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(points && queries && idx, DISPU_ENULL);
    return feature_knn_dispatch(b, n, m, c, k, skip, points, ldp, queries, ldq, dist, idx, stream);
}

extern "C" size_t dispu_dense_conv_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    return (size_t)b * n * kEcProjW * sizeof(float);
}

static long long *g_fk_dbg = nullptr;
extern "C" int dispu__fk_debug_trace(long long *host_out) {  // development only
    if (!g_fk_dbg || !host_out) return DISPU_EUNSUPPORTED;
    return (int)cudaMemcpy(host_out, g_fk_dbg, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
}

namespace dispu {
// tensor-core form of the per-edge MLP (edgeconv_tc.cu); returns false when the shape is not one it is built for
bool edge_mlp_tc_launch(long long rows, int n, int c, int k, int growth, const float *proj, const int *idx, const float *w1,
                        const float *w2, const float *feat, int ldf, float *out, int ldo, cudaStream_t stream, int *rc);
}  // namespace dispu

static int dense_conv_impl(int b, int n, int c, int k, int growth, const float *feat, int ldf, const int *idx, int *idx_out,
                           const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                           float *out, int ldo, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && c > 0 && k > 0 && growth > 0, DISPU_EINVAL);
    DISPU_REQUIRE(ldf >= c && ldo >= 3 * growth + c, DISPU_EINVAL);
    DISPU_REQUIRE(growth == kEcG && c % 4 == 0 && c <= 128 && ldf % 4 == 0 && ldo % 4 == 0, DISPU_EUNSUPPORTED);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(feat && (idx || idx_out) && w0 && w1 && w2 && out, DISPU_ENULL);
    DISPU_REQUIRE(ec_aligned16(feat) && ec_aligned16(out), DISPU_EUNSUPPORTED);
    DISPU_REQUIRE(workspace && workspace_bytes >= dispu_dense_conv_workspace_bytes(b, n) && ec_aligned16(workspace),
                  DISPU_EWORKSPACE);
    float *proj = reinterpret_cast<float *>(workspace);
    const long long rows = (long long)b * n;
    if (idx_out) {
        // neighbour search and projection in ONE kernel (get_edge_feature: k+1 nearest in feature space, self dropped)
        DISPU_REQUIRE(k + 1 <= n && k + 1 <= 32, DISPU_EUNSUPPORTED);
        static long long *dbg = nullptr;  // development only (DISPU_FK_TRACE)
        if (getenv("DISPU_FK_TRACE") && !dbg) { cudaMalloc(&dbg, 64 * sizeof(long long)); g_fk_dbg = dbg; }
        if (dbg) cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), stream);
        const int raw_async = ec_aligned16(w0) && ec_aligned16(w1) && ec_aligned16(w2);
        EcProjArgs pj{w0, b0, w1, b1, w2, b2, proj, raw_async, dbg};
        const int rc = feature_knn_dispatch(b, n, n, c, k + 1, 1, feat, ldf, feat, ldf, nullptr, idx_out, stream, pj);
        if (rc != DISPU_OK) return rc;
        idx = idx_out;
    } else {
        const size_t smem = ((size_t)c * kEcProjW + kEcProjW + (size_t)kEcProjRows * (c + 1)) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(edge_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        count_launch();
        edge_proj_kernel<<<(unsigned)ceil_div64(rows, kEcProjRows), kEcThreads, smem, stream>>>(rows, c, feat, ldf, w0, b0, w1, b1,
                                                                                              w2, b2, proj);
    }
    {   // k = 16: the per-edge products run on tcgen05 (activations in tensor memory); other k: packed-FFMA2 kernel below
        int rc = DISPU_OK;
        if (edge_mlp_tc_launch(rows, n, c, k, growth, proj, idx, w1, w2, feat, ldf, out, ldo, stream, &rc)) return rc;
    }
    count_launch();
    const unsigned g2 = (unsigned)ceil_div64(rows, kEcThreads / kEcLanes);
    if (k <= 2 * kEcLanes)
        edge_mlp_max_kernel<true><<<g2, kEcThreads, 0, stream>>>(rows, n, c, k, proj, idx, w1, w2, feat, ldf, out, ldo);
    else
        edge_mlp_max_kernel<false><<<g2, kEcThreads, 0, stream>>>(rows, n, c, k, proj, idx, w1, w2, feat, ldf, out, ldo);
    return finish_launch();
}

extern "C" int dispu_dense_conv(int b, int n, int c, int k, int growth, const float *feat, int ldf, const int *idx,
                                const float *w0, const float *b0, const float *w1, const float *b1, const float *w2,
                                const float *b2, float *out, int ldo, void *workspace, size_t workspace_bytes,
                                dispu_stream_t stream_) {
    DISPU_REQUIRE(idx, DISPU_ENULL);
    return dense_conv_impl(b, n, c, k, growth, feat, ldf, idx, nullptr, w0, b0, w1, b1, w2, b2, out, ldo, workspace, workspace_bytes,
                           (cudaStream_t)stream_);
}

extern "C" int dispu_dense_conv_knn(int b, int n, int c, int k, int growth, const float *feat, int ldf, int *idx_out,
                                    const float *w0, const float *b0, const float *w1, const float *b1, const float *w2,
                                    const float *b2, float *out, int ldo, void *workspace, size_t workspace_bytes,
                                    dispu_stream_t stream_) {
    DISPU_REQUIRE(idx_out, DISPU_ENULL);
    return dense_conv_impl(b, n, c, k, growth, feat, ldf, nullptr, idx_out, w0, b0, w1, b1, w2, b2, out, ldo, workspace,
                           workspace_bytes, (cudaStream_t)stream_);
}

extern "C" int dispu_weightnet_aggregate(int b, int m, int nsample, int C, int T, const float *xyz, const int *idx,
                                         const float *G, const float *Ww, const float *bw, float bn_scale, float *out,
                                         dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && m > 0 && nsample > 0 && C > 0 && T > 0, DISPU_EINVAL);
    DISPU_REQUIRE(nsample == kWaNs && C == kWaC && T == kWaT, DISPU_EUNSUPPORTED);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz && idx && G && Ww && out, DISPU_ENULL);
    DISPU_REQUIRE(ec_aligned16(G) && ec_aligned16(out), DISPU_EUNSUPPORTED);
    const long long pts = (long long)b * m;
    count_launch();
    weightnet_aggregate_kernel<<<(unsigned)ceil_div64(pts, kWaPts), kWaThreads, 0, stream>>>(pts, m, xyz, idx, G, Ww, bw,
                                                                                           bn_scale, out);
    return finish_launch();
}

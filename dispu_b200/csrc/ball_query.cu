// QueryBallPoint for sm_100a — replaces query_ball_point_gpu (tf_ops/grouping/tf_grouping_g.cu:3-36).
//
// Reference: one CTA per cloud, one THREAD per query streaming the whole cloud from global
// memory, re-loading its own query every iteration, IEEE sqrtf per candidate, strided 4-byte
// index stores.  Here: the cloud is staged once per CTA into shared memory (SoA, coalesced
// global reads), one WARP owns a query and its 32 lanes test 32 consecutive candidates per
// step; `__ballot_sync` + `__popc` of the lower-lane mask gives every hit its slot in the same
// ascending-index order the serial scan produces, with the same early exit at nsample.
// The sqrt is removed exactly: sqrtf is monotone and correctly rounded, so
//     max(sqrtf(s),1e-20f) < r   <=>   r > 1e-20f  and  !(s > s_max(r))
// where s_max is the largest float whose rounded square root is still below r (found once per
// thread by stepping ulps around r*r).  NaN distances are hits in both forms.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tile_stage.cuh"

namespace dispu {

constexpr int kBqWarps = 8;
constexpr int kBqThreads = kBqWarps * 32;
constexpr int kBqMaxQpc = 64;   // queries per CTA (<= 8 per warp)
constexpr int kBqQi = 4;        // queries a warp scans together (independent ballot / count chains)
constexpr int kBqTile = 4096;   // points per shared-memory tile (48 KB)

__device__ __forceinline__ float ball_s_max(float r, bool &any) {
    any = r > 1e-20f;  // false for NaN / non-positive radii: fmaxf(.,1e-20f) < r never holds
    if (!any) return 0.f;
    float c = __fmul_rn(r, r);
    for (int it = 0; it < 4096 && __fsqrt_rn(c) < r; ++it) c = __uint_as_float(__float_as_uint(c) + 1u);
    for (int it = 0; it < 4096 && !(__fsqrt_rn(c) < r); ++it) c = __uint_as_float(__float_as_uint(c) - 1u);
    return c;
}

// kTma: tiles land as raw AoS bytes by cp.async.bulk + mbarrier (tile_stage.cuh), double-buffered when the cloud spans
// several tiles, and the scan reads them in place (lane k touches floats 3k..3k+2: stride 3 is conflict-free), so no
// thread issues a staging load or store at all.  Otherwise: per-thread __ldg staging into x | y | z arrays.
template <bool kTma>
__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(int n, int m, const float *__restrict__ radius, int ns, const float *__restrict__ xyz1,
                  const float *__restrict__ xyz2, int *__restrict__ idx, int *__restrict__ pts_cnt, int qpc,
                  int tile) {
    extern __shared__ __align__(16) float s_pts[];  // x[tile] | y[tile] | z[tile]   or (kTma)   raw tile 0 [| raw tile 1]
    __shared__ int s_cnt[kBqMaxQpc], s_first[kBqMaxQpc];
    __shared__ __align__(8) uint64_t s_bar[2];
    float *sx = s_pts, *sy = s_pts + tile, *sz = s_pts + 2 * tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.y;
    const int q0 = blockIdx.x * qpc;
    const int nq = min(qpc, m - q0);
    const float *pts = xyz1 + (size_t)cloud * n * 3;
    const float *qs = xyz2 + ((size_t)cloud * m + q0) * 3;
    int *rows = idx + ((size_t)cloud * m + q0) * ns;

    bool any;
    const float smax = ball_s_max(__ldg(radius), any);  // only radius[0] is read (tf_grouping_g.cu:25)
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int q = tid; q < nq; q += kBqThreads) {
        s_cnt[q] = 0;
        s_first[q] = 0;
    }
    const int ntiles = ceil_div(n, tile);
    if (kTma) {
        if (tid == 0) {
            ts::mbar_init(&s_bar[0], 1);
            ts::mbar_init(&s_bar[1], 1);
            ts::fence_mbar_init();
            ts::issue_tile(s_pts, pts, min(tile, n), &s_bar[0]);
        }
        __syncthreads();
    }
    int pending = -1;  // kTma: tile whose copy is in flight and has not been waited for

    for (int i = 0; i < ntiles; ++i) {
        const int t0 = i * tile;
        const int tn = min(tile, n - t0);
        const float *raw = s_pts;
        if (kTma) {
            // landing buffer (i+1)&1 was last scanned in iteration i-1, which every thread left through the
            // __syncthreads_and below: refill it while this tile is scanned
            if (i + 1 < ntiles) {
                if (tid == 0)
                    ts::issue_tile(s_pts + ((i + 1) & 1) * tile * 3, pts + (size_t)(t0 + tile) * 3, min(tile, n - t0 - tile), &s_bar[(i + 1) & 1]);
                pending = i + 1;
            } else {
                pending = -1;
            }
            ts::mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));
            raw = s_pts + (i & 1) * tile * 3;
        } else {
            for (int e = tid; e < tn * 3; e += kBqThreads) {
                const int k = e / 3, c = e - k * 3;
                s_pts[c * tile + k] = __ldg(pts + (size_t)t0 * 3 + e);
            }
            __syncthreads();
        }
        bool all_done = true;
        // kBqQi queries per warp at a time: one candidate load feeds kBqQi distance tests and the kBqQi ballot -> count ->
        // store chains are independent, so their latencies overlap (a single query's scan is one long dependent chain:
        // ballot, popc, the slot store and the exit test all wait for the previous step's count)
        for (int qb = warp * kBqQi; qb < nq; qb += kBqWarps * kBqQi) {
            int cnt[kBqQi], first[kBqQi];
            float qx[kBqQi], qy[kBqQi], qz[kBqQi];
            int *row[kBqQi];
#pragma unroll
            for (int u = 0; u < kBqQi; ++u) {
                const int q = min(qb + u, nq - 1);
                cnt[u] = qb + u < nq ? s_cnt[q] : ns;  // queries past the end count as finished
                first[u] = s_first[q];
                qx[u] = __ldg(qs + q * 3); qy[u] = __ldg(qs + q * 3 + 1); qz[u] = __ldg(qs + q * 3 + 2);
                row[u] = rows + (size_t)q * ns;
            }
            for (int k0 = 0; k0 < tn; k0 += 32) {
                bool open = false;
#pragma unroll
                for (int u = 0; u < kBqQi; ++u) open = open || cnt[u] < ns;
                if (!open) break;
                const int k = k0 + lane;
                float px = 0.f, py = 0.f, pz = 0.f;
                if (k < tn) {
                    px = kTma ? raw[3 * k] : sx[k]; py = kTma ? raw[3 * k + 1] : sy[k]; pz = kTma ? raw[3 * k + 2] : sz[k];
                }
                unsigned mask[kBqQi];
                bool hit[kBqQi];
#pragma unroll
                for (int u = 0; u < kBqQi; ++u) {
                    const float s = d2_ref_gpu(__fsub_rn(qx[u], px), __fsub_rn(qy[u], py), __fsub_rn(qz[u], pz));
                    hit[u] = k < tn && any && !(s > smax) && cnt[u] < ns;
                    mask[u] = __ballot_sync(kFull, hit[u]);
                }
#pragma unroll
                for (int u = 0; u < kBqQi; ++u) {
                    if (mask[u]) {
                        if (cnt[u] == 0) first[u] = t0 + k0 + __ffs(mask[u]) - 1;
                        const int pos = cnt[u] + __popc(mask[u] & lt_mask);
                        if (hit[u] && pos < ns) row[u][pos] = t0 + k;
                        cnt[u] += __popc(mask[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kBqQi; ++u) {
                if (qb + u < nq) {
                    if (lane == 0) {
                        s_cnt[qb + u] = cnt[u];
                        s_first[qb + u] = first[u];
                    }
                    all_done = all_done && (cnt[u] >= ns);
                }
            }
        }
        if (__syncthreads_and(all_done)) break;
        pending = -1;  // the next iteration waits for it
    }
    // never leave with a bulk copy still writing into this CTA's shared memory (early exit above)
    if (kTma && pending >= 0) ts::mbar_wait(&s_bar[pending & 1], (uint32_t)((pending >> 1) & 1));
    __syncthreads();
    for (int q = warp; q < nq; q += kBqWarps) {
        const int cnt = min(s_cnt[q], ns);
        const int fill = cnt > 0 ? s_first[q] : 0;  // reference leaves such rows uninitialised; we write zeros
        int *row = rows + (size_t)q * ns;
        for (int l = cnt + lane; l < ns; l += 32) row[l] = fill;
        if (lane == 0) pts_cnt[(size_t)cloud * m + q0 + q] = cnt;
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_query_ball_point(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                                      const float *xyz2, int *idx, int *pts_cnt, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(nsample > 0, DISPU_EINVAL);  // attr nsample>0, tf_grouping.cpp:75
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(radius && xyz1 && xyz2 && idx && pts_cnt, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    // queries per CTA: as many as keeps >= 2 CTAs per SM in flight, 32..64
    int qpw = (int)(((long long)b * m) / ((long long)kBqWarps * 2 * kNumSMs));
    qpw = qpw < kBqQi ? kBqQi : (qpw > 8 ? 8 : qpw);  // at least kBqQi: a warp scans kBqQi queries together
    const int qpc = qpw * kBqWarps;
    const int tile = n < kBqTile ? ((n + 31) & ~31) : kBqTile;
    static const bool no_tma = getenv("DISPU_NO_TMA") != nullptr;  // development switch (A/B timing)
    const bool tma = !no_tma && ts::tma_ok(xyz1, n);
    const size_t smem = sizeof(float) * 3 * (size_t)tile * (tma && n > tile ? 2 : 1);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ball_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(float) * 3 * kBqTile * 2));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(ball_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(float) * 3 * kBqTile));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    count_launch();
    if (tma)
        ball_query_kernel<true><<<dim3(ceil_div(m, qpc), b), kBqThreads, smem, stream>>>(n, m, radius, nsample, xyz1, xyz2,
                                                                                        idx, pts_cnt, qpc, tile);
    else
        ball_query_kernel<false><<<dim3(ceil_div(m, qpc), b), kBqThreads, smem, stream>>>(n, m, radius, nsample, xyz1, xyz2,
                                                                                         idx, pts_cnt, qpc, tile);
    return finish_launch();
}

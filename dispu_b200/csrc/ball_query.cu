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

#include "common.cuh"

namespace dispu {

constexpr int kBqWarps = 8;
constexpr int kBqThreads = kBqWarps * 32;
constexpr int kBqMaxQpc = 64;   // queries per CTA (<= 8 per warp)
constexpr int kBqTile = 4096;   // points per shared-memory tile (48 KB)

__device__ __forceinline__ float ball_s_max(float r, bool &any) {
    any = r > 1e-20f;  // false for NaN / non-positive radii: fmaxf(.,1e-20f) < r never holds
    if (!any) return 0.f;
    float c = __fmul_rn(r, r);
    for (int it = 0; it < 4096 && __fsqrt_rn(c) < r; ++it) c = __uint_as_float(__float_as_uint(c) + 1u);
    for (int it = 0; it < 4096 && !(__fsqrt_rn(c) < r); ++it) c = __uint_as_float(__float_as_uint(c) - 1u);
    return c;
}

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(int n, int m, const float *__restrict__ radius, int ns, const float *__restrict__ xyz1,
                  const float *__restrict__ xyz2, int *__restrict__ idx, int *__restrict__ pts_cnt, int qpc,
                  int tile) {
    extern __shared__ float s_pts[];  // x[tile] | y[tile] | z[tile]
    __shared__ int s_cnt[kBqMaxQpc], s_first[kBqMaxQpc];
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

    for (int t0 = 0; t0 < n; t0 += tile) {
        const int tn = min(tile, n - t0);
        for (int e = tid; e < tn * 3; e += kBqThreads) {
            const int k = e / 3, c = e - k * 3;
            s_pts[c * tile + k] = __ldg(pts + (size_t)t0 * 3 + e);
        }
        __syncthreads();
        bool all_done = true;
        for (int q = warp; q < nq; q += kBqWarps) {
            int cnt = s_cnt[q];
            if (cnt >= ns) continue;
            int first = s_first[q];
            const float qx = __ldg(qs + q * 3), qy = __ldg(qs + q * 3 + 1), qz = __ldg(qs + q * 3 + 2);
            int *row = rows + (size_t)q * ns;
            for (int k0 = 0; k0 < tn && cnt < ns; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                if (k < tn) {
                    const float s = d2_ref_gpu(__fsub_rn(qx, sx[k]), __fsub_rn(qy, sy[k]), __fsub_rn(qz, sz[k]));
                    hit = any && !(s > smax);
                }
                const unsigned mask = __ballot_sync(kFull, hit);
                if (mask) {
                    if (cnt == 0) first = t0 + k0 + __ffs(mask) - 1;
                    const int pos = cnt + __popc(mask & lt_mask);
                    if (hit && pos < ns) row[pos] = t0 + k;
                    cnt += __popc(mask);
                }
            }
            if (lane == 0) {
                s_cnt[q] = cnt;
                s_first[q] = first;
            }
            all_done = all_done && (cnt >= ns);
        }
        if (__syncthreads_and(all_done)) break;
    }
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
    // queries per CTA: as many as keeps >= 2 CTAs per SM in flight, 8..64
    int qpw = (int)(((long long)b * m) / ((long long)kBqWarps * 2 * kNumSMs));
    qpw = qpw < 1 ? 1 : (qpw > 8 ? 8 : qpw);
    const int qpc = qpw * kBqWarps;
    const int tile = n < kBqTile ? ((n + 31) & ~31) : kBqTile;
    const size_t smem = sizeof(float) * 3 * (size_t)tile;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(float) * 3 * kBqTile));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    count_launch();
    ball_query_kernel<<<dim3(ceil_div(m, qpc), b), kBqThreads, smem, stream>>>(n, m, radius, nsample, xyz1, xyz2, idx,
                                                                              pts_cnt, qpc, tile);
    return finish_launch();
}

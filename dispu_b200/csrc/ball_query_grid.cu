// QueryBallPoint through a uniform grid — same results as the brute-force scan (ball_query.cu, and the
// reference's query_ball_point_gpu, tf_ops/grouping/tf_grouping_g.cu:3-36) with O(candidates in the 27
// neighbouring cells) work per query instead of O(n).  Used for large clouds when the caller supplies a
// workspace (dispu_query_ball_point_ws); bit-identical idx / pts_cnt by construction:
//   * the hit test is the same arithmetic (d2_ref_gpu, s <= s_max(r));
//   * the reference takes the FIRST nsample hits in ascending index order, i.e. the nsample smallest
//     indices among all hits: every hit lies in the 3x3x3 cells around the query (cell edge >=
//     r*(1+1e-4), which absorbs all rounding), so collecting all hits there and sorting them ascending
//     gives the same row; a query with more hits than the per-warp buffer falls back to the serial scan.
//
// Pipeline (5 small launches, all HBM/L2 streaming): bbox per cloud -> cell id per point + histogram ->
// exclusive scan of the histogram -> scatter points (x,y,z,original index) into cell order -> query
// (one warp per query, 9 contiguous x-runs of cells loaded with independent requests, ballot compaction
// into shared memory, rank-by-counting of the hit indices).
#include <float.h>

#include "common.cuh"

namespace dispu {

constexpr int kGridMax = 32;                          // cells per axis (<= 32768 cells per cloud)
constexpr int kGridCells = kGridMax * kGridMax * kGridMax;
constexpr int kGqWarps = 8;
constexpr int kGqCap = 256;                           // hit buffer per query

struct GridMeta {  // per cloud
    float minx, miny, minz, inv_cell;
    int gx, gy, gz, pad;
};

__device__ __forceinline__ float ball_s_max_g(float r, bool &any) {  // same as ball_query.cu
    any = r > 1e-20f;
    if (!any) return 0.f;
    float c = __fmul_rn(r, r);
    for (int it = 0; it < 4096 && __fsqrt_rn(c) < r; ++it) c = __uint_as_float(__float_as_uint(c) + 1u);
    for (int it = 0; it < 4096 && !(__fsqrt_rn(c) < r); ++it) c = __uint_as_float(__float_as_uint(c) - 1u);
    return c;
}

__device__ __forceinline__ int cell_coord(float p, float mn, float inv, int g) {
    const float u = fminf(fmaxf(__fmul_rn(__fsub_rn(p, mn), inv), -2.f), (float)g + 1.f);
    return (int)floorf(u);
}

// ---- 1. bounding box + grid dimensions per cloud ----------------------------------------------------
__global__ void __launch_bounds__(1024)
grid_bbox_kernel(int n, const float *__restrict__ xyz, const float *__restrict__ radius, GridMeta *__restrict__ meta) {
    __shared__ float s_min[3][32], s_max[3][32];
    const int cloud = blockIdx.x, t = threadIdx.x;
    const float *p = xyz + (size_t)cloud * n * 3;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const bool vec = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
    int k0 = 0;
    if (vec) {  // 4 points = 12 floats = 3 float4 per step, fully coalesced
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        const int groups = n / 4;
        for (int g = t; g < groups; g += 1024) {
            const float4 a = __ldg(p4 + 3 * g), b4 = __ldg(p4 + 3 * g + 1), c4 = __ldg(p4 + 3 * g + 2);
            const float xs[4] = {a.x, a.w, b4.z, c4.y}, ys[4] = {a.y, b4.x, b4.w, c4.z}, zs[4] = {a.z, b4.y, c4.x, c4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                mn[0] = fminf(mn[0], xs[u]); mx[0] = fmaxf(mx[0], xs[u]);
                mn[1] = fminf(mn[1], ys[u]); mx[1] = fmaxf(mx[1], ys[u]);
                mn[2] = fminf(mn[2], zs[u]); mx[2] = fmaxf(mx[2], zs[u]);
            }
        }
        k0 = groups * 4;
    }
    for (int k = k0 + t; k < n; k += 1024) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(p + (size_t)k * 3 + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(kFull, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(kFull, mx[c], o));
        }
        if ((t & 31) == 0) {
            s_min[c][t >> 5] = mn[c];
            s_max[c][t >> 5] = mx[c];
        }
    }
    __syncthreads();
    if (t == 0) {
        float lo[3], hi[3];
        for (int c = 0; c < 3; ++c) {
            lo[c] = s_min[c][0];
            hi[c] = s_max[c][0];
            for (int w = 1; w < 32; ++w) {
                lo[c] = fminf(lo[c], s_min[c][w]);
                hi[c] = fmaxf(hi[c], s_max[c][w]);
            }
        }
        const float r = __ldg(radius);
        const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
        // cell edge: at least r*(1+1e-4), enlarged so that no axis needs more than kGridMax cells
        float cell = fmaxf(r, 0.f) * 1.0001f;
        cell = fmaxf(cell, ext / (float)kGridMax * 1.0001f);
        if (!(cell > 0.f) || !isfinite(cell)) cell = 1.f;  // degenerate cloud (all points equal) or bad radius
        GridMeta m;
        m.minx = lo[0]; m.miny = lo[1]; m.minz = lo[2];
        m.inv_cell = 1.0f / cell;
        m.gx = min(kGridMax, max(1, (int)floorf((hi[0] - lo[0]) * m.inv_cell) + 1));
        m.gy = min(kGridMax, max(1, (int)floorf((hi[1] - lo[1]) * m.inv_cell) + 1));
        m.gz = min(kGridMax, max(1, (int)floorf((hi[2] - lo[2]) * m.inv_cell) + 1));
        m.pad = 0;
        meta[cloud] = m;
    }
}

// ---- 2. cell id per point + histogram -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_count_kernel(int n, const float *__restrict__ xyz, const GridMeta *__restrict__ meta, int *__restrict__ cell_of,
                  int *__restrict__ counts) {
    const int cloud = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const GridMeta m = meta[cloud];
    const float *p = xyz + ((size_t)cloud * n + k) * 3;
    const int cx = min(m.gx - 1, max(0, cell_coord(__ldg(p), m.minx, m.inv_cell, m.gx)));
    const int cy = min(m.gy - 1, max(0, cell_coord(__ldg(p + 1), m.miny, m.inv_cell, m.gy)));
    const int cz = min(m.gz - 1, max(0, cell_coord(__ldg(p + 2), m.minz, m.inv_cell, m.gz)));
    const int cid = (cz * m.gy + cy) * m.gx + cx;
    cell_of[(size_t)cloud * n + k] = cid;
    atomicAdd(counts + (size_t)cloud * kGridCells + cid, 1);
}

// ---- 3. exclusive scan of the histogram (one CTA per cloud; only the gx*gy*gz used cells) ----------------
// Warp w scans the contiguous slice [w*per, (w+1)*per) of cells in rows of 32 (coalesced), then the 32 warp
// totals are scanned and added.  `counts` is left zeroed: it becomes the scatter cursor.
__global__ void __launch_bounds__(1024)
grid_scan_kernel(const GridMeta *__restrict__ meta, int *__restrict__ counts, int *__restrict__ starts) {
    __shared__ int s_warp[32];
    const int cloud = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const GridMeta gm = meta[cloud];
    const int ncell = gm.gx * gm.gy * gm.gz;
    const int per = ((ncell + 31) / 32 + 31) & ~31;  // cells per warp, multiple of 32
    int *c = counts + (size_t)cloud * kGridCells;
    int *s = starts + (size_t)cloud * (kGridCells + 1);
    const int lo = warp * per, hi = min(ncell, lo + per);
    int carry = 0;
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        const int v = e < hi ? c[e] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += u;
        }
        if (e < hi) s[e] = carry + inc - v;  // exclusive, warp-local
        carry += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) s_warp[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        const int tot = s_warp[lane];
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += u;
        }
        s_warp[lane] = inc - tot;  // exclusive offset of each warp's slice
        if (lane == 31) s[ncell] = inc;
    }
    __syncthreads();
    const int off = s_warp[warp];
    for (int e = lo + lane; e < hi; e += 32) {
        s[e] += off;
        c[e] = 0;
    }
}

// ---- 4. scatter points into cell order -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_scatter_kernel(int n, const float *__restrict__ xyz, const int *__restrict__ cell_of, const int *__restrict__ starts,
                    int *__restrict__ cursor, float4 *__restrict__ sorted) {
    const int cloud = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const int cid = cell_of[(size_t)cloud * n + k];
    const int pos = starts[(size_t)cloud * (kGridCells + 1) + cid] + atomicAdd(cursor + (size_t)cloud * kGridCells + cid, 1);
    const float *p = xyz + ((size_t)cloud * n + k) * 3;
    sorted[(size_t)cloud * n + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float(k));
}

// ---- 5. query ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGqWarps * 32)
grid_query_kernel(int n, int m, const float *__restrict__ radius, int ns, const float *__restrict__ xyz1,
                  const float *__restrict__ xyz2, const GridMeta *__restrict__ meta, const int *__restrict__ starts,
                  const float4 *__restrict__ sorted, int *__restrict__ idx, int *__restrict__ pts_cnt) {
    __shared__ int s_hits[kGqWarps][kGqCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    const int q = blockIdx.x * kGqWarps + warp;
    if (q >= m) return;
    bool any;
    const float smax = ball_s_max_g(__ldg(radius), any);
    const unsigned lt_mask = (1u << lane) - 1u;
    const GridMeta gm = meta[cloud];
    const float *qp = xyz2 + ((size_t)cloud * m + q) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const int *st = starts + (size_t)cloud * (kGridCells + 1);
    const float4 *pts = sorted + (size_t)cloud * n;
    int *hits = s_hits[warp];
    int *row = idx + ((size_t)cloud * m + q) * ns;

    int H = 0;
    if (any) {
        const int cx = cell_coord(qx, gm.minx, gm.inv_cell, gm.gx);
        const int cy = cell_coord(qy, gm.miny, gm.inv_cell, gm.gy);
        const int cz = cell_coord(qz, gm.minz, gm.inv_cell, gm.gz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, gm.gx - 1);
        // lane r < 9 owns the contiguous x-run of cells (x0..x1, cy-1+r%3, cz-1+r/3): all 9 (begin,end)
        // pairs are fetched with ONE round of independent loads
        int beg = 0, len = 0;
        if (lane < 9 && x0 <= x1) {
            const int y = cy - 1 + lane % 3, z = cz - 1 + lane / 3;
            if (y >= 0 && y < gm.gy && z >= 0 && z < gm.gz) {
                const int base = (z * gm.gy + y) * gm.gx;
                beg = __ldg(st + base + x0);
                len = __ldg(st + base + x1 + 1) - beg;
            }
        }
        // exclusive prefix of the run lengths over lanes 0..8 -> flat candidate numbering
        int inc = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int u = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += u;
        }
        const int total = __shfl_sync(kFull, inc, 8);
        const int excl = inc - len;
        (void)total; (void)excl;
        // first 32 candidates of every run: 9 independent loads in flight, then the tests
        float4 c9[9];
        int l9[9], b9[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            l9[r] = __shfl_sync(kFull, len, r);
            b9[r] = __shfl_sync(kFull, beg, r);
            c9[r] = lane < l9[r] ? __ldg(pts + b9[r] + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            bool hit = false;
            if (lane < l9[r]) {
                const float s2 = d2_ref_gpu(__fsub_rn(qx, c9[r].x), __fsub_rn(qy, c9[r].y), __fsub_rn(qz, c9[r].z));
                hit = !(s2 > smax);
            }
            const unsigned mask = __ballot_sync(kFull, hit);
            const int hp = H + __popc(mask & lt_mask);
            if (hit && hp < kGqCap) hits[hp] = __float_as_int(c9[r].w);
            H += __popc(mask);
            for (int o = 32; o < l9[r]; o += 32) {  // runs longer than a warp (dense cells): rare
                bool hit2 = false;
                int k2 = 0;
                if (o + lane < l9[r]) {
                    const float4 c = __ldg(pts + b9[r] + o + lane);
                    const float s2 = d2_ref_gpu(__fsub_rn(qx, c.x), __fsub_rn(qy, c.y), __fsub_rn(qz, c.z));
                    hit2 = !(s2 > smax);
                    k2 = __float_as_int(c.w);
                }
                const unsigned m2 = __ballot_sync(kFull, hit2);
                const int hp2 = H + __popc(m2 & lt_mask);
                if (hit2 && hp2 < kGqCap) hits[hp2] = k2;
                H += __popc(m2);
            }
        }
    }
    __syncwarp();
    int cnt;
    if (H > kGqCap) {
        // more hits than the buffer holds: the serial definition, 32 candidates per step
        const float *pall = xyz1 + (size_t)cloud * n * 3;
        cnt = 0;
        int first = 0;
        for (int k0 = 0; k0 < n && cnt < ns; k0 += 32) {
            const int k = k0 + lane;
            bool hit = false;
            if (k < n) {
                const float s = d2_ref_gpu(__fsub_rn(qx, __ldg(pall + (size_t)k * 3)), __fsub_rn(qy, __ldg(pall + (size_t)k * 3 + 1)),
                                           __fsub_rn(qz, __ldg(pall + (size_t)k * 3 + 2)));
                hit = !(s > smax);
            }
            const unsigned mask = __ballot_sync(kFull, hit);
            if (mask) {
                if (cnt == 0) first = k0 + __ffs(mask) - 1;
                const int pos = cnt + __popc(mask & lt_mask);
                if (hit && pos < ns) row[pos] = k;
                cnt += __popc(mask);
            }
        }
        cnt = min(cnt, ns);
        for (int l = cnt + lane; l < ns; l += 32) row[l] = first;
    } else {
        // order the hit indices: rank(h) = number of hits smaller than h (indices are distinct); every lane
        // ranks its hits against broadcast reads of the whole buffer and scatters them into the row.
        cnt = min(H, ns);
        int first = 0x7fffffff;
        for (int i0 = 0; i0 < H; i0 += 32) {
            const int i = i0 + lane;
            const int h = i < H ? hits[i] : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < H; ++j) rank += hits[j] < h;
            if (i < H && rank < ns) row[rank] = h;
            first = min(first, h);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(kFull, first, o));
        const int fill = cnt > 0 ? first : 0;  // rows without a hit: zeros (see ball_query.cu)
        for (int l = cnt + lane; l < ns; l += 32) row[l] = fill;
    }
    if (lane == 0) pts_cnt[(size_t)cloud * m + q] = cnt;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_query_ball_point_workspace_bytes(int b, int n, int m) {
    (void)m;
    if (b <= 0 || n <= 0) return 0;
    return align256(sizeof(GridMeta) * (size_t)b) + align256(sizeof(int) * (size_t)b * n) +
           align256(sizeof(int) * (size_t)b * kGridCells) + align256(sizeof(int) * (size_t)b * (kGridCells + 1)) +
           align256(sizeof(float4) * (size_t)b * n);
}

extern "C" int dispu_query_ball_point_ws(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                                         const float *xyz2, int *idx, int *pts_cnt, void *workspace,
                                         size_t workspace_bytes, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(nsample > 0, DISPU_EINVAL);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(radius && xyz1 && xyz2 && idx && pts_cnt, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    // small clouds: the staged brute-force scan is already cheap and needs no workspace
    if (n < 2048 || workspace == nullptr) return dispu_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream_);
    DISPU_REQUIRE(workspace_bytes >= dispu_query_ball_point_workspace_bytes(b, n, m), DISPU_EWORKSPACE);
    uint8_t *w = reinterpret_cast<uint8_t *>(workspace);
    GridMeta *meta = reinterpret_cast<GridMeta *>(w);
    w += align256(sizeof(GridMeta) * (size_t)b);
    int *cell_of = reinterpret_cast<int *>(w);
    w += align256(sizeof(int) * (size_t)b * n);
    int *counts = reinterpret_cast<int *>(w);
    w += align256(sizeof(int) * (size_t)b * kGridCells);
    int *starts = reinterpret_cast<int *>(w);
    w += align256(sizeof(int) * (size_t)b * (kGridCells + 1));
    float4 *sorted = reinterpret_cast<float4 *>(w);
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)b * kGridCells, stream);
    if (e != cudaSuccess) return (int)e;
    count_launch(5);
    grid_bbox_kernel<<<b, 1024, 0, stream>>>(n, xyz1, radius, meta);
    grid_count_kernel<<<dim3(ceil_div(n, 256), b), 256, 0, stream>>>(n, xyz1, meta, cell_of, counts);
    grid_scan_kernel<<<b, 1024, 0, stream>>>(meta, counts, starts);
    grid_scatter_kernel<<<dim3(ceil_div(n, 256), b), 256, 0, stream>>>(n, xyz1, cell_of, starts, counts, sorted);
    grid_query_kernel<<<dim3(ceil_div(m, kGqWarps), b), kGqWarps * 32, 0, stream>>>(n, m, radius, nsample, xyz1, xyz2, meta,
                                                                                   starts, sorted, idx, pts_cnt);
    return finish_launch();
}

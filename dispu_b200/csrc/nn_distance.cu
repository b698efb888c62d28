// NnDistance (+Grad) and the fused Chamfer/Hausdorff reduction for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel (tf_ops/nn_distance/tf_nndistance_g.cu:5-157)
// and, for dispu_chamfer, the six TF reduce ops of chamfer()/hausdorff_loss()
// (Common/loss_utils.py:45-84) that follow them.
//
// The reference launches the search twice (once per direction, (32,16) CTAs each, :129-130)
// with a global-memory read-modify-write of the running result per 512-point tile (:119).
// Here both directions run in ONE launch (blockIdx.z), a thread keeps its query's running
// (min, argmin) in registers across all tiles, candidates are staged as float4 tiles so the
// inner loop is one broadcast LDS.128 + 9 FP32 ops per pair, and the grid covers queries x
// clouds x directions.  First minimum wins (strict '<' in ascending candidate order), which is
// what the reference's tile logic amounts to.  FMA contraction = d2_ref_gpu.
//
// dispu_chamfer never writes the (b,n)/(b,m) distance tensors: each CTA reduces its queries'
// minima to (sum, max) in a fixed order and stores one partial; a second tiny kernel reduces the
// partials in a fixed order -> per-cloud cd/hd and the batch mean / max / sum.  No float atomics,
// so results are run-to-run deterministic.
#include <stdlib.h>

#include "common.cuh"
#include "tile_stage.cuh"

namespace dispu {

constexpr int kNnThreads = 256;
constexpr int kNnTile = 1024;  // float4 candidates per tile (16 KB)

// One thread = one query of direction `dir` (0: xyz1 -> xyz2, 1: xyz2 -> xyz1).
// kTma: candidate tiles land as raw AoS bytes by cp.async.bulk + mbarrier, double-buffered (tile_stage.cuh), and are
// widened to the float4 tile the inner loop broadcasts; otherwise per-thread __ldg staging.
template <bool kTma>
__global__ void __launch_bounds__(kNnThreads)
nn_distance_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                   float *__restrict__ dist1, int *__restrict__ idx1, float *__restrict__ dist2,
                   int *__restrict__ idx2) {
    __shared__ float4 s_tile[kNnTile];
    __shared__ __align__(16) float s_raw[kTma ? 2 * kNnTile * 3 : 4];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int cloud = blockIdx.y, dir = blockIdx.z;
    const int nq = dir == 0 ? n : m, nc = dir == 0 ? m : n;
    if ((int)blockIdx.x * kNnThreads >= nq) return;  // this direction has fewer query chunks
    const float *qs = (dir == 0 ? xyz1 : xyz2) + (size_t)cloud * nq * 3;
    const float *cs = (dir == 0 ? xyz2 : xyz1) + (size_t)cloud * nc * 3;
    const int ntiles = ceil_div(nc, kNnTile);
    if (kTma) {
        if (threadIdx.x == 0) {
            ts::mbar_init(&s_bar[0], 1);
            ts::mbar_init(&s_bar[1], 1);
            ts::fence_mbar_init();
            ts::issue_tile(s_raw, cs, min(kNnTile, nc), &s_bar[0]);
        }
        __syncthreads();
    }
    const int j = blockIdx.x * kNnThreads + threadIdx.x;
    const bool active = j < nq;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        qx = __ldg(qs + (size_t)j * 3); qy = __ldg(qs + (size_t)j * 3 + 1); qz = __ldg(qs + (size_t)j * 3 + 2);
    }
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    for (int i = 0; i < ntiles; ++i) {
        const int t0 = i * kNnTile;
        const int tn = min(kNnTile, nc - t0);
        if (i > 0) __syncthreads();
        if (kTma) {
            if (threadIdx.x == 0 && i + 1 < ntiles)
                ts::issue_tile(s_raw + ((i + 1) & 1) * kNnTile * 3, cs + (size_t)(t0 + kNnTile) * 3, min(kNnTile, nc - t0 - kNnTile),
                               &s_bar[(i + 1) & 1]);
            ts::mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));
            const float *raw = s_raw + (i & 1) * kNnTile * 3;
            for (int e = threadIdx.x; e < tn; e += kNnThreads) s_tile[e] = make_float4(raw[e * 3], raw[e * 3 + 1], raw[e * 3 + 2], 0.f);
        } else {
            for (int e = threadIdx.x; e < tn; e += kNnThreads) {
                const float *p = cs + (size_t)(t0 + e) * 3;
                s_tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            }
        }
        __syncthreads();
        if (t0 == 0) {  // k==0 initialises unconditionally (tf_nndistance_g.cu:29), even for NaN/inf
            const float4 p = s_tile[0];
            best = d2_ref_gpu(__fsub_rn(p.x, qx), __fsub_rn(p.y, qy), __fsub_rn(p.z, qz));
            bi = 0;
        }
#pragma unroll 8
        for (int e = 0; e < tn; ++e) {
            const float4 p = s_tile[e];
            const float d = d2_ref_gpu(__fsub_rn(p.x, qx), __fsub_rn(p.y, qy), __fsub_rn(p.z, qz));
            if (d < best) {
                best = d;
                bi = t0 + e;
            }
        }
    }
    if (active) {
        (dir == 0 ? dist1 : dist2)[(size_t)cloud * nq + j] = best;
        (dir == 0 ? idx1 : idx2)[(size_t)cloud * nq + j] = bi;
    }
}

// ---- minimum-only search for the fused Chamfer/Hausdorff path ------------------------------------------------
// chamfer()/hausdorff_loss() consume only the distances, so the arg-min bookkeeping (a compare and two selects per
// pair) is dead weight there.  Candidates are staged as x / y / z arrays so that two neighbouring candidates load as
// one 64-bit pair, and the whole distance runs on packed FP32 (sub/mul/fma .f32x2: one issue slot, two IEEE results,
// bit-identical to d2_ref_gpu per half); the running minimum is one FMNMX per candidate.  NaN semantics of the
// reference (candidate 0 initialises unconditionally, later NaNs never win) are kept by handling candidate 0 apart.
typedef unsigned long long nn_f32x2;
__device__ __forceinline__ nn_f32x2 nn_pack(float lo, float hi) {
    nn_f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_sub2(nn_f32x2 a, nn_f32x2 b) {
    nn_f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_mul2(nn_f32x2 a, nn_f32x2 b) {
    nn_f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_fma2(nn_f32x2 a, nn_f32x2 b, nn_f32x2 c) {
    nn_f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Shape of the launch: a CTA of four warps owns kNmQueries = 32 * kNmQt queries; lane l holds kNmQt of them in registers
// (queries l, l+32, ...), and each WARP scans its own quarter of every candidate tile, so one broadcast LDS.64 triple
// (x, y, z of a candidate pair) feeds kNmQt * 2 distance evaluations instead of 2: the round-1 form (one query per
// thread, every thread scanning all candidates) was bound by the shared-memory pipe, not by FP32 (7.8 pair evaluations
// per clock and SM).  The four partial minima of a query meet in shared memory.  Small CTAs on purpose: 512 CTAs at the
// c2 shape = 3.5 per SM, all resident at once.
// kTma: candidate tiles arrive by cp.async.bulk into a raw AoS landing buffer (double-buffered: the copy of tile i+1
// flies while tile i is searched), completion on an mbarrier (tile_stage.cuh); the CTA then transposes the tile to
// the x / y / z arrays the packed inner loop reads.  Otherwise the threads stage with __ldg.
constexpr int kNmThreads = 128;
constexpr int kNmQt = 4;
constexpr int kNmQueries = 32 * kNmQt;

template <bool kTma>
__global__ void __launch_bounds__(kNmThreads)
nn_min_reduce_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                     float2 *__restrict__ partials, int chunks) {
    extern __shared__ __align__(16) float s_dyn[];  // x | y | z (kNnTile + 2 each, padded to 16 B) [| raw 0 | raw 1]
    constexpr int kArr = (kNnTile + 2 + 3) & ~3;
    float *s_x = s_dyn, *s_y = s_dyn + kArr, *s_z = s_dyn + 2 * kArr;
    float *s_raw = s_dyn + 3 * kArr;
    __shared__ float s_part[kNmThreads / 32][kNmQueries];
    __shared__ float s_red[2][kNmThreads / 32];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int cloud = blockIdx.y, dir = blockIdx.z;
    const int nq = dir == 0 ? n : m, nc = dir == 0 ? m : n;
    if ((int)blockIdx.x * kNmQueries >= nq) {  // this direction has fewer query chunks
        if (threadIdx.x == 0 && (int)blockIdx.x < chunks)
            partials[((size_t)cloud * 2 + dir) * chunks + blockIdx.x] = make_float2(0.f, -1.f);
        return;
    }
    const float *qs = (dir == 0 ? xyz1 : xyz2) + (size_t)cloud * nq * 3;
    const float *cs = (dir == 0 ? xyz2 : xyz1) + (size_t)cloud * nc * 3;
    const int ntiles = ceil_div(nc, kNnTile);
    if (kTma) {
        if (threadIdx.x == 0) {
            ts::mbar_init(&s_bar[0], 1);
            ts::mbar_init(&s_bar[1], 1);
            ts::fence_mbar_init();
            ts::issue_tile(s_raw, cs, min(kNnTile, nc), &s_bar[0]);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * kNmQueries;
    nn_f32x2 qx2[kNmQt], qy2[kNmQt], qz2[kNmQt];
    float best[kNmQt];
#pragma unroll
    for (int u = 0; u < kNmQt; ++u) {
        const int j = min(q0 + u * 32 + lane, nq - 1);
        const float qx = __ldg(qs + (size_t)j * 3), qy = __ldg(qs + (size_t)j * 3 + 1), qz = __ldg(qs + (size_t)j * 3 + 2);
        qx2[u] = nn_pack(qx, qx); qy2[u] = nn_pack(qy, qy); qz2[u] = nn_pack(qz, qz);
        best[u] = __int_as_float(0x7f800000);
    }
    for (int i = 0; i < ntiles; ++i) {
        const int t0 = i * kNnTile;
        const int tn = min(kNnTile, nc - t0);
        const int tn2 = (tn + 1) & ~1;  // odd tail: the last candidate is staged twice (the minimum is unchanged)
        if (i > 0) __syncthreads();     // everyone has finished searching the previous tile's x / y / z
        if (kTma) {
            // the other landing buffer was transposed in iteration i-1, before the barrier above: refill it now
            if (threadIdx.x == 0 && i + 1 < ntiles)
                ts::issue_tile(s_raw + ((i + 1) & 1) * kNnTile * 3, cs + (size_t)(t0 + kNnTile) * 3, min(kNnTile, nc - t0 - kNnTile),
                               &s_bar[(i + 1) & 1]);
            ts::mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));
            const float *raw = s_raw + (i & 1) * kNnTile * 3;
            for (int e = threadIdx.x; e < tn2; e += kNmThreads) {  // stride-3 reads: conflict-free (3 is odd)
                const float *p = raw + min(e, tn - 1) * 3;
                s_x[e] = p[0]; s_y[e] = p[1]; s_z[e] = p[2];
            }
        } else {
            for (int e = threadIdx.x; e < tn2; e += kNmThreads) {
                const float *p = cs + (size_t)(t0 + min(e, tn - 1)) * 3;
                s_x[e] = __ldg(p); s_y[e] = __ldg(p + 1); s_z[e] = __ldg(p + 2);
            }
        }
        __syncthreads();
        const nn_f32x2 *x2 = reinterpret_cast<const nn_f32x2 *>(s_x), *y2 = reinterpret_cast<const nn_f32x2 *>(s_y),
                       *z2 = reinterpret_cast<const nn_f32x2 *>(s_z);
        const int npairs = tn2 / 2;
        const int e_lo = (npairs * warp) / (kNmThreads / 32), e_hi = (npairs * (warp + 1)) / (kNmThreads / 32);
#pragma unroll 4
        for (int e = e_lo; e < e_hi; ++e) {
            const nn_f32x2 cx = x2[e], cy = y2[e], cz = z2[e];
#pragma unroll
            for (int u = 0; u < kNmQt; ++u) {
                const nn_f32x2 dx = nn_sub2(cx, qx2[u]), dy = nn_sub2(cy, qy2[u]), dz = nn_sub2(cz, qz2[u]);
                const nn_f32x2 d = nn_fma2(dz, dz, nn_fma2(dx, dx, nn_mul2(dy, dy)));
                float da, db;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(da), "=f"(db) : "l"(d));
                best[u] = fminf(best[u], fminf(da, db));
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kNmQt; ++u) s_part[warp][u * 32 + lane] = best[u];
    __syncthreads();
    // thread t = query q0 + t: merge the four warps' minima; candidate 0 apart (tf_nndistance_g.cu:29 initialises with it
    // unconditionally): the result is d0 if d0 is NaN, else the minimum (fminf drops later NaNs, like '<' does)
    const int j = q0 + threadIdx.x;
    const bool active = threadIdx.x < kNmQueries && j < nq;
    float bq = 0.f;
    if (active) {
        const float qx = __ldg(qs + (size_t)j * 3), qy = __ldg(qs + (size_t)j * 3 + 1), qz = __ldg(qs + (size_t)j * 3 + 2);
        const float d0 = d2_ref_gpu(__fsub_rn(__ldg(cs), qx), __fsub_rn(__ldg(cs + 1), qy), __fsub_rn(__ldg(cs + 2), qz));
        bq = s_part[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kNmThreads / 32; ++w) bq = fminf(bq, s_part[w][threadIdx.x]);
        bq = (d0 != d0) ? d0 : fminf(bq, d0);
    }
    float s = active ? bq : 0.f;
    float mx = active ? bq : -1.f;  // distances are >= 0
    s = warp_sum(s);
    mx = warp_max(mx);
    if (lane == 0) {
        s_red[0][warp] = s;
        s_red[1][warp] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts_ = 0.f, tm = -1.f;
#pragma unroll
        for (int w = 0; w < kNmThreads / 32; ++w) {
            ts_ += s_red[0][w];
            tm = fmaxf(tm, s_red[1][w]);
        }
        partials[((size_t)cloud * 2 + dir) * chunks + blockIdx.x] = make_float2(ts_, tm);
    }
}

// partials (b,2,chunks) -> per_cloud (b,2) and out[4]; one CTA, fixed order.
__global__ void __launch_bounds__(256)
chamfer_finalize_kernel(int b, int n_gt, int n_pred, int chunks, const float2 *__restrict__ partials,
                        const float *__restrict__ radius, float *__restrict__ out, float *__restrict__ per_cloud,
                        float *__restrict__ scratch /* (b,2) */, float *__restrict__ sum_out, float *__restrict__ max_out) {
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
        float sf = 0.f, sb = 0.f, mf = -1.f, mb = -1.f;
        for (int c = 0; c < chunks; ++c) {
            const float2 f = partials[((size_t)i * 2 + 0) * chunks + c];
            const float2 g = partials[((size_t)i * 2 + 1) * chunks + c];
            sf += f.x; mf = fmaxf(mf, f.y);
            sb += g.x; mb = fmaxf(mb, g.y);
        }
        const float r = radius ? radius[i] : 1.0f;
        // loss_utils.py:59-62: (mean fwd + mean bwd) / radius ; :79-82 with max
        const float cd = __fdiv_rn(__fadd_rn(__fdiv_rn(sf, (float)n_gt), __fdiv_rn(sb, (float)n_pred)), r);
        const float hd = __fdiv_rn(__fadd_rn(mf, mb), r);
        scratch[i * 2] = cd;
        scratch[i * 2 + 1] = hd;
        if (per_cloud) {
            per_cloud[i * 2] = cd;
            per_cloud[i * 2 + 1] = hd;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f, mx = -__int_as_float(0x7f800000);
        for (int i = 0; i < b; ++i) {
            sum += scratch[i * 2];
            mx = fmaxf(mx, scratch[i * 2 + 1]);
        }
        out[0] = __fdiv_rn(sum, (float)b);
        out[1] = mx;
        out[2] = sum;
        out[3] = (float)b;
        if (sum_out) *sum_out = sum;  // this rank's partial of the batch mean (the operand of the multi-GPU all-reduce)
        if (max_out) *max_out = mx;
    }
}

__global__ void __launch_bounds__(256)
nn_distance_grad_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                        const float *__restrict__ grad_dist1, const int *__restrict__ idx1,
                        const float *__restrict__ grad_dist2, const int *__restrict__ idx2,
                        float *__restrict__ grad_xyz1, float *__restrict__ grad_xyz2) {
    const int cloud = blockIdx.y, dir = blockIdx.z;
    const int na = dir == 0 ? n : m, nb = dir == 0 ? m : n;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= na) return;
    const float *a = (dir == 0 ? xyz1 : xyz2) + (size_t)cloud * na * 3;
    const float *bb = (dir == 0 ? xyz2 : xyz1) + (size_t)cloud * nb * 3;
    float *ga = (dir == 0 ? grad_xyz1 : grad_xyz2) + (size_t)cloud * na * 3;
    float *gb = (dir == 0 ? grad_xyz2 : grad_xyz1) + (size_t)cloud * nb * 3;
    const int j2 = (dir == 0 ? idx1 : idx2)[(size_t)cloud * na + j];
    const float g = __fmul_rn((dir == 0 ? grad_dist1 : grad_dist2)[(size_t)cloud * na + j], 2.0f);
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const float t = __fmul_rn(g, __fsub_rn(a[(size_t)j * 3 + l], bb[(size_t)j2 * 3 + l]));
        atomicAdd(ga + (size_t)j * 3 + l, t);
        atomicAdd(gb + (size_t)j2 * 3 + l, -t);
    }
}

static inline int nn_chunks(int n, int m) { return ceil_div(n > m ? n : m, kNmQueries); }

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                                 float *dist2, int *idx2, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m > 0, DISPU_EINVAL);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    static const bool no_tma = getenv("DISPU_NO_TMA") != nullptr;  // development switch (A/B timing)
    const dim3 grid(ceil_div(n > m ? n : m, kNnThreads), b, 2);
    count_launch();
    if (!no_tma && ts::tma_ok(xyz1, n) && ts::tma_ok(xyz2, m))
        nn_distance_kernel<true><<<grid, kNnThreads, 0, stream>>>(n, m, xyz1, xyz2, dist1, idx1, dist2, idx2);
    else
        nn_distance_kernel<false><<<grid, kNnThreads, 0, stream>>>(n, m, xyz1, xyz2, dist1, idx1, dist2, idx2);
    return finish_launch();
}

extern "C" int dispu_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                      const float *grad_dist1, const int *idx1, const float *grad_dist2,
                                      const int *idx2, float *grad_xyz1, float *grad_xyz2, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m > 0, DISPU_EINVAL);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    cudaError_t e = cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, stream);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    nn_distance_grad_kernel<<<dim3(ceil_div(n > m ? n : m, 256), b, 2), 256, 0, stream>>>(
        n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
    return finish_launch();
}

extern "C" size_t dispu_chamfer_workspace_bytes(int b, int n_gt, int n_pred) {
    if (b <= 0 || n_gt <= 0 || n_pred <= 0) return 0;
    return sizeof(float2) * (size_t)b * 2 * nn_chunks(n_gt, n_pred) + sizeof(float) * 2 * (size_t)b;
}

static int chamfer_impl(int b, int n_gt, const float *gt, int n_pred, const float *pred, const float *radius, float *out,
                        float *per_cloud, float *sum_out, float *max_out, void *workspace, size_t workspace_bytes,
                        dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b > 0 && n_gt > 0 && n_pred > 0, DISPU_EINVAL);
    DISPU_REQUIRE(gt && pred && out, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    DISPU_REQUIRE(workspace && workspace_bytes >= dispu_chamfer_workspace_bytes(b, n_gt, n_pred), DISPU_EWORKSPACE);
    const int chunks = nn_chunks(n_gt, n_pred);
    float2 *partials = reinterpret_cast<float2 *>(workspace);
    float *scratch = reinterpret_cast<float *>(partials + (size_t)b * 2 * chunks);
    // both clouds are candidate sets (one per direction): bulk copies need both to qualify
    static const bool no_tma = getenv("DISPU_NO_TMA") != nullptr;  // development switch (A/B timing)
    const bool tma = !no_tma && ts::tma_ok(gt, n_gt) && ts::tma_ok(pred, n_pred);
    constexpr int kArr = (kNnTile + 2 + 3) & ~3;
    const int ntiles = ceil_div(n_gt > n_pred ? n_gt : n_pred, kNnTile);
    const size_t smem = sizeof(float) * (3 * (size_t)kArr + (tma ? (ntiles > 1 ? 2 : 1) * (size_t)kNnTile * 3 : 0));
    count_launch(2);
    // nn_distance(gt, pred): dists_forward belongs to gt (loss_utils.py:51)
    if (tma)
        nn_min_reduce_kernel<true><<<dim3(chunks, b, 2), kNmThreads, smem, stream>>>(n_gt, n_pred, gt, pred, partials, chunks);
    else
        nn_min_reduce_kernel<false><<<dim3(chunks, b, 2), kNmThreads, smem, stream>>>(n_gt, n_pred, gt, pred, partials, chunks);
    chamfer_finalize_kernel<<<1, 256, 0, stream>>>(b, n_gt, n_pred, chunks, partials, radius, out, per_cloud, scratch, sum_out,
                                                   max_out);
    return finish_launch();
}

extern "C" int dispu_chamfer(int b, int n_gt, const float *gt, int n_pred, const float *pred, const float *radius,
                             float *out, float *per_cloud, void *workspace, size_t workspace_bytes,
                             dispu_stream_t stream) {
    return chamfer_impl(b, n_gt, gt, n_pred, pred, radius, out, per_cloud, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int dispu_chamfer_ex(int b, int n_gt, const float *gt, int n_pred, const float *pred, const float *radius,
                                float *out, float *per_cloud, float *sum_out, float *max_out, void *workspace,
                                size_t workspace_bytes, dispu_stream_t stream) {
    return chamfer_impl(b, n_gt, gt, n_pred, pred, radius, out, per_cloud, sum_out, max_out, workspace, workspace_bytes, stream);
}

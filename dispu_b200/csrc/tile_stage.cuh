// TMA staging of point tiles for the SIMT point kernels (NnDistance / Chamfer, kNN / ThreeNN, ball query):
// one elected thread issues a bulk asynchronous copy (cp.async.bulk, UBLKCP in SASS) of the tile's raw AoS bytes
// (12 B per point) from global into shared memory and the copy reports its bytes on an mbarrier; the CTA's threads
// wait on the barrier's phase instead of each issuing and waiting for their own global loads.  With two landing
// buffers the copy of tile i+1 is in flight while tile i is being searched (double buffering), which is what the
// reference's per-tile "load, __syncthreads, scan, __syncthreads" loops (tf_nndistance_g.cu:13-27,
// tf_grouping_g.cu:14-34) cannot do.
//
// Bulk copies need a 16-byte aligned source and a size that is a multiple of 16 bytes: a cloud qualifies when its
// base pointer is 16-byte aligned and n % 4 == 0 (then every cloud of the batch and every tile start is aligned and
// every tile length is a multiple of 4 points); other clouds take the kernels' per-thread __ldg staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dispu {
namespace ts {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy; completion = `bytes` transaction bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// host side: can the clouds of this (b, n, 3) tensor be staged by bulk copies?
static inline bool tma_ok(const void *base, int n) { return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (n % 4) == 0; }

// Issue the copy of `npts` points (npts % 4 == 0) starting at `src` into `dst`; call from ONE thread.
__device__ __forceinline__ void issue_tile(float *dst, const float *src, int npts, uint64_t *bar) {
    const uint32_t bytes = (uint32_t)npts * 12u;
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(dst, src, bytes, bar);
}

// ---- float4 tile staging shared by the kNN / ThreeNN kernels ------------------------------------------------------
// Protocol (all threads of the CTA call these in the same order):
//   stage_begin(...)                      once, before the tile loop: barrier init + copy of tile 0
//   for i in 0..ntiles-1:  stage_tile(i)  -> s_tile[0..tn) = points [i*kTile, i*kTile+tn) as float4 (w = 0)
// stage_tile contains the leading __syncthreads (the previous tile's readers are done), the refill of the other
// landing buffer with tile i+1, the wait for tile i, the AoS -> float4 widening and the trailing __syncthreads.
template <bool kTma>
__device__ __forceinline__ void stage_begin(float *s_raw, uint64_t *s_bar, const float *pts, int n, int tile_pts) {
    if (kTma) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            fence_mbar_init();
            issue_tile(s_raw, pts, n < tile_pts ? n : tile_pts, &s_bar[0]);
        }
        __syncthreads();
    }
}
template <bool kTma>
__device__ __forceinline__ int stage_tile(float4 *s_tile, float *s_raw, uint64_t *s_bar, const float *pts, int n, int tile_pts,
                                          int i, int ntiles) {
    const int t0 = i * tile_pts;
    const int tn = (n - t0) < tile_pts ? (n - t0) : tile_pts;
    if (i > 0) __syncthreads();
    if (kTma) {
        if (threadIdx.x == 0 && i + 1 < ntiles) {
            const int rest = n - t0 - tile_pts;
            issue_tile(s_raw + ((i + 1) & 1) * tile_pts * 3, pts + (size_t)(t0 + tile_pts) * 3, rest < tile_pts ? rest : tile_pts,
                       &s_bar[(i + 1) & 1]);
        }
        mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));
        const float *raw = s_raw + (i & 1) * tile_pts * 3;
        for (int e = threadIdx.x; e < tn; e += blockDim.x) s_tile[e] = make_float4(raw[e * 3], raw[e * 3 + 1], raw[e * 3 + 2], 0.f);
    } else {
        for (int e = threadIdx.x; e < tn; e += blockDim.x) {
            const float *p = pts + (size_t)(t0 + e) * 3;
            s_tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
    }
    __syncthreads();
    return tn;
}

}  // namespace ts
}  // namespace dispu

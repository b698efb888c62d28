// Farthest point sampling for sm_100a — replaces farthestpointsamplingKernel
// (tf_ops/sampling/tf_sampling_g.cu:105-170) behind farthestpointsamplingLauncher (:203).
//
// Design (not a translation): the reference keeps the running min-distance in GLOBAL memory
// and re-reads it (plus xyz beyond 3072 points) on every one of the m-1 dependent steps, with
// a 10-barrier shared-memory tree per step and at most one CTA per cloud.  Here a cloud lives
// ON CHIP for the whole run: every thread owns P points (x,y,z and the running min-distance in
// registers), a float4 copy of the cloud sits in shared memory only to look up the winner's
// coordinates, the per-step arg-max is two `redux.sync` per warp + one barrier + two more
// `redux.sync`, and clouds too large for one SM's register file (or too few to fill the GPU)
// are spread over a thread-block CLUSTER whose CTAs exchange their local winners through
// distributed shared memory through asynchronous DSMEM stores that complete transaction bytes on the
// receiver's mbarrier (no cluster-wide barrier inside the loop).
//
// Bit-exactness: indices must equal the reference's.  The reference's order among equal maxima
// is (k mod 512) ascending, then k ascending (strided per-thread scan with strict '>' at :146,
// tree reduction keeping the lower thread id at :158).  That order is carried explicitly as a
// 32-bit tie key next to the float's bit pattern, so the thread mapping is free.  Distances use
// the reference's FMA contraction (d2_ref_gpu).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dispu {

__device__ __forceinline__ uint32_t fps_tie_of(uint32_t k) { return ((k & 511u) << 23) | (k >> 9); }
__device__ __forceinline__ uint32_t fps_k_of_tie(uint32_t t) { return (t >> 23) | ((t & 0x7fffffu) << 9); }

constexpr uint32_t kTieNone = 0xffffffffu;

// (bits desc, tie asc) arg-max over a warp; every lane gets the winner.
__device__ __forceinline__ void warp_argmax(uint32_t bits, uint32_t tie, uint32_t &wb, uint32_t &wt) {
    wb = __reduce_max_sync(kFull, bits);
    wt = __reduce_min_sync(kFull, bits == wb ? tie : kTieNone);
}

// ---- cluster exchange primitives (DSMEM + remote mbarrier arrive) ---------------------------------
__device__ __forceinline__ uint32_t fps_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fps_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_smem_u32(bar)), "r"(count) : "memory");
}
// map a shared::cta address of this CTA to the shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t fps_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(saddr), "r"(rank));
    return raddr;
}
// Asynchronous remote stores that complete transaction bytes on the DESTINATION CTA's mbarrier: data and
// "it has arrived" travel together, so no release fence / second round trip is needed.
__device__ __forceinline__ void fps_st_async_v4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void fps_st_async_b32(uint32_t raddr, uint32_t rbar, uint32_t a) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void fps_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(fps_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

struct __align__(16) ClusterSlot {
    uint32_t bits, tie;
    float x, y;
    float z;
    uint32_t pad[3];
};

// T threads, P points per thread, cluster of C CTAs per cloud (C == 1: plain CTA).
template <int T, int P, int C>
__global__ void __launch_bounds__(T, 1)
fps_onchip_kernel(int b, int n, int m, const float *__restrict__ xyz, int *__restrict__ idx) {
    static_assert(T % 32 == 0 && T >= 64 && T <= 1024, "bad T");
    static_assert(C == 1 || (T * P) % 512 == 0, "cluster chunks must be 512-aligned for the tie order");
    constexpr int NW = T / 32;
    // Per-thread scan order that visits this thread's points in ascending tie order, so a strict
    // '>' keeps the reference's winner: points are k = base + t + i*T; with T < 512 their
    // (k mod 512) cycles with period R = 512/T, so visit residue classes in turn.
    constexpr int R = (T >= 512) ? 1 : 512 / T;

    extern __shared__ float4 s_xyz[];  // T*P local points (winner lookup only)
    __shared__ uint2 s_slot[2][32];
    __shared__ ClusterSlot s_cslot[2][C > 1 ? C : 1];
    __shared__ __align__(8) uint64_t s_cbar[2];  // one per slot buffer; 1 local arrival + 20*C transaction bytes per use

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int cloud, rank;
    if (C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = (int)cluster.block_rank();
        cloud = blockIdx.x / C;
    } else {
        rank = 0;
        cloud = blockIdx.x;
    }
    const int base = rank * (T * P);
    const float *p = xyz + (size_t)cloud * n * 3;
    int *out = idx + (size_t)cloud * m;

    float px[P], py[P], pz[P], md[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = base + t + i * T;
        if (k < n) {
            px[i] = __ldg(p + (size_t)k * 3 + 0);
            py[i] = __ldg(p + (size_t)k * 3 + 1);
            pz[i] = __ldg(p + (size_t)k * 3 + 2);
            md[i] = 1e38f;  // tf_sampling_g.cu:118
        } else {
            px[i] = py[i] = pz[i] = 0.f;
            md[i] = -1.f;  // never wins (reference idle threads hold best=-1, :125)
        }
        s_xyz[t + i * T] = make_float4(px[i], py[i], pz[i], 0.f);
    }
    float x1 = __ldg(p + 0), y1 = __ldg(p + 1), z1 = __ldg(p + 2);  // old = 0 (:114)
    if (t == 0 && rank == 0) out[0] = 0;
    if (C > 1 && t == 0) {
        fps_mbar_init(&s_cbar[0], 1);
        fps_mbar_init(&s_cbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (C > 1) cg::this_cluster().sync();  // barriers initialised cluster-wide before anyone arrives

    for (int j = 1; j < m; ++j) {
        const int buf = j & 1;
        float best = -1.f;
        int bi = -1;
#pragma unroll
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int i = a; i < P; i += R) {
                const float d = d2_ref_gpu(__fsub_rn(px[i], x1), __fsub_rn(py[i], y1), __fsub_rn(pz[i], z1));
                md[i] = fminf(d, md[i]);
                if (md[i] > best) {
                    best = md[i];
                    bi = i;
                }
            }
        }
        const uint32_t bits = bi >= 0 ? __float_as_uint(best) : 0u;
        const uint32_t tie = bi >= 0 ? fps_tie_of((uint32_t)(base + t + bi * T)) : kTieNone;
        uint32_t wb, wt;
        warp_argmax(bits, tie, wb, wt);
        if (lane == 0) s_slot[buf][warp] = make_uint2(wb, wt);
        __syncthreads();
        const uint2 v = lane < NW ? s_slot[buf][lane] : make_uint2(0u, kTieNone);
        uint32_t cb, ct;
        warp_argmax(v.x, v.y, cb, ct);

        if (C == 1) {
            const uint32_t kwin = fps_k_of_tie(ct);
            const float4 w = s_xyz[kwin];
            x1 = w.x; y1 = w.y; z1 = w.z;
            if (t == 0) out[j] = (int)kwin;
        } else {
            if (warp == 0) {
                // A CTA whose chunk lies wholly beyond n has no valid point: ct == kTieNone.
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ct != kTieNone) w = s_xyz[fps_k_of_tie(ct) - base];
                if (lane == 0) fps_mbar_arrive_expect_tx(&s_cbar[buf], 20u * C);  // what my own barrier will receive
                if (lane < C) {
                    // publish my CTA's winner into slot [rank] of CTA `lane`; the stores themselves complete
                    // 20 transaction bytes on that CTA's barrier.
                    const uint32_t rslot = fps_mapa(fps_smem_u32(&s_cslot[buf][rank]), (uint32_t)lane);
                    const uint32_t rbar = fps_mapa(fps_smem_u32(&s_cbar[buf]), (uint32_t)lane);
                    fps_st_async_v4(rslot, rbar, cb, ct, __float_as_uint(w.x), __float_as_uint(w.y));
                    fps_st_async_b32(rslot + 16, rbar, __float_as_uint(w.z));
                }
            }
            // buffer `buf` is used at steps j = 2u+1 (buf 1) / 2u+2 (buf 0): use number u = (j-1)>>1.
            // Re-use is safe: a peer can only be two steps ahead after it received my stores for the step in
            // between, which I issue after all my warps consumed this step's slots.
            fps_mbar_wait_cluster(&s_cbar[buf], (uint32_t)(((j - 1) >> 1) & 1));
            uint32_t sb = 0u, st = kTieNone;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            if (lane < C) {
                const ClusterSlot s = s_cslot[buf][lane];
                sb = s.bits; st = s.tie; sx = s.x; sy = s.y; sz = s.z;
            }
            uint32_t gb, gt;
            warp_argmax(sb, st, gb, gt);
            const int src = __ffs(__ballot_sync(kFull, sb == gb && st == gt)) - 1;
            x1 = __shfl_sync(kFull, sx, src);
            y1 = __shfl_sync(kFull, sy, src);
            z1 = __shfl_sync(kFull, sz, src);
            if (t == 0 && rank == 0) out[j] = (int)fps_k_of_tie(gt);
        }
    }
    // keep peers' shared memory alive until every CTA of the cluster has read its slots
    if (C > 1) cg::this_cluster().sync();
}

// Any n: running min-distance in the caller's scratch (the reference's `temp`, (32,n) floats),
// xyz streamed from global/L2.  Only used beyond the on-chip capacity.
__global__ void __launch_bounds__(512, 1)
fps_global_kernel(int b, int n, int m, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idx) {
    __shared__ uint2 s_slot[2][32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int cloud = blockIdx.x; cloud < b; cloud += gridDim.x) {
        const float *p = xyz + (size_t)cloud * n * 3;
        float *mind = temp + (size_t)blockIdx.x * n;
        int *out = idx + (size_t)cloud * m;
        for (int k = t; k < n; k += 512) mind[k] = 1e38f;
        if (t == 0) out[0] = 0;
        uint32_t last = 0;
        __syncthreads();
        for (int j = 1; j < m; ++j) {
            const int buf = j & 1;
            const float x1 = __ldg(p + (size_t)last * 3), y1 = __ldg(p + (size_t)last * 3 + 1),
                        z1 = __ldg(p + (size_t)last * 3 + 2);
            float best = -1.f;
            int bk = -1;
            for (int k = t; k < n; k += 512) {  // same residue class, k ascending: strict '>' suffices
                const float d = d2_ref_gpu(__fsub_rn(__ldg(p + (size_t)k * 3), x1),
                                           __fsub_rn(__ldg(p + (size_t)k * 3 + 1), y1),
                                           __fsub_rn(__ldg(p + (size_t)k * 3 + 2), z1));
                const float td = mind[k];
                const float d2 = fminf(d, td);
                if (d2 != td) mind[k] = d2;
                if (d2 > best) {
                    best = d2;
                    bk = k;
                }
            }
            const uint32_t bits = bk >= 0 ? __float_as_uint(best) : 0u;
            const uint32_t tie = bk >= 0 ? fps_tie_of((uint32_t)bk) : kTieNone;
            uint32_t wb, wt;
            warp_argmax(bits, tie, wb, wt);
            if (lane == 0) s_slot[buf][warp] = make_uint2(wb, wt);
            __syncthreads();
            const uint2 v = lane < 16 ? s_slot[buf][lane] : make_uint2(0u, kTieNone);
            uint32_t cb, ct;
            warp_argmax(v.x, v.y, cb, ct);
            last = fps_k_of_tie(ct);
            if (t == 0) out[j] = (int)last;
        }
        __syncthreads();
    }
}

template <int T, int P, int C>
static int launch_onchip(int b, int n, int m, const float *inp, int *out, cudaStream_t stream) {
    auto kern = fps_onchip_kernel<T, P, C>;
    const size_t smem = sizeof(float4) * (size_t)T * P;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * C);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (C > 1) {
        if (C > 8) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return (int)e;
        }
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    count_launch();
    e = cudaLaunchKernelEx(&cfg, kern, b, n, m, inp, out);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}


// Largest cluster the 1024-thread kernel can be launched with on THIS device / partition (16 needs the non-portable
// cluster size; a MIG slice or a future part may only give 8).  Queried once; decides the on-chip capacity the
// wrappers see, so "no temp needed" is never promised for a cloud that would have to fall back to the global kernel.
static int fps_max_cluster() {
    static const int cached = [] {
        auto kern = fps_onchip_kernel<1024, 8, 16>;
        const size_t smem = sizeof(float4) * 1024 * 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            (void)cudaGetLastError();
            return 8;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = smem;
        int c = 0;
        if (cudaOccupancyMaxPotentialClusterSize(&c, kern, &cfg) != cudaSuccess) {
            (void)cudaGetLastError();
            return 8;
        }
        return c >= 16 ? 16 : 8;
    }();
    return cached;
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_fps_onchip_capacity(void) { return fps_max_cluster() * 1024 * 8; }

extern "C" int dispu_fps(int b, int n, int m, const float *inp, float *temp, int *out, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0, DISPU_EINVAL);
    DISPU_REQUIRE(m > 0, DISPU_EINVAL);  // attr npoint>0, tf_sampling.cpp:99
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(inp && out, DISPU_ENULL);
    // single CTA per cloud: fewest warps that keep <= 8 points per thread
    if (n <= 128) return launch_onchip<128, 1, 1>(b, n, m, inp, out, stream);
    if (n <= 256) return launch_onchip<128, 2, 1>(b, n, m, inp, out, stream);
    if (n <= 512) return launch_onchip<128, 4, 1>(b, n, m, inp, out, stream);
    if (n <= 1024) return launch_onchip<128, 8, 1>(b, n, m, inp, out, stream);
    if (n <= 2048) return launch_onchip<256, 8, 1>(b, n, m, inp, out, stream);
    if (n <= 4096) return launch_onchip<512, 8, 1>(b, n, m, inp, out, stream);
    if (n <= 8192) return launch_onchip<1024, 8, 1>(b, n, m, inp, out, stream);
    // one cluster per cloud: the fewest CTAs that hold the cloud (<= 8 points per thread), then more
    // CTAs with fewer points each while that still leaves the GPU under-filled (small batches: the
    // per-step scan shrinks with P, the exchange cost barely grows with C).
    const int max_c = fps_max_cluster();
    if (n <= max_c * 8192) {
        int C = 2;
        while (C * 8192 < n) C *= 2;
        int P = ceil_div(n, C * 1024);
        while (C < max_c && P > 1 && (long long)b * C * 2 <= kNumSMs) {
            C *= 2;
            P = ceil_div(n, C * 1024);
        }
        if (const char *force = getenv("DISPU_FPS_FORCE_C")) {  // tuning hook (tools/), not a product knob
            const int fc = atoi(force);
            if ((fc == 2 || fc == 4 || fc == 8 || fc == 16) && fc <= max_c && (long long)fc * 8192 >= n) {
                C = fc;
                P = ceil_div(n, C * 1024);
            }
        }
        P = P <= 1 ? 1 : P <= 2 ? 2 : P <= 4 ? 4 : 8;
        int rc = DISPU_EUNSUPPORTED;
#define DISPU_FPS_CASE(CC, PP) \
    if (C == CC && P == PP) rc = launch_onchip<1024, PP, CC>(b, n, m, inp, out, stream);
        DISPU_FPS_CASE(2, 1) DISPU_FPS_CASE(2, 2) DISPU_FPS_CASE(2, 4) DISPU_FPS_CASE(2, 8)
        DISPU_FPS_CASE(4, 1) DISPU_FPS_CASE(4, 2) DISPU_FPS_CASE(4, 4) DISPU_FPS_CASE(4, 8)
        DISPU_FPS_CASE(8, 1) DISPU_FPS_CASE(8, 2) DISPU_FPS_CASE(8, 4) DISPU_FPS_CASE(8, 8)
        DISPU_FPS_CASE(16, 1) DISPU_FPS_CASE(16, 2) DISPU_FPS_CASE(16, 4) DISPU_FPS_CASE(16, 8)
#undef DISPU_FPS_CASE
        return rc;
    }
    DISPU_REQUIRE(temp != nullptr, DISPU_EWORKSPACE);
    count_launch();
    fps_global_kernel<<<b < 32 ? b : 32, 512, 0, stream>>>(b, n, m, inp, temp, out);
    return finish_launch();
}

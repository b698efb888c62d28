// Non-local attention of PointNonLocalCell (Common/ops.py:326-339) as ONE tcgen05 kernel, 3xTF32 on both products:
//     O[i,:] = sum_j softmax_j(q_i . k_j) v_j          q pre-scaled by 1/sqrt(d), d = 64, m keys per cloud
// so the (b,m,m) score tensor (134 MB at c2, 2.1 GB at the second stage of c4) is never written.
//
// One CTA per (cloud, 128-query tile), persistent.  Tensor memory (512 columns):
//     [  0,128)  Q operand, TF32 hi | lo (64 + 64 columns), written once per tile by the softmax warps
//     [128,256)  S = Q . K_j^T for one block of 128 keys
//     [256,384)  P ring: 2 slots x (32 hi + 32 lo columns) = exp(s - max) for 32 keys, the A operand of O += P . V
//     [384,448)  O accumulator (64 columns)
// Shared memory: two 64 KB buffers for the packed K blocks, one 64 KB buffer for the packed V block (both produced by
// dispu_dense_pack_strided per (cloud, key block): K as N=128 keys x K=64 channels, V as N=64 channels x K=128 keys).
// Two passes over the key blocks (exact softmax, no rescaling of O): pass 1 computes S with ONE TF32 product per K-step
// and only the row maximum is taken from it (any shift close to the maximum is exact for the softmax, as long as
// numerator and denominator share it), pass 2 recomputes S in 3xTF32, turns it into P chunk by chunk and accumulates O.
// Validated on B200 against a float64 softmax (tests/test_gpu_parity.py::test_attention_fused): 0.106 ms at the c2 shape
// (32 clouds x 1024 points) against 0.209 ms for the two batched products with the folded softmax it replaces.
// Roles: warps 0-3 softmax / epilogue (thread = query row = TMEM lane), warp 4 lane 0 issues every MMA, warp 5 lane 0
// issues the TMA bulk copies.  All hand-offs are mbarriers; MMA completion is signalled with tcgen05.commit.
// The barrier protocol (not the arithmetic) is model-checked on the host by tools/attention_protocol_sim.py: random
// interleavings of the six roles with the hardware's phase/parity semantics, no deadlock, no read-before-produce or
// overwrite-before-consume on K/V buffers, S, the P ring, Q and O (o_empty is implied by q_full and kept as a guard).
#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kAtThreads = 6 * 32;
constexpr int kAtD = 64;               // channels of q / k / v
constexpr int kAtBK = 128;             // keys per block
constexpr uint32_t kAtBlockBytes = 65536;  // packed K block == packed V block
constexpr uint32_t kAtColQ = 0, kAtColS = 128, kAtColP = 256, kAtColO = 384;

__device__ __forceinline__ void at_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(kAtThreads, 1)
attention_tf32x3_kernel(int b, int m, const float *__restrict__ q, long long ldq, const uint8_t *__restrict__ k_packed,
                        const uint8_t *__restrict__ v_packed, float *__restrict__ out, long long ldo) {
    extern __shared__ __align__(1024) uint8_t at_smem[];  // K buffer 0 | K buffer 1 | V buffer
    __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty, s_full, s_empty, p_full[2], p_empty[2],
        o_full, o_empty;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = m / 128, nb = m / kAtBK;
    const int n_tiles = b * qt;
    const uint32_t smem_base = smem_u32(at_smem);

    if (tid == 0) {
        mbar_init(&q_full, 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(&v_full, 1);
        mbar_init(&v_empty, 1);
        mbar_init(&s_full, 1);
        mbar_init(&s_empty, 4);
        mbar_init(&o_full, 1);
        mbar_init(&o_empty, 4);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&s_tmem, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem;

    if (warp == 5) {
        // ================= TMA: K block twice per (tile, block) (once per pass), V block in pass 2 =================
        if (lane == 0) {
            uint32_t kc = 0, vc = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int cloud = tile / qt;
                for (int pass = 0; pass < 2; ++pass) {
                    for (int j = 0; j < nb; ++j) {
                        const size_t blk = ((size_t)cloud * nb + j) * kAtBlockBytes;
                        const int buf = kc & 1;
                        mbar_wait(&k_empty[buf], ((kc >> 1) & 1) ^ 1);
                        at_arrive_expect_tx(&k_full[buf], kAtBlockBytes);
                        at_bulk_g2s(smem_base + (uint32_t)buf * kAtBlockBytes, k_packed + blk, 32768, &k_full[buf]);
                        at_bulk_g2s(smem_base + (uint32_t)buf * kAtBlockBytes + 32768, k_packed + blk + 32768, 32768, &k_full[buf]);
                        ++kc;
                        if (pass == 1) {
                            mbar_wait(&v_empty, (vc & 1) ^ 1);
                            at_arrive_expect_tx(&v_full, kAtBlockBytes);
                            at_bulk_g2s(smem_base + 2 * kAtBlockBytes, v_packed + blk, 32768, &v_full);
                            at_bulk_g2s(smem_base + 2 * kAtBlockBytes + 32768, v_packed + blk + 32768, 32768, &v_full);
                            ++vc;
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_tf32(128, kAtBK), idesc_o = make_idesc_tf32(128, kAtD);
            const uint32_t lbo_k = (kAtBK / 8) * 128u, lbo_v = (kAtD / 8) * 128u;  // 2048, 1024
            const uint32_t plane_k = kAtBK * 32 * 4u, plane_v = kAtD * 32 * 4u;    // 16 KB, 8 KB per 32-K chunk plane
            uint32_t kc = 0, vc = 0, sc = 0, pc = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                mbar_wait(&q_full, tcount & 1);
                tc_fence_after_sync();
                for (int pass = 0; pass < 2; ++pass) {
                    for (int j = 0; j < nb; ++j) {
                        const int buf = kc & 1;
                        mbar_wait(&k_full[buf], (kc >> 1) & 1);
                        mbar_wait(&s_empty, (sc & 1) ^ 1);
                        tc_fence_after_sync();
                        // S = Q . K_j^T : 64 channels = 2 chunks of 32 = 8 K-steps of 8
                        const uint32_t kb = smem_base + (uint32_t)buf * kAtBlockBytes;
#pragma unroll
                        for (int ks = 0; ks < kAtD / 8; ++ks) {
                            const uint32_t chunk = kb + (uint32_t)(ks >> 2) * 2u * plane_k + (uint32_t)(ks & 3) * 2u * lbo_k;
                            const uint64_t dbh = make_smem_desc(chunk, lbo_k, 128u);
                            const uint64_t dbl = make_smem_desc(chunk + plane_k, lbo_k, 128u);
                            const uint32_t a_hi = tmem_base + kAtColQ + (uint32_t)ks * 8u, a_lo = a_hi + 64u;
                            if (pass == 0) {
                                // pass 1 only needs a row shift close to the maximum (softmax is shift-invariant as long as
                                // numerator and denominator use the same shift): one TF32 product instead of three
                                mma_tf32_ts(tmem_base + kAtColS, a_hi, dbh, idesc_s, ks != 0);
                            } else {
                                mma_tf32_ts(tmem_base + kAtColS, a_lo, dbh, idesc_s, ks != 0);
                                mma_tf32_ts(tmem_base + kAtColS, a_hi, dbl, idesc_s, 1);
                                mma_tf32_ts(tmem_base + kAtColS, a_hi, dbh, idesc_s, 1);
                            }
                        }
                        mma_commit(&k_empty[buf]);
                        mma_commit(&s_full);
                        ++kc;
                        ++sc;
                        if (pass == 1) {
                            mbar_wait(&v_full, vc & 1);
                            if (j == 0) mbar_wait(&o_empty, (tcount & 1) ^ 1);  // previous tile's O has been read out
                            const uint32_t vb = smem_base + 2 * kAtBlockBytes;
                            for (int cc = 0; cc < kAtBK / 32; ++cc) {
                                const int slot = pc & 1;
                                mbar_wait(&p_full[slot], (pc >> 1) & 1);
                                tc_fence_after_sync();
                                const uint32_t pa = tmem_base + kAtColP + (uint32_t)slot * 64u;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t chunk = vb + (uint32_t)cc * 2u * plane_v + (uint32_t)ks * 2u * lbo_v;
                                    const uint64_t dbh = make_smem_desc(chunk, lbo_v, 128u);
                                    const uint64_t dbl = make_smem_desc(chunk + plane_v, lbo_v, 128u);
                                    const uint32_t a_hi = pa + (uint32_t)ks * 8u, a_lo = a_hi + 32u;
                                    mma_tf32_ts(tmem_base + kAtColO, a_lo, dbh, idesc_o, (j | cc | ks) != 0);
                                    mma_tf32_ts(tmem_base + kAtColO, a_hi, dbl, idesc_o, 1);
                                    mma_tf32_ts(tmem_base + kAtColO, a_hi, dbh, idesc_o, 1);
                                }
                                mma_commit(&p_empty[slot]);
                                ++pc;
                            }
                            mma_commit(&v_empty);
                            ++vc;
                        }
                    }
                }
                mma_commit(&o_full);
            }
        }
    } else {
        // ================= softmax / epilogue: thread = query row = TMEM lane =================
        const int qd = warp;  // lane quarter == warp % 4
        const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
        const float kLog2e = 1.4426950408889634f;
        uint32_t sc = 0, pc = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int cloud = tile / qt;
            const long long grow = (long long)cloud * m + (long long)(tile - cloud * qt) * 128 + qd * 32 + lane;
            // ---- Q operand (the previous tile's S products are complete: their last s_full was observed below)
            {
                const float4 *q4 = reinterpret_cast<const float4 *>(q + grow * ldq);
#pragma unroll
                for (int c8 = 0; c8 < kAtD; c8 += 8) {
                    const float4 x0 = __ldg(q4 + c8 / 4), x1 = __ldg(q4 + c8 / 4 + 1);
                    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    uint32_t hi8[8], lo8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float hh = tf32_round_finite(xs[u]);
                        hi8[u] = __float_as_uint(hh);
                        lo8[u] = __float_as_uint(tf32_round_finite(__fsub_rn(xs[u], hh)));
                    }
                    tmem_st8(lane_base + kAtColQ + (uint32_t)c8, hi8);
                    tmem_st8(lane_base + kAtColQ + 64u + (uint32_t)c8, lo8);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) at_arrive(&q_full);
            }
            // ---- pass 1: row maximum
            float mx = -__int_as_float(0x7f800000);
            for (int j = 0; j < nb; ++j, ++sc) {
                mbar_wait(&s_full, sc & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < kAtBK; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + kAtColS + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 32; ++u) mx = fmaxf(mx, __uint_as_float(v[u]));
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) at_arrive(&s_empty);
            }
            // ---- pass 2: P = exp(S - max) chunk by chunk into the ring, row sum on the side
            const float ml2 = __fmul_rn(mx, kLog2e);
            float sum = 0.f;
            for (int j = 0; j < nb; ++j, ++sc) {
                mbar_wait(&s_full, sc & 1);
                tc_fence_after_sync();
                for (int cc = 0; cc < kAtBK / 32; ++cc, ++pc) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + kAtColS + (uint32_t)cc * 32u, v);
                    tmem_ld_wait();
                    const int slot = pc & 1;
                    mbar_wait(&p_empty[slot], ((pc >> 1) & 1) ^ 1);
                    tc_fence_after_sync();
#pragma unroll
                    for (int c8 = 0; c8 < 32; c8 += 8) {
                        uint32_t hi8[8], lo8[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float e2;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(__fmaf_rn(__uint_as_float(v[c8 + u]), kLog2e, -ml2)));
                            sum = __fadd_rn(sum, e2);
                            const float hh = tf32_round_finite(e2);
                            hi8[u] = __float_as_uint(hh);
                            lo8[u] = __float_as_uint(tf32_round_finite(__fsub_rn(e2, hh)));
                        }
                        tmem_st8(lane_base + kAtColP + (uint32_t)slot * 64u + (uint32_t)c8, hi8);
                        tmem_st8(lane_base + kAtColP + (uint32_t)slot * 64u + 32u + (uint32_t)c8, lo8);
                    }
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) at_arrive(&p_full[slot]);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) at_arrive(&s_empty);
            }
            // ---- output row = O / sum
            mbar_wait(&o_full, tcount & 1);
            tc_fence_after_sync();
            const float inv = __fdividef(1.f, sum);
            float4 *o4 = reinterpret_cast<float4 *>(out + grow * ldo);
#pragma unroll
            for (int c0 = 0; c0 < kAtD; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(lane_base + kAtColO + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 32; u += 4)
                    o4[(c0 + u) / 4] = make_float4(__fmul_rn(__uint_as_float(v[u]), inv), __fmul_rn(__uint_as_float(v[u + 1]), inv),
                                                   __fmul_rn(__uint_as_float(v[u + 2]), inv), __fmul_rn(__uint_as_float(v[u + 3]), inv));
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) at_arrive(&o_empty);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace dispu

using namespace dispu;

// q (b,m,64) pre-scaled by 1/sqrt(64), row stride ldq;
// k_packed / v_packed: b*(m/128) operands of 64 KB each from dispu_dense_pack_strided (K: K=64,N=128 of key block j;
// V: K=128,N=64 of key block j); out (b,m,64) row stride ldo.  m % 128 == 0, 16-byte aligned q / out rows.
extern "C" int dispu_attention_tf32x3(int b, int m, const float *q, long long ldq, const void *k_packed,
                                                   const void *v_packed, float *out, long long ldo, dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && m > 0 && ldq >= kAtD && ldo >= kAtD, DISPU_EINVAL);
    DISPU_REQUIRE(m % 128 == 0 && ldq % 4 == 0 && ldo % 4 == 0, DISPU_EUNSUPPORTED);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(q && k_packed && v_packed && out, DISPU_ENULL);
    DISPU_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(k_packed) |
                    reinterpret_cast<uintptr_t>(v_packed)) & 15u) == 0, DISPU_EUNSUPPORTED);
    const int n_tiles = b * (m / 128);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = 3 * (size_t)kAtBlockBytes;
    cudaError_t e = cudaFuncSetAttribute(attention_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    attention_tf32x3_kernel<<<n_tiles < sms ? n_tiles : sms, kAtThreads, smem, (cudaStream_t)stream>>>(
        b, m, q, ldq, reinterpret_cast<const uint8_t *>(k_packed), reinterpret_cast<const uint8_t *>(v_packed), out, ldo);
    return finish_launch();
}

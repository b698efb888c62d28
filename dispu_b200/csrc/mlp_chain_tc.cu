// Chained dense layers on the tensor cores (tcgen05, 3xTF32, FP32-level accuracy): a whole per-point MLP
//     H1 = act(X . W1 + b1)   (K0 -> N1 <= 256)
//     H2 = act(H1 . W2 + b2)  (N1 -> N2 <= 64)
//     H3 = act(H2 . W3 + b3)  (N2 -> N3 <= 256)
// in ONE launch with the intermediate activations never leaving tensor memory.  This is the coordinate regressor of
// the live generator (Common/ops.py:1089-1110: conv1d 256 -> 64 -> 3, each a tf_util.conv2d 1x1 with bias + ReLU,
// Common/tf_util.py:120-185) behind its 128/256 -> 256 input layer.  As three launches of the dense-layer kernel the
// regressor paid, per launch, the pipeline fill and drain of a persistent kernel that owns fewer than two tiles per
// SM, wrote a (b,m,256) and a (b,m,64) tensor and read them back: 78 us (coarse) and 86 us (fine) at c2 for 5 GFLOP.
//
// One CTA per SM, 128-row tiles, 18 warps:
//   warps 0-7   producers (layer 1 only): raw global loads of the tile's X rows two 32-column chunks ahead in registers,
//               split into TF32 hi/lo and written straight into the tensor-memory A ring with tcgen05.st (as in dense_tc.cu);
//   warps 8-15  converters / epilogue: when a layer's accumulator is complete they read it back 16 columns at a time
//               (tcgen05.ld), add the bias, apply the activation, split into hi/lo and write the NEXT layer's A operand
//               chunk into the same ring ("epilogue as producer"); after the last layer they write the result rows
//               (optionally as  base + sigmoid(x) - 0.5  for the offset regressor, Common/ops.py:1106-1108, and to several
//               destinations at once: the coordinates are also columns of two wider tensors the next stage reads);
//   warp  16    one thread issues every tcgen05.mma (TS form: A from tensor memory, B from shared memory) in program order
//               layer 1 chunks, layer 2 chunks, layer 3 chunks, releasing ring slots with tcgen05.commit;
//   warp  17    one thread streams the pre-packed [hi | lo] weight chunks of all three layers through a 3-slot shared
//               memory ring with TMA bulk copies (cp.async.bulk + mbarrier transaction bytes), in consumption order.
// A chunk = 32 K-columns of the current layer's A operand (TMEM ring slot: hi 32 + lo 32 columns) together with the
// matching weight chunk (shared-memory ring slot); slot i % 3 of both rings belongs to the i-th chunk of the CTA's
// global chunk sequence, so ONE full barrier (8 writer warps + the TMA's transaction bytes) and ONE empty barrier
// (tcgen05.commit) per slot serve both rings.
// Tensor memory: [0,256) accumulator of layer 1 (and, once it has been consumed, of layer 3), [256,320) accumulator of
// layer 2, [320,512) the A ring.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dispu {

using namespace tc;

constexpr int kMcKC = 32;
constexpr int kMcSlots = 3;
constexpr int kMcProducerWarps = 8;
constexpr int kMcConvWarps = 8;
constexpr int kMcThreads = (kMcProducerWarps + kMcConvWarps + 2) * 32;
constexpr uint32_t kMcSlotBytes = 65536;  // weight chunk of a 256-column layer: 2 planes x 256 rows x 32 k x 4 B
constexpr uint32_t kMcColAcc1 = 0, kMcColAcc2 = 256, kMcColRing = 320;

__host__ __device__ __forceinline__ int mc_npad(int N) { return (N + 15) & ~15; }
__host__ __device__ __forceinline__ int mc_chunks(int K) { return (K + kMcKC - 1) / kMcKC; }

struct ChainParams {
    long long M;
    int K0;
    const float *X;
    long long ldx;
    int L;  // 2 or 3 layers
    int N[3];
    const uint8_t *W[3];
    const float *bias[3];
    int relu[3];
    // result of the last layer: the first `head` (<= 4) columns — packed in a 16-column block of their own — go to up to
    // three destinations, optionally as base + sigmoid(x) - 0.5; the remaining columns (from column 16 on when head > 0)
    // go to Yt
    int head;
    float *Yh[3];
    long long ldh[3];
    const float *head_base;
    long long ldb;
    float *Yt;
    long long ldt;
    int n_tiles;
};

__device__ __forceinline__ void mc_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mc_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mc_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float4 mc_load4(const float *__restrict__ row, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) {
        v = __ldg(reinterpret_cast<const float4 *>(row + k));
    } else {
        if (k < K) v.x = __ldg(row + k);
        if (k + 1 < K) v.y = __ldg(row + k + 1);
        if (k + 2 < K) v.z = __ldg(row + k + 2);
        if (k + 3 < K) v.w = __ldg(row + k + 3);
    }
    return v;
}
// 8 values -> TF32 hi / lo planes of the A operand (lo rounded, not truncated: halves the error of the lo*w_hi term)
__device__ __forceinline__ void mc_split8(const float (&x)[8], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float h = tf32_round_finite(x[u]);
        hi[u] = __float_as_uint(h);
        lo[u] = __float_as_uint(tf32_round_finite(__fsub_rn(x[u], h)));
    }
}

__global__ void __launch_bounds__(kMcThreads, 1)
mlp_chain_kernel(const __grid_constant__ ChainParams prm) {
    extern __shared__ __align__(1024) uint8_t mc_smem[];  // 3 weight slots of 64 KB
    __shared__ __align__(8) uint64_t s_full[kMcSlots], s_empty[kMcSlots], s_acc[3], s_drained;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_bias[3][256];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long M = prm.M;
    const int L = prm.L;
    const int n_tiles = prm.n_tiles;
    const int np1 = mc_npad(prm.N[0]), np2 = mc_npad(prm.N[1]), np3 = L == 3 ? mc_npad(prm.N[2]) : 0;
    const int nch1 = mc_chunks(prm.K0), nch2 = mc_chunks(prm.N[0]), nch3 = L == 3 ? mc_chunks(prm.N[1]) : 0;
    const int cpt = nch1 + nch2 + nch3;  // chunks per tile
    const uint32_t smem_base = smem_u32(mc_smem);
    constexpr int kMmaWarp = kMcProducerWarps + kMcConvWarps, kTmaWarp = kMmaWarp + 1;

    if (tid == 0) {
        for (int s = 0; s < kMcSlots; ++s) {
            mbar_init(&s_full[s], kMcProducerWarps + 1);  // 8 writer warps (producers or converters) + the TMA's expect_tx arrival
            mbar_init(&s_empty[s], 1);                    // tcgen05.commit
        }
        for (int a = 0; a < 3; ++a) mbar_init(&s_acc[a], 1);  // accumulator of layer a complete (tcgen05.commit)
        mbar_init(&s_drained, kMcConvWarps);                  // last accumulator read out: the next tile may overwrite [0,256)
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 512);
    for (int e = tid; e < 3 * 256; e += kMcThreads) {
        const int l = e >> 8, col = e & 255;
        s_bias[l][col] = (l < L && prm.bias[l] && col < prm.N[l]) ? __ldg(prm.bias[l] + col) : 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem;
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp < kMcProducerWarps) {
        // ================= producers: layer-1 A operand =================
        // thread (r, h): row r = tid & 127 (TMEM lane r; this warp's lane quarter = warp % 4), K-columns 16h..16h+15 of a chunk
        constexpr int kDepth = 2;
        const int r = tid & 127, h = tid >> 7;
        const uint32_t tm_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kMcColRing + (uint32_t)h * 16u;
        const bool vec = ((reinterpret_cast<uintptr_t>(prm.X) & 15u) == 0) && (prm.ldx % 4 == 0);
        const int K0 = prm.K0;
        const int total = my_tiles * nch1;
        int ld_t = 0, ld_c = 0;  // load cursor: (tile ordinal of this CTA, chunk)
        auto row_ptr = [&](int t_ord, bool &valid) -> const float * {
            const long long row = ((long long)blockIdx.x + (long long)t_ord * gridDim.x) * 128 + r;
            valid = t_ord < my_tiles && row < M;
            return prm.X + (valid ? row : 0) * prm.ldx;
        };
        bool rvalid = false;
        const float *xr = row_ptr(0, rvalid);
        float4 xb[kDepth][4];
        auto load_chunk = [&](float4 (&xv)[4]) {
            const int k0 = ld_c * kMcKC + h * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[j] = rvalid ? mc_load4(xr, k0 + j * 4, K0, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (++ld_c == nch1) {
                ld_c = 0;
                ++ld_t;
                xr = row_ptr(ld_t, rvalid);
                if (rvalid) {  // pull the head of the next tile's row into L2 while this tile is converted
                    for (int cc = h; cc < nch1 && cc < 8; cc += 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + cc * kMcKC));
                }
            }
        };
#pragma unroll
        for (int d = 0; d < kDepth; ++d)
            if (d < total) load_chunk(xb[d]);
        int t_ord = 0, c = 0;
        for (int base = 0; base < total; base += kDepth) {
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                if (base + d < total) {
                    // A ring slot may be filled by a DIFFERENT thread group next (the converters): parity waits tell only
                    // "one phase ahead" from "in step", so the producers re-synchronise at every tile boundary — they start
                    // tile t only when every chunk of tile t-1 has been consumed (its last accumulator is complete)
                    if (c == 0 && t_ord > 0) mbar_wait(&s_acc[L - 1], (uint32_t)((t_ord - 1) & 1));
                    const uint32_t it = (uint32_t)t_ord * (uint32_t)cpt + (uint32_t)c;  // global chunk sequence number
                    const int s = it % kMcSlots;
                    const uint32_t ph = (it / kMcSlots) & 1;
                    mbar_wait(&s_empty[s], ph ^ 1);
                    tc_fence_after_sync();
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        const float4 v0 = xb[d][2 * j2], v1 = xb[d][2 * j2 + 1];
                        const float xs[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        uint32_t hi8[8], lo8[8];
                        mc_split8(xs, hi8, lo8);
                        tmem_st8(tm_row + (uint32_t)s * 64u + (uint32_t)j2 * 8u, hi8);
                        tmem_st8(tm_row + (uint32_t)s * 64u + 32u + (uint32_t)j2 * 8u, lo8);
                    }
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mc_arrive(&s_full[s]);
                    if (base + d + kDepth < total) load_chunk(xb[d]);
                    if (++c == nch1) { c = 0; ++t_ord; }
                }
            }
        }
    } else if (warp == kTmaWarp) {
        // ================= weight stream: every chunk of every layer, in consumption order =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                for (int l = 0; l < L; ++l) {
                    const int npl = l == 0 ? np1 : l == 1 ? np2 : np3;
                    const int nch = l == 0 ? nch1 : l == 1 ? nch2 : nch3;
                    const uint32_t bytes = 2u * (uint32_t)npl * kMcKC * 4u;
                    for (int c = 0; c < nch; ++c, ++it) {
                        const int s = it % kMcSlots;
                        mbar_wait(&s_empty[s], ((it / kMcSlots) & 1) ^ 1);
                        mc_arrive_expect_tx(&s_full[s], bytes);
                        mc_bulk_g2s(smem_base + (uint32_t)s * kMcSlotBytes, prm.W[l] + (size_t)c * bytes, bytes, &s_full[s]);
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                if (t > 0) {  // the previous tile's last accumulator (aliased with this tile's first) has been read out
                    mbar_wait(&s_drained, (uint32_t)((t - 1) & 1));
                    tc_fence_after_sync();
                }
                for (int l = 0; l < L; ++l) {
                    const int npl = l == 0 ? np1 : l == 1 ? np2 : np3;
                    const int nch = l == 0 ? nch1 : l == 1 ? nch2 : nch3;
                    const uint32_t d = tmem_base + (l == 1 ? kMcColAcc2 : kMcColAcc1);
                    const uint32_t idesc = make_idesc_tf32(128, npl);
                    const uint32_t lboB = (uint32_t)(npl / 8) * 128u;
                    const uint32_t b_plane = (uint32_t)npl * kMcKC * 4u;
                    for (int c = 0; c < nch; ++c, ++it) {
                        const int s = it % kMcSlots;
                        mbar_wait(&s_full[s], (it / kMcSlots) & 1);
                        tc_fence_after_sync();
                        const uint32_t ta = tmem_base + kMcColRing + (uint32_t)s * 64u;
                        const uint32_t bh = smem_base + (uint32_t)s * kMcSlotBytes, bl = bh + b_plane;
#pragma unroll
                        for (int ks = 0; ks < kMcKC / 8; ++ks) {
                            const uint64_t dbh = make_smem_desc(bh + ks * 2 * lboB, lboB, 128u);
                            const uint64_t dbl = make_smem_desc(bl + ks * 2 * lboB, lboB, 128u);
                            mma_tf32_ts(d, ta + 32u + ks * 8u, dbh, idesc, (c | ks) != 0);  // a_lo * b_hi
                            mma_tf32_ts(d, ta + ks * 8u, dbl, idesc, 1);                    // a_hi * b_lo
                            mma_tf32_ts(d, ta + ks * 8u, dbh, idesc, 1);                    // a_hi * b_hi
                        }
                        mma_commit(&s_empty[s]);
                    }
                    mma_commit(&s_acc[l]);
                }
            }
        }
    } else {
        // ================= converters / epilogue (8 warps: two per TMEM lane quarter) =================
        const int ew = warp - kMcProducerWarps;
        const int q = ew & 3;      // == warp % 4: the lane quarter this warp may access
        const int half = ew >> 2;  // which 16 of a chunk's 32 columns (conversion) / which 16-column blocks (epilogue)
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t tph = (uint32_t)(t & 1);
            const long long row = ((long long)blockIdx.x + (long long)t * gridDim.x) * 128 + q * 32 + lane;
            uint32_t it = (uint32_t)t * (uint32_t)cpt + (uint32_t)nch1;
            // ---- accumulator of layer l -> A operand chunks of layer l + 1
            for (int l = 0; l + 1 < L; ++l) {
                const int Nl = prm.N[l];
                const int nch = l == 0 ? nch2 : nch3;
                const uint32_t acc = lane_base + (l == 0 ? kMcColAcc1 : kMcColAcc2);
                const float *sb = s_bias[l];
                const int relu = prm.relu[l];
                mbar_wait(&s_acc[l], tph);
                tc_fence_after_sync();
                for (int c = 0; c < nch; ++c, ++it) {
                    const int s = it % kMcSlots;
                    const int col0 = c * kMcKC + half * 16;
                    uint32_t v[32];
                    tmem_ld16(acc + (uint32_t)col0, v);  // (columns past the layer's padded width are never multiplied: zeroed below)
                    tmem_ld_wait();
                    mbar_wait(&s_empty[s], ((it / kMcSlots) & 1) ^ 1);
                    tc_fence_after_sync();
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        float xs[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int col = col0 + j2 * 8 + u;
                            float x = __fadd_rn(__uint_as_float(v[j2 * 8 + u]), sb[col & 255]);
                            x = relu ? fmaxf(x, 0.f) : x;
                            xs[u] = col < Nl ? x : 0.f;
                        }
                        uint32_t hi8[8], lo8[8];
                        mc_split8(xs, hi8, lo8);
                        const uint32_t dst = lane_base + kMcColRing + (uint32_t)s * 64u + (uint32_t)half * 16u + (uint32_t)j2 * 8u;
                        tmem_st8(dst, hi8);
                        tmem_st8(dst + 32u, lo8);
                    }
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mc_arrive(&s_full[s]);
                }
            }
            // ---- result of the last layer
            {
                const int l = L - 1;
                const int Nl = prm.N[l], npl = l == 1 ? np2 : np3;
                const uint32_t acc = lane_base + (l == 1 ? kMcColAcc2 : kMcColAcc1);
                const float *sb = s_bias[l];
                const int relu = prm.relu[l];
                mbar_wait(&s_acc[l], tph);
                tc_fence_after_sync();
                for (int c0 = half * 16; c0 < npl; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld16(acc + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (row >= M) {
                        // (rows past M: nothing to store; stay in the loop — tcgen05.ld is warp-collective)
                    } else if (prm.head > 0 && c0 == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j < prm.head) {
                                float x = __fadd_rn(__uint_as_float(v[j]), sb[j]);
                                x = relu ? fmaxf(x, 0.f) : x;
                                if (prm.head_base)  // offset regressor: base + sigmoid(x) - 0.5 (Common/ops.py:1106-1108, generator.py:78-81)
                                    x = __fadd_rn(__ldg(prm.head_base + row * prm.ldb + j), __fsub_rn(__fdividef(1.f, 1.f + __expf(-x)), 0.5f));
#pragma unroll
                                for (int dsti = 0; dsti < 3; ++dsti)
                                    if (prm.Yh[dsti]) prm.Yh[dsti][row * prm.ldh[dsti] + j] = x;
                            }
                        }
                    } else if (prm.Yt) {
                        const int tcol = prm.head > 0 ? c0 - 16 : c0;  // column in the tail tensor
                        float *yr = prm.Yt + row * prm.ldt + tcol;
                        const bool v4 = ((reinterpret_cast<uintptr_t>(yr) & 15u) == 0) && c0 + 16 <= Nl;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float o[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                o[u] = __fadd_rn(__uint_as_float(v[j + u]), sb[c0 + j + u]);
                                o[u] = relu ? fmaxf(o[u], 0.f) : o[u];
                            }
                            if (v4) {
                                *reinterpret_cast<float4 *>(yr + j) = make_float4(o[0], o[1], o[2], o[3]);
                            } else {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    if (c0 + j + u < Nl) yr[j + u] = o[u];
                            }
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mc_arrive(&s_drained);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_mlp_chain_args_size(void) { return sizeof(dispu_mlp_chain_args); }

extern "C" int dispu_mlp_chain_tf32x3(const dispu_mlp_chain_args *a, dispu_stream_t stream) {
    DISPU_REQUIRE(a, DISPU_ENULL);
    DISPU_REQUIRE(a->M >= 0 && a->K0 > 0 && (a->n_layers == 2 || a->n_layers == 3), DISPU_EINVAL);
    const int L = a->n_layers;
    for (int l = 0; l < L; ++l) DISPU_REQUIRE(a->N[l] > 0, DISPU_EINVAL);
    DISPU_REQUIRE(a->N[0] <= 256 && a->N[1] <= 64 && (L == 2 || a->N[2] <= 256), DISPU_EUNSUPPORTED);
    DISPU_REQUIRE(a->head >= 0 && a->head <= 4, DISPU_EINVAL);
    if (a->M == 0) return DISPU_OK;
    DISPU_REQUIRE(a->X, DISPU_ENULL);
    for (int l = 0; l < L; ++l) DISPU_REQUIRE(a->W[l], DISPU_ENULL);
    const int NL = a->N[L - 1];
    const int tail = a->head > 0 ? NL - 16 : NL;
    DISPU_REQUIRE(a->head == 0 || NL >= 16, DISPU_EINVAL);
    DISPU_REQUIRE(a->head == 0 || a->Yh[0] || a->Yh[1] || a->Yh[2], DISPU_ENULL);
    DISPU_REQUIRE(tail <= 0 || a->Yt || a->head > 0, DISPU_ENULL);
    const long long ldx = a->ldx > 0 ? a->ldx : a->K0;
    DISPU_REQUIRE(ldx >= a->K0, DISPU_EINVAL);
    const long long n_tiles_ll = ceil_div64(a->M, 128);
    DISPU_REQUIRE(n_tiles_ll <= 0x7fffffffLL, DISPU_EUNSUPPORTED);
    ChainParams p = {};
    p.M = a->M; p.K0 = a->K0; p.X = a->X; p.ldx = ldx; p.L = L;
    for (int l = 0; l < 3; ++l) {
        p.N[l] = l < L ? a->N[l] : 0;
        p.W[l] = l < L ? reinterpret_cast<const uint8_t *>(a->W[l]) : nullptr;
        p.bias[l] = l < L ? a->bias[l] : nullptr;
        p.relu[l] = l < L ? a->relu[l] : 0;
    }
    p.head = a->head;
    for (int d = 0; d < 3; ++d) {
        p.Yh[d] = a->Yh[d];
        p.ldh[d] = a->ldh[d] > 0 ? a->ldh[d] : (a->head > 0 ? a->head : 1);
    }
    p.head_base = a->head_base; p.ldb = a->ldb > 0 ? a->ldb : (a->head > 0 ? a->head : 1);
    p.Yt = a->Yt; p.ldt = a->ldt > 0 ? a->ldt : (tail > 0 ? tail : 1);
    DISPU_REQUIRE(!a->Yt || p.ldt >= tail, DISPU_EINVAL);
    p.n_tiles = (int)n_tiles_ll;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    const size_t smem = (size_t)kMcSlots * kMcSlotBytes;
    const cudaError_t e = cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    mlp_chain_kernel<<<grid, kMcThreads, smem, (cudaStream_t)stream>>>(p);
    return finish_launch();
}

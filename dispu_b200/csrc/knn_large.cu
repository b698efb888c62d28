// k-nearest neighbours for LARGE k (64 < k <= 1024): the patch extraction of the inference driver
// (DisPU/model.py:331 -> Common/pc_util.py:83-91 extract_knn_patch: sklearn NearestNeighbors with k = 256 on the host)
// kept on the GPU.  One CTA per query:
//   1. squared distances (host-loop contraction, like dispu_knn) to all n points -> order-preserving 32-bit keys
//      in a global scratch row;
//   2. 4-pass radix select (8 bits per pass, shared-memory histograms) for the k-th smallest key;
//   3. ordered compaction: every key below the threshold, then threshold ties in ascending index until k are taken
//      (ties to the lower index, the order of a stable sort);
//   4. bitonic sort of the k (key, index) pairs in shared memory -> ascending distance, ties by index.
#include "common.cuh"

namespace dispu {

constexpr int kKlThreads = 256;
constexpr int kKlMaxK = 1024;

__global__ void __launch_bounds__(kKlThreads)
knn_large_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
                 float *__restrict__ dist, int *__restrict__ idx, uint32_t *__restrict__ scratch) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_sel[kKlMaxK];  // (key << 32) | index
    __shared__ unsigned s_warp[kKlThreads / 32][2];
    __shared__ unsigned s_prefix, s_remaining, s_taken_less, s_taken_eq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.y, j = blockIdx.x;
    const float *pts = data + (size_t)cloud * n * 3;
    const float *q = query + ((size_t)cloud * m + j) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    uint32_t *keys = scratch + ((size_t)cloud * m + j) * n;

    // ---- 1. keys (squared distances are >= 0, so the float bit pattern is already order-preserving)
    for (int p = tid; p < n; p += kKlThreads) {
        const float d = d2_ref_cpu(__fsub_rn(qx, __ldg(pts + (size_t)p * 3)), __fsub_rn(qy, __ldg(pts + (size_t)p * 3 + 1)),
                                   __fsub_rn(qz, __ldg(pts + (size_t)p * 3 + 2)));
        keys[p] = __float_as_uint(d);
    }
    if (tid == 0) { s_prefix = 0; s_remaining = (unsigned)k; }
    __syncthreads();

    // ---- 2. radix select, most significant byte first
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int p = tid; p < n; p += kKlThreads) {
            const unsigned key = keys[p];
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rem = s_remaining, bin = 0;
            for (; bin < 256; ++bin) {
                const unsigned c = s_hist[bin];
                if (c >= rem) break;
                rem -= c;
            }
            s_prefix = prefix | (bin << shift);
            s_remaining = rem;  // how many of the keys equal to the (eventual) threshold prefix are still needed
        }
        __syncthreads();
    }
    const unsigned thr = s_prefix;         // the k-th smallest key
    const unsigned need_eq = s_remaining;  // ties to take, in ascending index
    if (tid == 0) { s_taken_less = 0; s_taken_eq = 0; }
    __syncthreads();

    // ---- 3. ordered compaction, 256 candidates per step
    for (int p0 = 0; p0 < n; p0 += kKlThreads) {
        const int p = p0 + tid;
        const unsigned key = p < n ? keys[p] : 0xffffffffu;
        const bool less = p < n && key < thr, eq = p < n && key == thr;
        const unsigned bl = __ballot_sync(kFull, less), be = __ballot_sync(kFull, eq);
        if (lane == 0) { s_warp[warp][0] = __popc(bl); s_warp[warp][1] = __popc(be); }
        __syncthreads();
        unsigned off_l = s_taken_less, off_e = s_taken_eq, tot_l = 0, tot_e = 0;
#pragma unroll
        for (int w = 0; w < kKlThreads / 32; ++w) {
            if (w < warp) { off_l += s_warp[w][0]; off_e += s_warp[w][1]; }
            tot_l += s_warp[w][0];
            tot_e += s_warp[w][1];
        }
        const unsigned lt = (1u << lane) - 1u;
        if (less) {
            const unsigned pos = off_l + __popc(bl & lt);                  // slots [0, k - need_eq)
            s_sel[pos] = ((unsigned long long)key << 32) | (unsigned)p;
        } else if (eq) {
            const unsigned r = off_e + __popc(be & lt);
            if (r < need_eq) s_sel[(unsigned)k - need_eq + r] = ((unsigned long long)key << 32) | (unsigned)p;
        }
        __syncthreads();
        if (tid == 0) { s_taken_less += tot_l; s_taken_eq += tot_e; }
        __syncthreads();
    }
    // ---- 4. bitonic sort of the k pairs (padded with +inf keys to a power of two)
    int kp = 1;
    while (kp < k) kp <<= 1;
    for (int e = k + tid; e < kp; e += kKlThreads) s_sel[e] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= kp; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = tid; e < kp / 2; e += kKlThreads) {
                const int lo = 2 * e - (e & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = s_sel[lo], b2 = s_sel[hi];
                if ((a > b2) == up) { s_sel[lo] = b2; s_sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    const size_t o = ((size_t)cloud * m + j) * k;
    for (int e = tid; e < k; e += kKlThreads) {
        const unsigned long long v = s_sel[e];
        idx[o + e] = (int)(unsigned)(v & 0xffffffffull);
        if (dist) dist[o + e] = __uint_as_float((unsigned)(v >> 32));
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_knn_large_workspace_bytes(int b, int n, int m) {
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return (size_t)b * m * n * sizeof(uint32_t);
}

extern "C" int dispu_knn_large(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *dist, int *idx,
                               void *workspace, size_t workspace_bytes, dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0 && k > 0 && k <= n, DISPU_EINVAL);
    DISPU_REQUIRE(k <= kKlMaxK && b <= 65535, DISPU_EUNSUPPORTED);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && idx, DISPU_ENULL);
    DISPU_REQUIRE(workspace && workspace_bytes >= dispu_knn_large_workspace_bytes(b, n, m), DISPU_EWORKSPACE);
    count_launch();
    knn_large_kernel<<<dim3(m, b), kKlThreads, 0, (cudaStream_t)stream>>>(n, m, k, xyz1, xyz2, dist, idx,
                                                                        reinterpret_cast<uint32_t *>(workspace));
    return finish_launch();
}

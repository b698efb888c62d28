// Brute-force k-nearest-neighbour kernels in 3-D for sm_100a.
//
//  * dispu_knn      — kNN grouping: replaces the nanoflann KD-tree round trip the live generator
//                     makes every step (libs/nearest_neighbors/knn_.cxx:104-135 through
//                     tf.py_func, Common/ops.py:110-118,165: device -> host -> KD-tree build per
//                     cloud -> host -> device) and the pure-TF knn_point
//                     (tf_ops/grouping/tf_grouping.py:116-141), which materialises a (b,m,n)
//                     distance tensor before a full top_k.
//  * dispu_three_nn — ThreeNN: replaces the single-threaded host loop threenn_cpu
//                     (tf_ops/interpolation/tf_interpolate.cpp:60-103); the reference registers
//                     no GPU kernel for it.
//
// Both: the dataset cloud is staged in shared memory as float4 tiles (one broadcast LDS.128 per
// candidate per warp), one thread owns one query and keeps its running top-k sorted in registers
// (statically indexed, strict '<' insertion so equal distances keep the lower index first — the
// order of tf.nn.top_k and of the reference's cascades).  Distances use the unfused host-loop
// contraction (d2_ref_cpu) so they are bit-identical to nanoflann's / threenn_cpu's floats.
#include "common.cuh"

namespace dispu {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 2048;  // float4 points per tile (32 KB)

template <int KT>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int n, int m, int k, const float *__restrict__ data, const float *__restrict__ query,
           float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKnnTile];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kKnnThreads + threadIdx.x;
    const float *pts = data + (size_t)cloud * n * 3;
    const bool active = j < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *q = query + ((size_t)cloud * m + j) * 3;
        qx = __ldg(q); qy = __ldg(q + 1); qz = __ldg(q + 2);
    }
    float bd[KT];
    int bi[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        bd[i] = __int_as_float(0x7f800000);  // +inf
        bi[i] = 0;
    }
    for (int t0 = 0; t0 < n; t0 += kKnnTile) {
        const int tn = min(kKnnTile, n - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < tn; e += kKnnThreads) {
            const float *p = pts + (size_t)(t0 + e) * 3;
            s_tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (active) {
            for (int e = 0; e < tn; ++e) {
                const float4 p = s_tile[e];
                // diff = query - data, accumulated dimension by dimension (nanoflann.hpp:323-348)
                const float d = d2_ref_cpu(__fsub_rn(qx, p.x), __fsub_rn(qy, p.y), __fsub_rn(qz, p.z));
                if (d < bd[KT - 1]) {
                    bd[KT - 1] = d;
                    bi[KT - 1] = t0 + e;
#pragma unroll
                    for (int i = KT - 1; i > 0; --i) {
                        if (bd[i] < bd[i - 1]) {
                            const float tf = bd[i]; bd[i] = bd[i - 1]; bd[i - 1] = tf;
                            const int ti = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (active) {
        const size_t o = ((size_t)cloud * m + j) * k;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            if (i < k) {
                if (dist) dist[o + i] = bd[i];
                idx[o + i] = bi[i];
            }
        }
    }
}

// The KT-slot list holds the KT smallest; the first k of them are the k smallest, so running
// with KT >= k is exact.
template <int KT>
static int launch_knn(int b, int n, int m, int k, const float *data, const float *query, float *dist, int *idx,
                      cudaStream_t stream) {
    count_launch();
    knn_kernel<KT><<<dim3(ceil_div(m, kKnnThreads), b), kKnnThreads, 0, stream>>>(n, m, k, data, query, dist, idx);
    return finish_launch();
}

// ThreeNN: (x2-x1) with x2 the known point, strict '<' cascades (tf_interpolate.cpp:77-91).
// The reference compares in double against 1e40; for float inputs that is the float order with a
// +inf start (a finite float is always < 1e40, inf/NaN never are), and float(1e40) = +inf is what
// it stores when fewer than three known points exist.
__global__ void __launch_bounds__(kKnnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 s_tile[kKnnTile];
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kKnnThreads + threadIdx.x;
    const float *known = xyz2 + (size_t)cloud * m * 3;
    const bool active = j < n;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float *u = xyz1 + ((size_t)cloud * n + j) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    const float inf = __int_as_float(0x7f800000);
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int t0 = 0; t0 < m; t0 += kKnnTile) {
        const int tn = min(kKnnTile, m - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < tn; e += kKnnThreads) {
            const float *p = known + (size_t)(t0 + e) * 3;
            s_tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (active) {
            for (int e = 0; e < tn; ++e) {
                const float4 p = s_tile[e];
                const float d = d2_ref_cpu(__fsub_rn(p.x, ux), __fsub_rn(p.y, uy), __fsub_rn(p.z, uz));
                const int k = t0 + e;
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else if (d < b3) {
                    b3 = d; i3 = k;
                }
            }
        }
    }
    if (active) {
        const size_t o = ((size_t)cloud * n + j) * 3;
        dist[o] = b1; dist[o + 1] = b2; dist[o + 2] = b3;
        idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *dist, int *idx,
                         dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(k >= 1 && k <= n, DISPU_EINVAL);
    DISPU_REQUIRE(k <= 64, DISPU_EUNSUPPORTED);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && idx, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    if (k <= 4) return launch_knn<4>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 8) return launch_knn<8>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 16) return launch_knn<16>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 24) return launch_knn<24>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    if (k <= 32) return launch_knn<32>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
    return launch_knn<64>(b, n, m, k, xyz1, xyz2, dist, idx, stream);
}

extern "C" int dispu_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                              dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n >= 0 && m > 0, DISPU_EINVAL);
    if (b == 0 || n == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz1 && xyz2 && dist && idx, DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    count_launch();
    three_nn_kernel<<<dim3(ceil_div(n, kKnnThreads), b), kKnnThreads, 0, stream>>>(n, m, xyz1, xyz2, dist, idx);
    return finish_launch();
}

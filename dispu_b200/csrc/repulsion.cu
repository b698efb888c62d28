// Repulsion loss, fused — the other live consumer of QueryBallPoint / GroupPoint in the training step
// (Common/loss_utils.py:271-298 get_repulsion_loss, DisPU/model.py:83):
//     idx = query_ball_point(radius, nsample, pred, pred); g = group_point(pred, idx) - pred[:, :, None]
//     d2 = sum(g^2, -1); val = top_k(-d2, 5)[:, :, 1:]; loss = mean(max(0, h + val))
// The reference materialises (b,n,nsample,3), (b,n,nsample) and runs a full top_k.  Here one thread per point
// reads its idx row, keeps the five smallest squared distances in registers, drops the smallest (the point
// itself) and accumulates the hinge; CTAs write fixed-order partial sums and a one-CTA pass finalises, so
// the scalar is deterministic.  The ball query itself is dispu_query_ball_point(_ws).
#include "common.cuh"

namespace dispu {

constexpr int kRepThreads = 256;

__global__ void __launch_bounds__(kRepThreads)
repulsion_partial_kernel(long long total, int n, int ns, float h, const float *__restrict__ pred,
                         const int *__restrict__ idx, float *__restrict__ partials) {
    __shared__ float s_red[kRepThreads / 32];
    const long long p = (long long)blockIdx.x * kRepThreads + threadIdx.x;
    float acc = 0.f;
    if (p < total) {
        const long long cloud = p / n;
        const float *c = pred + p * 3;
        const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
        const float inf = __int_as_float(0x7f800000);
        float d0 = inf, d1 = inf, d2 = inf, d3 = inf, d4 = inf;  // ascending
        const int *row = idx + p * ns;
        for (int s = 0; s < ns; ++s) {
            const float *q = pred + (cloud * n + __ldg(row + s)) * 3;
            const float gx = __fsub_rn(__ldg(q), cx), gy = __fsub_rn(__ldg(q + 1), cy), gz = __fsub_rn(__ldg(q + 2), cz);
            float d = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
            // insert into the sorted five (compare-exchange chain)
            float t;
            if (d < d0) { t = d0; d0 = d; d = t; }
            if (d < d1) { t = d1; d1 = d; d = t; }
            if (d < d2) { t = d2; d2 = d; d = t; }
            if (d < d3) { t = d3; d3 = d; d = t; }
            if (d < d4) { d4 = d; }
        }
        // top_k(-d2, 5)[1:] : drop the smallest, hinge on the next four (fewer than 5 samples: +inf -> 0)
        acc = __fadd_rn(__fadd_rn(fmaxf(0.f, __fsub_rn(h, d1)), fmaxf(0.f, __fsub_rn(h, d2))),
                        __fadd_rn(fmaxf(0.f, __fsub_rn(h, d3)), fmaxf(0.f, __fsub_rn(h, d4))));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRepThreads / 32; ++w) t += s_red[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
repulsion_finalize_kernel(int nparts, long long count, const float *__restrict__ partials, float *__restrict__ out) {
    __shared__ double s_red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) acc += (double)partials[i];
    s_red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 256; ++i) t += s_red[i];
        out[0] = (float)(t / (double)count);  // mean over b*n*4
        out[1] = (float)t;                    // sum (multi-GPU partial)
    }
}

// One percentage term of get_uniform_loss (Common/loss_utils.py:238-269), fused after FPS + ball query:
//     g = group_point(pcd, idx) (b,npoint,nsample,3); per group the 2-NN of every grouped point INSIDE the group
//     (knn_point(2, g, g): the smallest is the point itself), u = sqrt(|d2 + 1e-8|),
//     term = mean over (b,npoint,nsample) of (u - expect_len)^2 / (expect_len + 1e-8).
// One thread per grouped point scans its group's nsample entries (padding repeats included, exactly like top_k on
// the materialised (nsample,nsample) matrix) and keeps the two smallest squared distances.
__global__ void __launch_bounds__(kRepThreads)
uniform_partial_kernel(long long total, int n, int m_ns, int ns, float expect_len, const float *__restrict__ pcd,
                       const int *__restrict__ idx, float *__restrict__ partials) {
    __shared__ float s_red[kRepThreads / 32];
    const long long e = (long long)blockIdx.x * kRepThreads + threadIdx.x;  // (b, j, s)
    float acc = 0.f;
    if (e < total) {
        const long long cloud = e / m_ns;
        const long long grp = e / ns;
        const int *row = idx + grp * ns;
        const float *c = pcd + (cloud * n + __ldg(idx + e)) * 3;
        const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
        const float inf = __int_as_float(0x7f800000);
        float d0 = inf, d1 = inf;
        for (int s = 0; s < ns; ++s) {
            const float *q = pcd + (cloud * n + __ldg(row + s)) * 3;
            const float gx = __fsub_rn(__ldg(q), cx), gy = __fsub_rn(__ldg(q + 1), cy), gz = __fsub_rn(__ldg(q + 2), cz);
            float d = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
            if (d < d0) { const float t = d0; d0 = d; d = t; }
            if (d < d1) d1 = d;
        }
        if (ns < 2) d1 = 0.f;
        const float u = __fsqrt_rn(fabsf(__fadd_rn(d1, 1e-8f)));
        const float t = __fsub_rn(u, expect_len);
        acc = __fdiv_rn(__fmul_rn(t, t), __fadd_rn(expect_len, 1e-8f));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRepThreads / 32; ++w) t += s_red[w];
        partials[blockIdx.x] = t;
    }
}

}  // namespace dispu

using namespace dispu;

extern "C" size_t dispu_uniform_workspace_bytes(int b, int npoint, int nsample) {
    if (b <= 0 || npoint <= 0 || nsample <= 0) return 0;
    return sizeof(float) * (size_t)ceil_div64((long long)b * npoint * nsample, kRepThreads);
}

extern "C" int dispu_uniform_loss_term(int b, int n, int npoint, int nsample, float expect_len, const float *pcd,
                                       const int *idx, float *out, void *workspace, size_t workspace_bytes,
                                       dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b > 0 && n > 0 && npoint > 0 && nsample > 0, DISPU_EINVAL);
    DISPU_REQUIRE(pcd && idx && out, DISPU_ENULL);
    DISPU_REQUIRE(workspace && workspace_bytes >= dispu_uniform_workspace_bytes(b, npoint, nsample), DISPU_EWORKSPACE);
    const long long total = (long long)b * npoint * nsample;
    const int nparts = (int)ceil_div64(total, kRepThreads);
    count_launch(2);
    uniform_partial_kernel<<<nparts, kRepThreads, 0, stream>>>(total, n, npoint * nsample, nsample, expect_len, pcd, idx,
                                                               reinterpret_cast<float *>(workspace));
    repulsion_finalize_kernel<<<1, 256, 0, stream>>>(nparts, total, reinterpret_cast<float *>(workspace), out);
    return finish_launch();
}

extern "C" size_t dispu_repulsion_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    return sizeof(float) * (size_t)ceil_div64((long long)b * n, kRepThreads);
}

extern "C" int dispu_repulsion_loss(int b, int n, int nsample, float h, const float *pred, const int *idx, float *out,
                                    void *workspace, size_t workspace_bytes, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b > 0 && n > 0 && nsample > 0, DISPU_EINVAL);
    DISPU_REQUIRE(pred && idx && out, DISPU_ENULL);
    DISPU_REQUIRE(workspace && workspace_bytes >= dispu_repulsion_workspace_bytes(b, n), DISPU_EWORKSPACE);
    const long long total = (long long)b * n;
    const int nparts = (int)ceil_div64(total, kRepThreads);
    count_launch(2);
    repulsion_partial_kernel<<<nparts, kRepThreads, 0, stream>>>(total, n, nsample, h, pred, idx,
                                                                 reinterpret_cast<float *>(workspace));
    repulsion_finalize_kernel<<<1, 256, 0, stream>>>(nparts, total * 4, reinterpret_cast<float *>(workspace), out);
    return finish_launch();
}

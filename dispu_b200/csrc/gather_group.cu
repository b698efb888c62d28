// Index-driven copies of the path: GatherPoint(+Grad), GroupPoint(+Grad), ThreeInterpolate(+Grad)
// and the inverse-distance weights.  All are HBM/L2-bound gathers: one thread per OUTPUT element
// (or float4 of elements) so stores are fully coalesced and the gathered rows are read through
// L1/L2 (each source row is reused nsample/ npoint times).  Grid = enough CTAs to cover the
// output, capped at a multiple of the SM count with a grid-stride loop.
//
// Reference kernels replaced: gatherpointKernel / scatteraddpointKernel
// (tf_ops/sampling/tf_sampling_g.cu:172,183), group_point_gpu / group_point_grad_gpu
// (tf_ops/grouping/tf_grouping_g.cu:40,61), threeinterpolate_cpu / _grad_cpu
// (tf_ops/interpolation/tf_interpolate.cpp:107,131 — host loops in the reference).
#include "common.cuh"

namespace dispu {

constexpr int kThreads = 256;
static inline int grid_for(long long work_items) {
    long long g = ceil_div64(work_items, kThreads);
    const long long cap = (long long)kNumSMs * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// out[row, :] = src[cloud(row), idx[row], :]   rows = b*rows_per_cloud, c floats each.
// VEC = 4 when c % 4 == 0 (16-byte rows), else scalar.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(long long total_vec, int rows_per_cloud, int n, int c_vec, const float *__restrict__ src,
                   const int *__restrict__ idx, float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total_vec;
         e += (long long)gridDim.x * kThreads) {
        const long long row = e / c_vec;
        const int l = (int)(e - row * c_vec);
        const long long cloud = row / rows_per_cloud;
        const int ii = __ldg(idx + row);
        const long long s = (cloud * n + ii) * (long long)c_vec + l;
        if (VEC == 4)
            reinterpret_cast<float4 *>(out)[e] = __ldg(reinterpret_cast<const float4 *>(src) + s);
        else
            out[e] = __ldg(src + s);
    }
}

// dst[cloud(row), idx[row], :] += g[row, :]   (dst pre-zeroed)
__global__ void __launch_bounds__(kThreads)
scatter_add_rows_kernel(long long total, int rows_per_cloud, int n, int c, const float *__restrict__ g,
                        const int *__restrict__ idx, float *__restrict__ dst) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total;
         e += (long long)gridDim.x * kThreads) {
        const long long row = e / c;
        const int l = (int)(e - row * c);
        const long long cloud = row / rows_per_cloud;
        const int ii = __ldg(idx + row);
        atomicAdd(dst + (cloud * n + ii) * (long long)c + l, __ldg(g + e));
    }
}

// out[b,j,l] = (p[i1,l]*w1 + p[i2,l]*w2) + p[i3,l]*w3   — unfused like the host loop
// (tf_interpolate.cpp:121).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
three_interpolate_kernel(long long total_vec, int n, int m, int c_vec, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total_vec;
         e += (long long)gridDim.x * kThreads) {
        const long long row = e / c_vec;  // b*n + j
        const int l = (int)(e - row * c_vec);
        const long long cloud = row / n;
        const int i1 = __ldg(idx + row * 3), i2 = __ldg(idx + row * 3 + 1), i3 = __ldg(idx + row * 3 + 2);
        const float w1 = __ldg(weight + row * 3), w2 = __ldg(weight + row * 3 + 1), w3 = __ldg(weight + row * 3 + 2);
        const long long pb = cloud * m;
        if (VEC == 4) {
            const float4 *p4 = reinterpret_cast<const float4 *>(points);
            const float4 a = __ldg(p4 + (pb + i1) * c_vec + l), bq = __ldg(p4 + (pb + i2) * c_vec + l),
                         cq = __ldg(p4 + (pb + i3) * c_vec + l);
            float4 o;
            o.x = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w1), __fmul_rn(bq.x, w2)), __fmul_rn(cq.x, w3));
            o.y = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w1), __fmul_rn(bq.y, w2)), __fmul_rn(cq.y, w3));
            o.z = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w1), __fmul_rn(bq.z, w2)), __fmul_rn(cq.z, w3));
            o.w = __fadd_rn(__fadd_rn(__fmul_rn(a.w, w1), __fmul_rn(bq.w, w2)), __fmul_rn(cq.w, w3));
            reinterpret_cast<float4 *>(out)[e] = o;
        } else {
            const float a = __ldg(points + (pb + i1) * c_vec + l), bq = __ldg(points + (pb + i2) * c_vec + l),
                        cq = __ldg(points + (pb + i3) * c_vec + l);
            out[e] = __fadd_rn(__fadd_rn(__fmul_rn(a, w1), __fmul_rn(bq, w2)), __fmul_rn(cq, w3));
        }
    }
}

__global__ void __launch_bounds__(kThreads)
three_interpolate_grad_kernel(long long total, int n, int m, int c, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total;
         e += (long long)gridDim.x * kThreads) {
        const long long row = e / c;
        const int l = (int)(e - row * c);
        const long long cloud = row / n;
        const float go = __ldg(grad_out + e);
        float *g = grad_points + cloud * m * (long long)c + l;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            atomicAdd(g + (long long)__ldg(idx + row * 3 + r) * c, __fmul_rn(go, __ldg(weight + row * 3 + r)));
    }
}

// pointnet_util.py:205-208: d=max(d,1e-10); inv=1/d; w=inv/(inv0+inv1+inv2).  IEEE division.
__global__ void __launch_bounds__(kThreads)
three_nn_weights_kernel(long long rows, const float *__restrict__ dist, float *__restrict__ weight) {
    for (long long r = (long long)blockIdx.x * kThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kThreads) {
        const float i0 = __fdiv_rn(1.0f, fmaxf(__ldg(dist + r * 3), 1e-10f));
        const float i1 = __fdiv_rn(1.0f, fmaxf(__ldg(dist + r * 3 + 1), 1e-10f));
        const float i2 = __fdiv_rn(1.0f, fmaxf(__ldg(dist + r * 3 + 2), 1e-10f));
        const float norm = __fadd_rn(__fadd_rn(i0, i1), i2);
        weight[r * 3] = __fdiv_rn(i0, norm);
        weight[r * 3 + 1] = __fdiv_rn(i1, norm);
        weight[r * 3 + 2] = __fdiv_rn(i2, norm);
    }
}


// Fused GroupPoint + centre term + ReLU:  out[b,i,s,:] = relu(P[b,idx[b,i,s],:] + sign * Q[b,i,:]).
// This is the first layer of a grouped MLP after hoisting it to per-point GEMMs: for grouped input
// [x_j - x_i, x_j, f_j] and weights [Wa; Wb; Wc],  row.W + b = (x_j(Wa+Wb) + f_j Wc + b) - x_i Wa = P[j] - Q[i]
// (Common/ops.py:1033-1058 PointShuffle2 conv0; :1897-1905 dense_conv l0).  The (b,m,ns,C_in) grouped /
// concatenated tensors of the reference are never written.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
group_affine_relu_kernel(long long total_vec, int m_ns, int ns, int n, int c_vec, float sign,
                         const float *__restrict__ P, const float *__restrict__ Q, const int *__restrict__ idx,
                         float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total_vec;
         e += (long long)gridDim.x * kThreads) {
        const long long row = e / c_vec;  // (b, i, s)
        const int l = (int)(e - row * c_vec);
        const long long cloud = row / m_ns;
        const long long grp = row / ns;   // (b, i)
        const int ii = __ldg(idx + row);
        const long long ps = (cloud * n + ii) * (long long)c_vec + l;
        const long long qs = grp * (long long)c_vec + l;
        if (VEC == 4) {
            const float4 p = __ldg(reinterpret_cast<const float4 *>(P) + ps);
            const float4 q = Q ? __ldg(reinterpret_cast<const float4 *>(Q) + qs) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = fmaxf(__fmaf_rn(sign, q.x, p.x), 0.f);
            o.y = fmaxf(__fmaf_rn(sign, q.y, p.y), 0.f);
            o.z = fmaxf(__fmaf_rn(sign, q.z, p.z), 0.f);
            o.w = fmaxf(__fmaf_rn(sign, q.w, p.w), 0.f);
            reinterpret_cast<float4 *>(out)[e] = o;
        } else {
            const float q = Q ? __ldg(Q + qs) : 0.f;
            out[e] = fmaxf(__fmaf_rn(sign, q, __ldg(P + ps)), 0.f);
        }
    }
}

// Fused GroupPoint + reduce_max over nsample:  out[b,i,:] = max_s points[b,idx[b,i,s],:]
// (e.g. the skip connection of PointShuffle2, Common/ops.py:1049: reduce_max(grouped_feat, axis=2)).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
group_max_kernel(long long total_vec, int m, int ns, int n, int c_vec, const float *__restrict__ points, long long ldp_vec,
                 const int *__restrict__ idx, float *__restrict__ out, long long ldo_vec) {
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total_vec;
         e += (long long)gridDim.x * kThreads) {
        const long long grp = e / c_vec;  // (b, i)
        const int l = (int)(e - grp * c_vec);
        const long long cloud = grp / m;
        const int *row = idx + grp * ns;
        if (VEC == 4) {
            const float4 *p4 = reinterpret_cast<const float4 *>(points);
            float4 mx = __ldg(p4 + (cloud * n + __ldg(row)) * ldp_vec + l);
            for (int s = 1; s < ns; ++s) {
                const float4 v = __ldg(p4 + (cloud * n + __ldg(row + s)) * ldp_vec + l);
                mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
            }
            reinterpret_cast<float4 *>(out)[grp * ldo_vec + l] = mx;
        } else {
            float mx = __ldg(points + (cloud * n + __ldg(row)) * ldp_vec + l);
            for (int s = 1; s < ns; ++s) mx = fmaxf(mx, __ldg(points + (cloud * n + __ldg(row + s)) * ldp_vec + l));
            out[grp * ldo_vec + l] = mx;
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int gather_rows(int b, int n, int c, long long rows_per_cloud, const float *src, const int *idx, float *out,
                       cudaStream_t stream) {
    const long long rows = (long long)b * rows_per_cloud;
    if (rows == 0 || c == 0) return DISPU_OK;
    count_launch();
    if (c % 4 == 0 && aligned16(src) && aligned16(out)) {
        const long long total = rows * (c / 4);
        gather_rows_kernel<4><<<grid_for(total), kThreads, 0, stream>>>(total, (int)rows_per_cloud, n, c / 4, src, idx, out);
    } else {
        const long long total = rows * c;
        gather_rows_kernel<1><<<grid_for(total), kThreads, 0, stream>>>(total, (int)rows_per_cloud, n, c, src, idx, out);
    }
    return finish_launch();
}

static int scatter_add_rows(int b, int n, int c, long long rows_per_cloud, const float *g, const int *idx, float *dst,
                            cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)b * n * c, stream);
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)b * rows_per_cloud * c;
    if (total == 0) return DISPU_OK;
    count_launch();
    scatter_add_rows_kernel<<<grid_for(total), kThreads, 0, stream>>>(total, (int)rows_per_cloud, n, c, g, idx, dst);
    return finish_launch();
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out,
                                  dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(inp && idx && out, DISPU_ENULL);
    return gather_rows(b, n, 3, m, inp, idx, out, (cudaStream_t)stream);
}

extern "C" int dispu_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g,
                                       dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(inp_g && (m == 0 || (out_g && idx)), DISPU_ENULL);
    return scatter_add_rows(b, n, 3, m, out_g, idx, inp_g, (cudaStream_t)stream);
}

extern "C" int dispu_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                                 float *out, dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample >= 0, DISPU_EINVAL);
    if (b == 0 || m == 0 || nsample == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(points && idx && out, DISPU_ENULL);
    return gather_rows(b, n, c, (long long)m * nsample, points, idx, out, (cudaStream_t)stream);
}

extern "C" int dispu_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                      float *grad_points, dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample >= 0, DISPU_EINVAL);
    if (b == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(grad_points && (m == 0 || nsample == 0 || (grad_out && idx)), DISPU_ENULL);
    return scatter_add_rows(b, n, c, (long long)m * nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}

extern "C" int dispu_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                                       const float *weight, float *out, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && m > 0 && c >= 0 && n >= 0, DISPU_EINVAL);
    if (b == 0 || n == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(points && idx && weight && out, DISPU_ENULL);
    count_launch();
    if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
        const long long total = (long long)b * n * (c / 4);
        three_interpolate_kernel<4><<<grid_for(total), kThreads, 0, stream>>>(total, n, m, c / 4, points, idx, weight, out);
    } else {
        const long long total = (long long)b * n * c;
        three_interpolate_kernel<1><<<grid_for(total), kThreads, 0, stream>>>(total, n, m, c, points, idx, weight, out);
    }
    return finish_launch();
}

extern "C" int dispu_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                            const float *weight, float *grad_points, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && m > 0 && c >= 0 && n >= 0, DISPU_EINVAL);
    if (b == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(grad_points && (n == 0 || (grad_out && idx && weight)), DISPU_ENULL);
    cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, stream);
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)b * n * c;
    if (total == 0) return DISPU_OK;
    count_launch();
    three_interpolate_grad_kernel<<<grid_for(total), kThreads, 0, stream>>>(total, n, m, c, grad_out, idx, weight, grad_points);
    return finish_launch();
}

extern "C" int dispu_three_nn_weights(int b, int n, const float *dist, float *weight, dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n >= 0, DISPU_EINVAL);
    const long long rows = (long long)b * n;
    if (rows == 0) return DISPU_OK;
    DISPU_REQUIRE(dist && weight, DISPU_ENULL);
    count_launch();
    three_nn_weights_kernel<<<grid_for(rows), kThreads, 0, (cudaStream_t)stream>>>(rows, dist, weight);
    return finish_launch();
}

extern "C" int dispu_group_affine_relu(int b, int n, int c, int m, int nsample, const float *P, const float *Q,
                                       float sign, const int *idx, float *out, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample >= 0, DISPU_EINVAL);
    if (b == 0 || m == 0 || nsample == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(P && idx && out, DISPU_ENULL);
    const long long rows = (long long)b * m * nsample;
    count_launch();
    if (c % 4 == 0 && aligned16(P) && aligned16(out) && (!Q || aligned16(Q))) {
        const long long total = rows * (c / 4);
        group_affine_relu_kernel<4><<<grid_for(total), kThreads, 0, stream>>>(total, m * nsample, nsample, n, c / 4, sign, P, Q,
                                                                            idx, out);
    } else {
        const long long total = rows * c;
        group_affine_relu_kernel<1><<<grid_for(total), kThreads, 0, stream>>>(total, m * nsample, nsample, n, c, sign, P, Q, idx,
                                                                            out);
    }
    return finish_launch();
}

extern "C" int dispu_group_max_ld(int b, int n, int c, int m, int nsample, const float *points, int ldp, const int *idx,
                                  float *out, int ldo, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && nsample > 0, DISPU_EINVAL);
    DISPU_REQUIRE(ldp >= c && ldo >= c, DISPU_EINVAL);
    if (b == 0 || m == 0 || c == 0) return DISPU_OK;
    DISPU_REQUIRE(points && idx && out, DISPU_ENULL);
    const long long groups = (long long)b * m;
    count_launch();
    if (c % 4 == 0 && ldp % 4 == 0 && ldo % 4 == 0 && aligned16(points) && aligned16(out)) {
        const long long total = groups * (c / 4);
        group_max_kernel<4><<<grid_for(total), kThreads, 0, stream>>>(total, m, nsample, n, c / 4, points, ldp / 4, idx, out,
                                                                      ldo / 4);
    } else {
        const long long total = groups * c;
        group_max_kernel<1><<<grid_for(total), kThreads, 0, stream>>>(total, m, nsample, n, c, points, ldp, idx, out, ldo);
    }
    return finish_launch();
}

extern "C" int dispu_group_max(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                               dispu_stream_t stream_) {
    return dispu_group_max_ld(b, n, c, m, nsample, points, c, idx, out, c, stream_);
}

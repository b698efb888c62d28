// Fused set-abstraction body for sm_100a: group -> centre-subtract -> concat -> 1x1-conv MLP
// (bias, ReLU) -> max over nsample.  Replaces the chain the reference builds out of separate TF
// ops, each of which round-trips a (b,npoint,nsample,C) tensor through HBM
// (Common/pointnet_util.py:46-53 group_point x2 / tile / subtract / concat, :117-121
// tf_util.conv2d per layer = Common/tf_util.py:120-185, :133 reduce_max).  Nothing of shape
// (b,npoint,nsample,*) is ever written to global memory here: the only HBM traffic is idx, the
// gathered rows (L2-resident after first touch), the weights and the (b,npoint,C_out) result.
//
// This file holds the exact-FP32 path (DISPU_MLP_FP32): one thread owns one grouped row, its
// activations ping-pong between two private shared-memory rows (odd stride: conflict-free), each
// layer is a register-tiled FMA chain over C_in in ascending order, and the pool is a column max
// over the CTA's rows.  The tensor-core path (DISPU_MLP_TF32X3, tcgen05) lives in
// sa_mlp_tc.cu.
#include "sa_mlp.cuh"

namespace dispu {

constexpr int kMlpThreads = 128;

// rows of a CTA pass: kMlpThreads; groups per pass G = kMlpThreads / ns when ns divides 128,
// else one group per pass processed in chunks of 128 rows.
__global__ void __launch_bounds__(kMlpThreads)
sa_mlp_fp32_kernel(int n, int m, int ns, int c_pts, int layout, int use_xyz, int pooling,
                   const float *__restrict__ xyz, const float *__restrict__ points, const float *__restrict__ new_xyz,
                   const int *__restrict__ idx, MlpParams prm, float *__restrict__ out, long long n_groups,
                   int stride /* odd, >= max width */) {
    extern __shared__ float s_buf[];  // 2 * kMlpThreads * stride
    __shared__ float s_wt[kMlpThreads];  // weighted_avg: exp(-5 |grouped_xyz|) of each row (pointnet_util.py:125-127)
    float *buf0 = s_buf + (size_t)threadIdx.x * stride;
    float *buf1 = s_buf + (size_t)(kMlpThreads + threadIdx.x) * stride;
    const int t = threadIdx.x;
    const int c_in = 3 + c_pts;
    const int c_out = prm.widths[prm.n_layers];
    const bool multi = ns <= kMlpThreads && (kMlpThreads % ns) == 0;
    const int G = multi ? kMlpThreads / ns : 1;
    const int chunks = multi ? 1 : ceil_div(ns, kMlpThreads);
    const int o_stride = pooling == DISPU_POOL_MAX_AND_AVG ? 2 * c_out : c_out;
    float wtot = 0.f;  // chunked weighted_avg: running sum of the group's weights

    for (long long g0 = (long long)blockIdx.x * G; g0 < n_groups; g0 += (long long)gridDim.x * G) {
        for (int ch = 0; ch < chunks; ++ch) {
            // ---- which grouped row is mine
            const int gi = multi ? t / ns : 0;
            const int s = multi ? t % ns : ch * kMlpThreads + t;
            const long long g = g0 + gi;
            const bool valid = g < n_groups && s < ns;
            float *x = buf0, *y = buf1;
            if (valid) {
                const long long cloud = g / m;
                const int ii = __ldg(idx + g * ns + s);
                const float *pxyz = xyz + (cloud * n + ii) * 3;
                const float *ctr = new_xyz + g * 3;
                const float gx = __fsub_rn(__ldg(pxyz), __ldg(ctr));
                const float gy = __fsub_rn(__ldg(pxyz + 1), __ldg(ctr + 1));
                const float gz = __fsub_rn(__ldg(pxyz + 2), __ldg(ctr + 2));
                const int xo = layout == 0 ? 0 : c_pts, po = (layout == 0 && use_xyz) ? 3 : 0;
                if (use_xyz) { x[xo] = gx; x[xo + 1] = gy; x[xo + 2] = gz; }
                if (pooling == DISPU_POOL_WEIGHTED_AVG)
                    s_wt[t] = expf(-5.f * sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz))));
                if (c_pts) {
                    const float *pf = points + (cloud * n + ii) * (long long)c_pts;
                    for (int l = 0; l < c_pts; ++l) x[po + l] = __ldg(pf + l);
                }
                // ---- MLP, private to this thread's row
                for (int l = 0; l < prm.n_layers; ++l) {
                    const int ci = prm.widths[l], co = prm.widths[l + 1];
                    const float *W = prm.w[l];
                    const float *B = prm.bias[l];
                    const float *S = prm.bn_scale[l], *T = prm.bn_shift[l];
                    int oc = 0;
                    for (; oc + 8 <= co; oc += 8) {
                        float acc[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
                        for (int ic = 0; ic < ci; ++ic) {
                            const float xv = x[ic];
                            const float *wr = W + (size_t)ic * co + oc;
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc[u] = __fmaf_rn(xv, __ldg(wr + u), acc[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float v = B ? __fadd_rn(acc[u], __ldg(B + oc + u)) : acc[u];
                            if (S) v = __fmaf_rn(v, __ldg(S + oc + u), __ldg(T + oc + u));
                            y[oc + u] = fmaxf(v, 0.f);
                        }
                    }
                    for (; oc < co; ++oc) {
                        float acc = 0.f;
                        for (int ic = 0; ic < ci; ++ic) acc = __fmaf_rn(x[ic], __ldg(W + (size_t)ic * co + oc), acc);
                        float v = B ? __fadd_rn(acc, __ldg(B + oc)) : acc;
                        if (S) v = __fmaf_rn(v, __ldg(S + oc), __ldg(T + oc));
                        y[oc] = fmaxf(v, 0.f);
                    }
                    float *tmp = x; x = y; y = tmp;
                }
            }
            __syncthreads();
            // ---- pool: the result rows sit in buffer (n_layers & 1)
            const float *res = s_buf + (size_t)((prm.n_layers & 1) ? kMlpThreads : 0) * stride;
            const bool last = ch == chunks - 1;
            if (pooling == DISPU_POOL_WEIGHTED_AVG && !multi) {  // this chunk's share of the group's weight sum
                const int rows = min(kMlpThreads, ns - ch * kMlpThreads);
                if (ch == 0) wtot = 0.f;
                for (int r = 0; r < rows; ++r) wtot = __fadd_rn(wtot, s_wt[r]);
            }
            for (int e = t; e < G * c_out; e += kMlpThreads) {
                const int pg = e / c_out, oc = e - pg * c_out;
                const long long gg = g0 + pg;
                if (gg >= n_groups) continue;
                const int r0 = multi ? pg * ns : 0;
                const int rows = multi ? ns : min(kMlpThreads, ns - ch * kMlpThreads);
                float *o = out + gg * o_stride + oc;
                if (pooling == DISPU_POOL_MAX || pooling == DISPU_POOL_MIN || pooling == DISPU_POOL_MAX_AND_AVG) {
                    // 'min' is max_pool2d(-1 * x) in the reference (pointnet_util.py:135-137): the result is -min(x)
                    const float sg = pooling == DISPU_POOL_MIN ? -1.f : 1.f;
                    float mx = sg * res[(size_t)r0 * stride + oc];
                    for (int r = 1; r < rows; ++r) mx = fmaxf(mx, sg * res[(size_t)(r0 + r) * stride + oc]);
                    if (ch == 0) *o = mx; else *o = fmaxf(*o, mx);
                }
                if (pooling == DISPU_POOL_AVG || pooling == DISPU_POOL_MAX_AND_AVG) {  // avg_pool2d: sum / nsample
                    float *oa = pooling == DISPU_POOL_AVG ? o : o + c_out;  // max_and_avg: [max | avg] (:139-141)
                    float sum = ch == 0 ? 0.f : *oa;
                    for (int r = 0; r < rows; ++r) sum = __fadd_rn(sum, res[(size_t)(r0 + r) * stride + oc]);
                    *oa = last ? __fdiv_rn(sum, (float)ns) : sum;
                }
                if (pooling == DISPU_POOL_WEIGHTED_AVG) {  // sum_r x_r * (w_r / sum_r w_r) (:124-131)
                    if (multi) {
                        float tot = 0.f;
                        for (int r = 0; r < rows; ++r) tot = __fadd_rn(tot, s_wt[r0 + r]);
                        float acc = 0.f;
                        for (int r = 0; r < rows; ++r)
                            acc = __fadd_rn(acc, __fmul_rn(res[(size_t)(r0 + r) * stride + oc], __fdiv_rn(s_wt[r0 + r], tot)));
                        *o = acc;
                    } else {  // chunked group: accumulate the numerator, normalise once the weight sum is complete
                        float acc = ch == 0 ? 0.f : *o;
                        for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, __fmul_rn(res[(size_t)r * stride + oc], s_wt[r]));
                        *o = last ? __fdiv_rn(acc, wtot) : acc;
                    }
                }
            }
            __syncthreads();
        }
    }
    (void)c_in;
}

}  // namespace dispu

using namespace dispu;

static int sa_mlp_impl(int b, int n, int m, int nsample, int c_pts, int layout, int use_xyz, int pooling, const float *xyz,
                       const float *points, const float *new_xyz, const int *idx, int n_layers, const int *widths,
                       const float *const *weights, const float *const *biases, const float *const *bn_scale,
                       const float *const *bn_shift, float *out, int precision, dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample > 0 && c_pts >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(layout == 0 || layout == 1, DISPU_EINVAL);
    DISPU_REQUIRE(pooling >= DISPU_POOL_MAX && pooling <= DISPU_POOL_MAX_AND_AVG, DISPU_EINVAL);
    DISPU_REQUIRE(use_xyz || c_pts > 0, DISPU_EINVAL);
    DISPU_REQUIRE(n_layers >= 1 && n_layers <= kMlpMaxLayers, DISPU_EINVAL);
    DISPU_REQUIRE(widths && weights && biases, DISPU_ENULL);
    DISPU_REQUIRE(widths[0] == (use_xyz ? 3 : 0) + c_pts, DISPU_EINVAL);
    DISPU_REQUIRE(precision == DISPU_MLP_FP32 || precision == DISPU_MLP_TF32X3 || precision == DISPU_MLP_TF32,
                  DISPU_EINVAL);
    if (b == 0 || m == 0) return DISPU_OK;
    DISPU_REQUIRE(xyz && new_xyz && idx && out && (c_pts == 0 || points), DISPU_ENULL);
    MlpParams prm;
    prm.n_layers = n_layers;
    int wmax = widths[0];
    bool any_bn = false;
    for (int l = 0; l <= n_layers; ++l) {
        DISPU_REQUIRE(widths[l] > 0, DISPU_EINVAL);
        prm.widths[l] = widths[l];
        if (widths[l] > wmax) wmax = widths[l];
    }
    for (int l = 0; l < kMlpMaxLayers; ++l) { prm.w[l] = prm.bias[l] = prm.bn_scale[l] = prm.bn_shift[l] = nullptr; }
    for (int l = 0; l < n_layers; ++l) {
        DISPU_REQUIRE(weights[l], DISPU_ENULL);
        prm.w[l] = weights[l];
        prm.bias[l] = biases[l];
        if (bn_scale && bn_scale[l]) {
            DISPU_REQUIRE(bn_shift && bn_shift[l], DISPU_ENULL);
            prm.bn_scale[l] = bn_scale[l];
            prm.bn_shift[l] = bn_shift[l];
            any_bn = true;
        }
    }
    if (precision != DISPU_MLP_FP32) {
        // the tensor-core kernel is built for the hot-path form (max pooling, xyz channels, no BN)
        DISPU_REQUIRE(pooling == DISPU_POOL_MAX && use_xyz && !any_bn, DISPU_EUNSUPPORTED);
        bool handled = false;
        int rc = sa_mlp_tc_launch(b, n, m, nsample, c_pts, layout, xyz, points, new_xyz, idx, prm, out, precision, stream,
                                  &handled);
        if (handled) return rc;
        return DISPU_EUNSUPPORTED;  // shape outside the tensor-core kernel's instantiations
    }
    const int stride = wmax | 1;
    const size_t smem = sizeof(float) * 2 * kMlpThreads * (size_t)stride;
    DISPU_REQUIRE(smem <= 200 * 1024, DISPU_EUNSUPPORTED);
    cudaError_t e = cudaFuncSetAttribute(sa_mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long n_groups = (long long)b * m;
    const bool multi = nsample <= kMlpThreads && (kMlpThreads % nsample) == 0;
    const int G = multi ? kMlpThreads / nsample : 1;
    long long grid = ceil_div64(n_groups, G);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    count_launch();
    sa_mlp_fp32_kernel<<<(unsigned)grid, kMlpThreads, smem, stream>>>(n, m, nsample, c_pts, layout, use_xyz, pooling, xyz,
                                                                     points, new_xyz, idx, prm, out, n_groups, stride);
    return finish_launch();
}

extern "C" int dispu_sa_mlp(int b, int n, int m, int nsample, int c_pts, int layout, const float *xyz,
                            const float *points, const float *new_xyz, const int *idx, int n_layers,
                            const int *widths, const float *const *weights, const float *const *biases, float *out,
                            int precision, dispu_stream_t stream) {
    return sa_mlp_impl(b, n, m, nsample, c_pts, layout, 1, DISPU_POOL_MAX, xyz, points, new_xyz, idx, n_layers, widths, weights,
                       biases, nullptr, nullptr, out, precision, stream);
}

extern "C" int dispu_sa_mlp_ex(int b, int n, int m, int nsample, int c_pts, int layout, int use_xyz, int pooling,
                               const float *xyz, const float *points, const float *new_xyz, const int *idx, int n_layers,
                               const int *widths, const float *const *weights, const float *const *biases,
                               const float *const *bn_scale, const float *const *bn_shift, float *out, int precision,
                               dispu_stream_t stream) {
    return sa_mlp_impl(b, n, m, nsample, c_pts, layout, use_xyz, pooling, xyz, points, new_xyz, idx, n_layers, widths, weights,
                       biases, bn_scale, bn_shift, out, precision, stream);
}

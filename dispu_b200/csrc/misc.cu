// Ops registered for API completeness (no live caller in the reference):
// SelectionSort (tf_ops/grouping/tf_grouping_g.cu:83-123) and ProbSample
// (tf_ops/sampling/tf_sampling_g.cu:7-104).  Correct and coalesced, not tuned.
#include "common.cuh"

namespace dispu {

// One warp per (b,m) row: copy dist -> out / iota -> outi, then k rounds of "find the first
// minimum of out[s..n) and swap it to position s" — the reference's partial selection sort with
// strict '<' (first minimum wins), executed by 32 lanes instead of one thread.
__global__ void __launch_bounds__(256)
selection_sort_kernel(long long rows, int n, int k, const float *__restrict__ dist, int *__restrict__ outi,
                      float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *src = dist + row * n;
    float *pd = out + row * n;
    int *pi = outi + row * n;
    for (int s = lane; s < n; s += 32) {
        pd[s] = src[s];
        pi[s] = s;
    }
    __syncwarp();
    const int kk = k < n ? k : n;
    for (int s = 0; s < kk; ++s) {
        float best = __int_as_float(0x7f800000);
        int bt = 0x7fffffff;
        for (int t = s + lane; t < n; t += 32) {
            const float v = pd[t];
            if (bt == 0x7fffffff || v < best) {  // first element seen initialises (handles inf rows)
                best = v;
                bt = t;
            }
        }
        // arg-min with lowest position on ties; position s itself wins ties (reference starts min=s)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(kFull, best, o);
            const int ot = __shfl_xor_sync(kFull, bt, o);
            if (ot != 0x7fffffff && (bt == 0x7fffffff || ob < best || (ob == best && ot < bt))) {
                best = ob;
                bt = ot;
            }
        }
        if (lane == 0 && bt != s && bt != 0x7fffffff) {
            const float tf = pd[bt]; pd[bt] = pd[s]; pd[s] = tf;
            const int ti = pi[bt]; pi[bt] = pi[s]; pi[s] = ti;
        }
        __syncwarp();
    }
}

// Inclusive prefix sum per cloud (one CTA per cloud; per-thread contiguous segment in double,
// fixed-order combine) then lower-bound search of r*total, falling back to n-1 like the
// reference's binarysearchKernel (tf_sampling_g.cu:90-104).
__global__ void __launch_bounds__(256)
cumsum_kernel(int n, const float *__restrict__ inp, float *__restrict__ out) {
    __shared__ double s_part[256];
    const int cloud = blockIdx.x, t = threadIdx.x;
    const float *src = inp + (size_t)cloud * n;
    float *dst = out + (size_t)cloud * n;
    const int seg = ceil_div(n, 256);
    const int lo = min(n, t * seg), hi = min(n, lo + seg);
    double acc = 0.0;
    for (int i = lo; i < hi; ++i) acc += (double)src[i];
    s_part[t] = acc;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < t; ++w) base += s_part[w];
    acc = base;
    for (int i = lo; i < hi; ++i) {
        acc += (double)src[i];
        dst[i] = (float)acc;
    }
}

__global__ void __launch_bounds__(256)
binary_search_kernel(int n, int m, const float *__restrict__ cum, const float *__restrict__ query,
                     int *__restrict__ result) {
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float *c = cum + (size_t)cloud * n;
    const float q = __fmul_rn(query[(size_t)cloud * m + j], c[n - 1]);
    int lo = 0, hi = n - 1;  // smallest index with c[idx] >= q, else n-1
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] >= q) hi = mid; else lo = mid + 1;
    }
    result[(size_t)cloud * m + j] = lo;
}

}  // namespace dispu

using namespace dispu;

extern "C" int dispu_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out,
                                    dispu_stream_t stream) {
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    DISPU_REQUIRE(k > 0, DISPU_EINVAL);  // tf_grouping.cpp attr k>0
    const long long rows = (long long)b * m;
    if (rows == 0) return DISPU_OK;
    DISPU_REQUIRE(dist && outi && out, DISPU_ENULL);
    count_launch();
    selection_sort_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(rows, n, k, dist, outi, out);
    return finish_launch();
}

extern "C" int dispu_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out,
                                 dispu_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DISPU_REQUIRE(b >= 0 && n > 0 && m >= 0, DISPU_EINVAL);
    if (b == 0) return DISPU_OK;
    DISPU_REQUIRE(inp_p && temp && (m == 0 || (inp_r && out)), DISPU_ENULL);
    DISPU_REQUIRE(b <= 65535, DISPU_EUNSUPPORTED);
    count_launch();
    cumsum_kernel<<<b, 256, 0, stream>>>(n, inp_p, temp);
    if (m > 0) {
        count_launch();
        binary_search_kernel<<<dim3(ceil_div(m, 256), b), 256, 0, stream>>>(n, m, temp, inp_r, out);
    }
    return finish_launch();
}

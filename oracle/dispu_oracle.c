/*
 * dispu_oracle.c — CPU restatement of the Dis-PU sampling / grouping / interpolation /
 * Chamfer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (dispu_b200/) never does and fails loudly when its
 * CUDA library is missing.
 *
 * Every function follows the reference semantics cited next to it (paths relative to the
 * reference checkout).  It is written from those semantics, not copied: one restatement per
 * op, plain C99, single thread, 64-bit indexing.
 *
 * Two distance contractions exist in the reference and both are reproduced bit-exactly:
 *   ORC_GPU : the CUDA kernels as nvcc (default -fmad=true) compiles them for sm_100a:
 *             d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy))     (checked in SASS: FMUL y, FFMA x, FFMA z)
 *   ORC_CPU : the host loops as g++ -O2 on x86-64 compiles them (no FMA contraction):
 *             d2 = ((dx*dx) + (dy*dy)) + (dz*dz)
 * Build with -ffp-contract=off so the C compiler adds no contraction of its own.
 *
 * Pinning status: the reference ships NO golden vectors for the forward results of this path
 * (SURVEY.md §4/§8c).  This restatement is pinned instead against the reference's own code
 * compiled from /root/reference (oracle/build_ref.sh -> oracle/_ref/): the host loops
 * (nnsearch, threenn_cpu, threeinterpolate_cpu, query_ball_point_cpu, group_point_cpu,
 * nanoflann kNN) in tests/test_oracle.py on CPU, and the recompiled CUDA kernels
 * (FPS, ball query, group, NnDistance) in the -m gpu tests; fixtures generated that way are
 * committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_GPU 0
#define ORC_CPU 1

static inline float d2_gpu(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}
static inline float d2_cpu(float dx, float dy, float dz) {
    float a = dx * dx;
    float b = dy * dy;
    float c = dz * dz;
    float s = a + b;
    return s + c;
}
static inline float d2_mode(int mode, float dx, float dy, float dz) {
    return mode == ORC_GPU ? d2_gpu(dx, dy, dz) : d2_cpu(dx, dy, dz);
}

/* ---------------------------------------------------------------------------------------
 * Farthest point sampling — tf_ops/sampling/tf_sampling_g.cu:105-170 (kernel), :203 (launch
 * <<<32,512>>>).  idx[0]=0; running min-distance starts at 1e38f; each step updates
 * mind[k]=min(d2(p[k],p[last]),mind[k]) (:142-143) and picks the arg-max.  Tie order of the
 * reference's strided scan (:130,:146 strict '>') plus its tree reduction (:153-163, strict
 * '<' keeps the lower thread id): among equal maxima the smallest (k mod 512) wins, then the
 * smallest k.  Works for m>n (all mind reach 0 -> index 0 repeats).
 * ------------------------------------------------------------------------------------- */
int orc_fps(int b, int n, int m, const float *xyz, int *idx) {
    if (b < 0 || n <= 0 || m < 0) return -1;
    if (m == 0) return 0; /* :106 early return */
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    if (!mind) return -2;
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        int *out = idx + (size_t)i * m;
        for (int k = 0; k < n; ++k) mind[k] = 1e38f;
        int last = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = p[last * 3 + 0], y1 = p[last * 3 + 1], z1 = p[last * 3 + 2];
            float best = -1.0f;
            int besti = 0;
            unsigned bestlane = 0;
            for (int k = 0; k < n; ++k) {
                float d = d2_gpu(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
                float d2 = fminf(d, mind[k]);
                mind[k] = d2;
                unsigned lane = (unsigned)k & 511u;
                /* (value desc, k mod 512 asc, k asc); k ascends so the last clause is implicit */
                if (d2 > best || (d2 == best && lane < bestlane)) {
                    best = d2;
                    besti = k;
                    bestlane = lane;
                }
            }
            last = besti;
            out[j] = last;
        }
    }
    free(mind);
    return 0;
}

/* GatherPoint — tf_sampling_g.cu:172-181: out[b,j,:] = inp[b,idx[b,j],:] (3 channels). */
int orc_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l)
                out[((size_t)i * m + j) * 3 + l] = inp[((size_t)i * n + a) * 3 + l];
        }
    return 0;
}

/* GatherPointGrad — tf_sampling_g.cu:183-192 after the zero fill at tf_sampling.cpp:174.
 * Sequential accumulation in (i,j) order; the GPU's atomic order is unspecified. */
int orc_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l)
                inp_g[((size_t)i * n + a) * 3 + l] += out_g[((size_t)i * m + j) * 3 + l];
        }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * QueryBallPoint — tf_ops/grouping/tf_grouping_g.cu:3-36 (GPU, ORC_GPU contraction) and
 * tf_ops/grouping/query_ball_point.cpp:19-47 (host loop, ORC_CPU).  Ascending scan, hit iff
 * max(sqrtf(d2),1e-20f) < radius[0] (only element 0 is read, :25), first hit fills the whole
 * row (:26-29), stop at nsample (:15-16).  Rows without any hit are left untouched by the
 * reference; `fill_empty` selects the product's documented choice (zero row) instead.
 * pts_cnt may be NULL (the host loop has none).
 * ------------------------------------------------------------------------------------- */
int orc_query_ball_point(int mode, int b, int n, int m, const float *radius, int nsample,
                         const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                         int fill_empty) {
    const float r = radius[0];
    for (int i = 0; i < b; ++i) {
        const float *p = xyz1 + (size_t)i * n * 3;
        const float *q = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < m; ++j) {
            int *row = idx + ((size_t)i * m + j) * nsample;
            int cnt = 0;
            float x2 = q[j * 3 + 0], y2 = q[j * 3 + 1], z2 = q[j * 3 + 2];
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float s = d2_mode(mode, x2 - p[k * 3 + 0], y2 - p[k * 3 + 1], z2 - p[k * 3 + 2]);
                float d = fmaxf(sqrtf(s), 1e-20f);
                if (d < r) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt++] = k;
                }
            }
            if (cnt == 0 && fill_empty)
                for (int l = 0; l < nsample; ++l) row[l] = 0;
            if (pts_cnt) pts_cnt[(size_t)i * m + j] = cnt;
        }
    }
    return 0;
}

/* GroupPoint — tf_grouping_g.cu:40-57 / query_ball_point.cpp:52-66. */
int orc_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                    float *out) {
    for (int i = 0; i < b; ++i)
        for (size_t r = 0; r < (size_t)m * nsample; ++r) {
            int ii = idx[(size_t)i * m * nsample + r];
            memcpy(out + ((size_t)i * m * nsample + r) * c, points + ((size_t)i * n + ii) * c,
                   sizeof(float) * (size_t)c);
        }
    return 0;
}

/* GroupPointGrad — tf_grouping_g.cu:61-78 after the zero fill at tf_grouping.cpp:208;
 * host loop query_ball_point.cpp:70-84 (which does not zero). */
int orc_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out,
                         const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
    for (int i = 0; i < b; ++i)
        for (size_t r = 0; r < (size_t)m * nsample; ++r) {
            int ii = idx[(size_t)i * m * nsample + r];
            for (int l = 0; l < c; ++l)
                grad_points[((size_t)i * n + ii) * c + l] += grad_out[((size_t)i * m * nsample + r) * c + l];
        }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * kNN grouping (3-D or c-D, squared L2).  Three reference call sites share this definition:
 *   - tf_ops/grouping/tf_grouping.py:116-141 knn_point: top_k(-sum_c (xyz1-xyz2)^2), ties go
 *     to the lower index (tf.nn.top_k), returned value is the NEGATED distance;
 *   - libs/nearest_neighbors/knn_.cxx:104-135 + nanoflann.hpp:323-348 (L2_Adaptor): the
 *     distance is accumulated dimension by dimension in float, diff = query - data, no FMA;
 *   - Common/ops.py:110-118,154-179 (grouping/knn_query) which casts the indices to int32.
 * d = ((dx*dx + dy*dy) + dz*dz ...) left to right; result ascending in d, stable in index.
 * dist receives +d (callers negate for knn_point).  dist may be NULL.
 * ------------------------------------------------------------------------------------- */
int orc_knn(int b, int n, int m, int c, int k, const float *data, const float *query, float *dist,
            int *idx) {
    if (k > n || k <= 0) return -1;
    float *bd = (float *)malloc(sizeof(float) * (size_t)k);
    int *bi = (int *)malloc(sizeof(int) * (size_t)k);
    if (!bd || !bi) return -2;
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *q = query + ((size_t)i * m + j) * c;
            int cnt = 0;
            for (int p = 0; p < n; ++p) {
                const float *x = data + ((size_t)i * n + p) * c;
                float d = 0.0f;
                for (int l = 0; l < c; ++l) {
                    float diff = q[l] - x[l];
                    float sq = diff * diff;
                    d = (l == 0) ? sq : d + sq;
                }
                if (cnt < k || d < bd[cnt - 1]) {
                    int pos = cnt < k ? cnt : k - 1;
                    while (pos > 0 && bd[pos - 1] > d) { /* strict: equal keeps earlier index first */
                        bd[pos] = bd[pos - 1];
                        bi[pos] = bi[pos - 1];
                        --pos;
                    }
                    bd[pos] = d;
                    bi[pos] = p;
                    if (cnt < k) ++cnt;
                }
            }
            for (int s = 0; s < k; ++s) {
                if (dist) dist[((size_t)i * m + j) * k + s] = bd[s];
                idx[((size_t)i * m + j) * k + s] = bi[s];
            }
        }
    free(bd);
    free(bi);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * ThreeNN — tf_ops/interpolation/tf_interpolate.cpp:60-103.  Squared distances in float
 * (ORC_CPU contraction, (x2-x1) with x2 = known point), compared as double against an initial
 * 1e40; strict '<' cascades so the first index wins at every rank; stores float(best) which
 * is +inf when fewer than 3 known points exist, with index 0.
 * ------------------------------------------------------------------------------------- */
int orc_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float *u = xyz1 + ((size_t)i * n + j) * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                const float *v = xyz2 + ((size_t)i * m + k) * 3;
                double d = (double)d2_cpu(v[0] - u[0], v[1] - u[1], v[2] - u[2]);
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else if (d < b3) {
                    b3 = d; i3 = k;
                }
            }
            float *od = dist + ((size_t)i * n + j) * 3;
            int *oi = idx + ((size_t)i * n + j) * 3;
            od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
    return 0;
}

/* ThreeInterpolate — tf_interpolate.cpp:107-127: out = (p1*w1 + p2*w2) + p3*w3, no FMA. */
int orc_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                          const float *weight, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const int *id = idx + ((size_t)i * n + j) * 3;
            const float *w = weight + ((size_t)i * n + j) * 3;
            const float *p = points + (size_t)i * m * c;
            for (int l = 0; l < c; ++l) {
                float t1 = p[(size_t)id[0] * c + l] * w[0];
                float t2 = p[(size_t)id[1] * c + l] * w[1];
                float t3 = p[(size_t)id[2] * c + l] * w[2];
                float s = t1 + t2;
                out[((size_t)i * n + j) * c + l] = s + t3;
            }
        }
    return 0;
}

/* ThreeInterpolateGrad — tf_interpolate.cpp:131-153 after the zero fill at :258. */
int orc_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const int *id = idx + ((size_t)i * n + j) * 3;
            const float *w = weight + ((size_t)i * n + j) * 3;
            float *g = grad_points + (size_t)i * m * c;
            for (int l = 0; l < c; ++l) {
                float go = grad_out[((size_t)i * n + j) * c + l];
                g[(size_t)id[0] * c + l] += go * w[0];
                g[(size_t)id[1] * c + l] += go * w[1];
                g[(size_t)id[2] * c + l] += go * w[2];
            }
        }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * NnDistance — tf_ops/nn_distance/tf_nndistance_g.cu:5-127 (ORC_GPU; strict '<' inside a
 * 512-point tile :29-62, strict '>' across tiles :119 => the first minimum wins) and
 * tf_ops/nn_distance/tf_nndistance.cpp:21-43 nnsearch (ORC_CPU; compared in double, stored as
 * float).  Difference is xyz2 - xyz1.  Produces both directions like the launcher (:128-131).
 * ------------------------------------------------------------------------------------- */
static void nn_one_direction(int mode, int b, int n, int m, const float *xyz1, const float *xyz2,
                             float *dist, int *idx) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float *u = xyz1 + ((size_t)i * n + j) * 3;
            float best = 0.0f;
            int bi = 0;
            for (int k = 0; k < m; ++k) {
                const float *v = xyz2 + ((size_t)i * m + k) * 3;
                float d = d2_mode(mode, v[0] - u[0], v[1] - u[1], v[2] - u[2]);
                if (k == 0 || d < best) {
                    best = d;
                    bi = k;
                }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = bi;
        }
}
int orc_nn_distance(int mode, int b, int n, int m, const float *xyz1, const float *xyz2,
                    float *dist1, int *idx1, float *dist2, int *idx2) {
    if (n <= 0 || m <= 0) return -1;
    nn_one_direction(mode, b, n, m, xyz1, xyz2, dist1, idx1);
    nn_one_direction(mode, b, m, n, xyz2, xyz1, dist2, idx2);
    return 0;
}

/* NnDistanceGrad — tf_nndistance_g.cu:132-157 / tf_nndistance.cpp:126-163: g = 2*grad_dist;
 * grad_xyz1[i] += g*(x1-x2[idx1[i]]); grad_xyz2[idx1[i]] -= the same; then the mirror. */
int orc_nn_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2,
                         const float *grad_dist1, const int *idx1, const float *grad_dist2,
                         const int *idx2, float *grad_xyz1, float *grad_xyz2) {
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    for (int i = 0; i < b; ++i) {
        for (int dir = 0; dir < 2; ++dir) {
            int na = dir == 0 ? n : m, nb = dir == 0 ? m : n;
            const float *a = dir == 0 ? xyz1 + (size_t)i * n * 3 : xyz2 + (size_t)i * m * 3;
            const float *bb = dir == 0 ? xyz2 + (size_t)i * m * 3 : xyz1 + (size_t)i * n * 3;
            const float *gd = dir == 0 ? grad_dist1 + (size_t)i * n : grad_dist2 + (size_t)i * m;
            const int *id = dir == 0 ? idx1 + (size_t)i * n : idx2 + (size_t)i * m;
            float *ga = dir == 0 ? grad_xyz1 + (size_t)i * n * 3 : grad_xyz2 + (size_t)i * m * 3;
            float *gb = dir == 0 ? grad_xyz2 + (size_t)i * m * 3 : grad_xyz1 + (size_t)i * n * 3;
            (void)nb;
            for (int j = 0; j < na; ++j) {
                int j2 = id[j];
                float g = gd[j] * 2.0f;
                for (int l = 0; l < 3; ++l) {
                    float t = g * (a[j * 3 + l] - bb[j2 * 3 + l]);
                    ga[j * 3 + l] += t;
                    gb[j2 * 3 + l] -= t;
                }
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * chamfer / hausdorff_loss — Common/loss_utils.py:45-64 and :67-84, called as
 * chamfer(pred, gt) -> nn_distance(gt, pred): dists_forward belongs to gt.
 *   cd = mean_b((mean_i fwd + mean_j bwd) / radius_b)
 *   hd = max_b ((max_i  fwd + max_j  bwd) / radius_b)
 * Accumulated in double here (TF's float reduction order is unspecified); per-cloud values
 * are returned as float in per_cd / per_hd (either may be NULL).
 * ------------------------------------------------------------------------------------- */
int orc_chamfer(int mode, int b, int n_gt, int n_pred, const float *gt, const float *pred,
                const float *radius, float *cd, float *hd, float *per_cd, float *per_hd) {
    size_t mx = (size_t)(n_gt > n_pred ? n_gt : n_pred);
    float *d = (float *)malloc(sizeof(float) * mx);
    int *di = (int *)malloc(sizeof(int) * mx);
    if (!d || !di) return -2;
    double cd_acc = 0.0, hd_max = -INFINITY;
    for (int i = 0; i < b; ++i) {
        const float *g = gt + (size_t)i * n_gt * 3;
        const float *p = pred + (size_t)i * n_pred * 3;
        double sf = 0, sb = 0, mf = -INFINITY, mb = -INFINITY;
        nn_one_direction(mode, 1, n_gt, n_pred, g, p, d, di);
        for (int j = 0; j < n_gt; ++j) { sf += d[j]; if (d[j] > mf) mf = d[j]; }
        nn_one_direction(mode, 1, n_pred, n_gt, p, g, d, di);
        for (int j = 0; j < n_pred; ++j) { sb += d[j]; if (d[j] > mb) mb = d[j]; }
        double r = radius ? (double)radius[i] : 1.0;
        double c1 = (sf / n_gt + sb / n_pred) / r;
        double h1 = (mf + mb) / r;
        if (per_cd) per_cd[i] = (float)c1;
        if (per_hd) per_hd[i] = (float)h1;
        cd_acc += c1;
        if (h1 > hd_max) hd_max = h1;
    }
    if (cd) *cd = (float)(cd_acc / b);
    if (hd) *hd = (float)hd_max;
    free(d);
    free(di);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Grouped MLP + pool — Common/pointnet_util.py:117-133 (pointnet_sa_module) on the tensor
 * built by sample_and_group (:41-59): per group row r = (b, j, s):
 *     X = concat(xyz[idx]-new_xyz[j], points[idx])      (xyz first, :53)   layout 0
 *     X = concat(points[idx], xyz[idx]-new_xyz[j])      (_msg, :179)       layout 1
 * then for each layer  X = relu(X.W + bias)  (Common/tf_util.py:120-185, 1x1 conv, NHWC,
 * W is (C_in, C_out) row-major = the (1,1,C_in,C_out) kernel), then max over the nsample
 * rows of each group.  FP32 with sequential float accumulation over C_in.
 * points may be NULL (c_pts = 0).  last layer keeps its ReLU (activation_fn default).
 * ------------------------------------------------------------------------------------- */
int orc_sa_mlp(int b, int n, int m, int nsample, int c_pts, int layout, const float *xyz,
               const float *points, const float *new_xyz, const int *idx, int n_layers,
               const int *widths /* n_layers+1: C_in, C1, ... */, const float *const *W,
               const float *const *bias, float *out /* (b,m,C_last) */) {
    int c_in = 3 + c_pts;
    if (widths[0] != c_in) return -1;
    int wmax = 0;
    for (int l = 0; l <= n_layers; ++l) if (widths[l] > wmax) wmax = widths[l];
    float *x = (float *)malloc(sizeof(float) * (size_t)wmax);
    float *y = (float *)malloc(sizeof(float) * (size_t)wmax);
    if (!x || !y) return -2;
    int c_out = widths[n_layers];
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            float *o = out + ((size_t)i * m + j) * c_out;
            const float *ctr = new_xyz + ((size_t)i * m + j) * 3;
            for (int s = 0; s < nsample; ++s) {
                int ii = idx[((size_t)i * m + j) * nsample + s];
                const float *pxyz = xyz + ((size_t)i * n + ii) * 3;
                float g[3] = {pxyz[0] - ctr[0], pxyz[1] - ctr[1], pxyz[2] - ctr[2]};
                const float *pf = c_pts ? points + ((size_t)i * n + ii) * c_pts : NULL;
                if (layout == 0) {
                    x[0] = g[0]; x[1] = g[1]; x[2] = g[2];
                    for (int l = 0; l < c_pts; ++l) x[3 + l] = pf[l];
                } else {
                    for (int l = 0; l < c_pts; ++l) x[l] = pf[l];
                    x[c_pts] = g[0]; x[c_pts + 1] = g[1]; x[c_pts + 2] = g[2];
                }
                for (int l = 0; l < n_layers; ++l) {
                    int ci = widths[l], co = widths[l + 1];
                    for (int oc = 0; oc < co; ++oc) {
                        float acc = 0.0f;
                        for (int ic = 0; ic < ci; ++ic) acc += x[ic] * W[l][(size_t)ic * co + oc];
                        acc += bias[l] ? bias[l][oc] : 0.0f;
                        y[oc] = acc > 0.0f ? acc : 0.0f;
                    }
                    float *t = x; x = y; y = t;
                }
                for (int oc = 0; oc < c_out; ++oc)
                    o[oc] = (s == 0 || x[oc] > o[oc]) ? x[oc] : o[oc];
            }
        }
    free(x);
    free(y);
    return 0;
}

/* get_repulsion_loss — Common/loss_utils.py:271-298 with the live defaults (ball query, squared L2):
 * per point the squared distances to its nsample grouped points (idx row, pad repeats included),
 * top_k(-d2,5)[1:] = the 2nd..5th smallest, loss = mean over (b,n,4) of max(0, h - d2). */
int orc_repulsion_loss(int b, int n, int nsample, float h, const float *pred, const int *idx, float *loss) {
    double acc = 0.0;
    float *d = (float *)malloc(sizeof(float) * (size_t)nsample);
    if (!d) return -2;
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float *c = pred + ((size_t)i * n + j) * 3;
            for (int s = 0; s < nsample; ++s) {
                const float *q = pred + ((size_t)i * n + idx[((size_t)i * n + j) * nsample + s]) * 3;
                d[s] = d2_cpu(q[0] - c[0], q[1] - c[1], q[2] - c[2]);
            }
            for (int a = 0; a < 5 && a < nsample; ++a) { /* partial selection sort of the 5 smallest */
                int mn = a;
                for (int t = a + 1; t < nsample; ++t) if (d[t] < d[mn]) mn = t;
                float tf = d[a]; d[a] = d[mn]; d[mn] = tf;
            }
            for (int a = 1; a < 5 && a < nsample; ++a) acc += (double)fmaxf(0.0f, h - d[a]);
        }
    free(d);
    *loss = (float)(acc / ((double)b * n * 4));
    return 0;
}

/* SelectionSort — tf_grouping_g.cu:83-123 / selection_sort.cpp:20-63: copy dist, then a
 * partial selection sort of the first k entries per row (strict '<' => first minimum). */
int orc_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
    for (size_t r = 0; r < (size_t)b * m; ++r) {
        float *pd = out + r * n;
        int *pi = outi + r * n;
        for (int s = 0; s < n; ++s) { pd[s] = dist[r * n + s]; pi[s] = s; }
        for (int s = 0; s < k; ++s) {
            int mn = s;
            for (int t = s + 1; t < n; ++t) if (pd[t] < pd[mn]) mn = t;
            if (mn != s) {
                float tf = pd[mn]; pd[mn] = pd[s]; pd[s] = tf;
                int ti = pi[mn]; pi[mn] = pi[s]; pi[s] = ti;
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Feature-space kNN — knn_point_2 (tf_ops/grouping/tf_grouping.py:95-114) on the matmul-form
 * distance matrix batch_distance_matrix_general (:61-66):
 *     D[q][p] = (r_q - 2*dot(q,p)) + r_p,   r_x = sum_c x_c^2
 * then tf.nn.top_k(-D, k): ascending D, ties to the lower index.  The `unique` flag is a no-op
 * (prepare_for_unique_top_k :89-91 rebinds a local).  Call site: get_edge_feature
 * (Common/ops.py:1866-1868: k+1 neighbours of the cloud in itself, column 0 dropped).
 * TensorFlow leaves the accumulation order of reduce_sum / matmul unspecified; this restatement
 * (and the CUDA kernel, bit for bit) uses sequential fmaf over the channels for both.
 * ld* = row strides in floats.  dist may be NULL.
 * ------------------------------------------------------------------------------------- */
int orc_feature_knn(int b, int n, int m, int c, int k, const float *points, int ldp, const float *queries,
                    int ldq, float *dist, int *idx) {
    if (k <= 0 || k > n || c <= 0) return -1;
    float *bd = (float *)malloc(sizeof(float) * (size_t)k);
    int *bi = (int *)malloc(sizeof(int) * (size_t)k);
    float *rp = (float *)malloc(sizeof(float) * (size_t)n);
    if (!bd || !bi || !rp) return -2;
    for (int i = 0; i < b; ++i) {
        for (int p = 0; p < n; ++p) {
            const float *x = points + ((size_t)i * n + p) * ldp;
            float r = 0.0f;
            for (int l = 0; l < c; ++l) r = fmaf(x[l], x[l], r);
            rp[p] = r;
        }
        for (int j = 0; j < m; ++j) {
            const float *q = queries + ((size_t)i * m + j) * ldq;
            float rq = 0.0f;
            for (int l = 0; l < c; ++l) rq = fmaf(q[l], q[l], rq);
            int cnt = 0;
            for (int p = 0; p < n; ++p) {
                const float *x = points + ((size_t)i * n + p) * ldp;
                float dot = 0.0f;
                for (int l = 0; l < c; ++l) dot = fmaf(q[l], x[l], dot);
                float d = fmaf(-2.0f, dot, rq) + rp[p];
                if (cnt < k || d < bd[cnt - 1]) {
                    int pos = cnt < k ? cnt : k - 1;
                    while (pos > 0 && bd[pos - 1] > d) {
                        bd[pos] = bd[pos - 1];
                        bi[pos] = bi[pos - 1];
                        --pos;
                    }
                    bd[pos] = d;
                    bi[pos] = p;
                    if (cnt < k) ++cnt;
                }
            }
            for (int s = 0; s < k; ++s) {
                if (dist) dist[((size_t)i * m + j) * k + s] = bd[s];
                idx[((size_t)i * m + j) * k + s] = bi[s];
            }
        }
    }
    free(bd); free(bi); free(rp);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * EdgeConv dense block — dense_conv (Common/ops.py:1897-1915) on the edge feature of
 * get_edge_feature (:1856-1877), n=3 layers, growth g, conv2d = 1x1 conv + bias (+ReLU)
 * (Common/tf_util.py:120-185), literal op order (nothing hoisted):
 *     e   = [f_i, f_j - f_i]                       (2c)     j = idx[i][s]
 *     y0  = relu(e.W0 + b0)          z1 = [y0, f_i]         (g + c)
 *     y1  = relu(z1.W1 + b1)         z2 = [y1, y0, f_i]     (2g + c)
 *     y2  =      z2.W2 + b2          z3 = [y2, y1, y0, f_i] (3g + c)
 *     out[i] = max_s z3
 * Accumulation in double, rounded to float per activation: the comparator for a 1e-5 bound, not a
 * bit-exact twin (the CUDA kernel hoists the f_i terms to per-point products).
 * feat (b,n,ldf) ; idx (b,n,k) ; out (b,n,ldo) receives 3g+c columns.
 * ------------------------------------------------------------------------------------- */
int orc_dense_conv(int b, int n, int c, int k, int g, const float *feat, int ldf, const int *idx,
                   const float *w0, const float *b0, const float *w1, const float *b1, const float *w2,
                   const float *b2, float *out, int ldo) {
    if (c <= 0 || g <= 0 || k <= 0) return -1;
    const int wz = 3 * g + c;
    float *z = (float *)malloc(sizeof(float) * (size_t)wz);
    float *e = (float *)malloc(sizeof(float) * (size_t)(2 * c));
    if (!z || !e) return -2;
    for (int i = 0; i < b; ++i)
        for (int p = 0; p < n; ++p) {
            const float *fi = feat + ((size_t)i * n + p) * ldf;
            float *o = out + ((size_t)i * n + p) * ldo;
            for (int s = 0; s < k; ++s) {
                const int j = idx[((size_t)i * n + p) * k + s];
                const float *fj = feat + ((size_t)i * n + j) * ldf;
                for (int l = 0; l < c; ++l) { e[l] = fi[l]; e[c + l] = fj[l] - fi[l]; }
                float *y2 = z, *y1 = z + g, *y0 = z + 2 * g, *zf = z + 3 * g;
                for (int l = 0; l < c; ++l) zf[l] = fi[l];
                for (int oc = 0; oc < g; ++oc) {
                    double acc = 0.0;
                    for (int ic = 0; ic < 2 * c; ++ic) acc += (double)e[ic] * w0[(size_t)ic * g + oc];
                    acc += b0 ? b0[oc] : 0.0;
                    y0[oc] = acc > 0.0 ? (float)acc : 0.0f;
                }
                for (int oc = 0; oc < g; ++oc) { /* z1 = [y0, f_i] */
                    double acc = 0.0;
                    for (int ic = 0; ic < g; ++ic) acc += (double)y0[ic] * w1[(size_t)ic * g + oc];
                    for (int ic = 0; ic < c; ++ic) acc += (double)fi[ic] * w1[(size_t)(g + ic) * g + oc];
                    acc += b1 ? b1[oc] : 0.0;
                    y1[oc] = acc > 0.0 ? (float)acc : 0.0f;
                }
                for (int oc = 0; oc < g; ++oc) { /* z2 = [y1, y0, f_i], no activation */
                    double acc = 0.0;
                    for (int ic = 0; ic < g; ++ic) acc += (double)y1[ic] * w2[(size_t)ic * g + oc];
                    for (int ic = 0; ic < g; ++ic) acc += (double)y0[ic] * w2[(size_t)(g + ic) * g + oc];
                    for (int ic = 0; ic < c; ++ic) acc += (double)fi[ic] * w2[(size_t)(2 * g + ic) * g + oc];
                    acc += b2 ? b2[oc] : 0.0;
                    y2[oc] = (float)acc;
                }
                for (int l = 0; l < wz; ++l) o[l] = (s == 0 || z[l] > o[l]) ? z[l] : o[l];
            }
        }
    free(z); free(e);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Weighted local aggregation of PointShuffle2 (Common/ops.py:1062-1068) with weight_net_hidden
 * (:181-191, one hidden layer of 16 units, bn=True evaluated with its initial moving statistics:
 * mean 0, variance 1, eps 1e-3 => a scale of 1/sqrt(1+eps)):
 *     gx[s]   = xyz[idx[i][s]] - xyz[i]
 *     wn[s][t] = relu((gx[s].Ww[:,t] + bw[t]) * bn_scale)
 *     H[i][ch][t] = sum_s G[i][s][ch] * wn[s][t]           (= transpose(G) . wn, :1066-1067)
 * out (b,m,C*T) row = (ch, t) flattened, the operand of the [1,C] "after_conv" (:1068-1072).
 * ------------------------------------------------------------------------------------- */
int orc_weightnet_aggregate(int b, int m, int ns, int C, int T, const float *xyz, const int *idx, const float *G,
                            const float *Ww, const float *bw, float bn_scale, float *out) {
    float *wn = (float *)malloc(sizeof(float) * (size_t)ns * T);
    if (!wn) return -2;
    for (int i = 0; i < b; ++i)
        for (int p = 0; p < m; ++p) {
            const float *ctr = xyz + ((size_t)i * m + p) * 3;
            const int *row = idx + ((size_t)i * m + p) * ns;
            for (int s = 0; s < ns; ++s) {
                const float *q = xyz + ((size_t)i * m + row[s]) * 3;
                float gx[3] = {q[0] - ctr[0], q[1] - ctr[1], q[2] - ctr[2]};
                for (int t = 0; t < T; ++t) {
                    double a = (double)gx[0] * Ww[t] + (double)gx[1] * Ww[T + t] + (double)gx[2] * Ww[2 * T + t];
                    a = (a + (bw ? bw[t] : 0.0)) * bn_scale;
                    wn[s * T + t] = a > 0.0 ? (float)a : 0.0f;
                }
            }
            const float *g = G + ((size_t)i * m + p) * ns * C;
            float *o = out + ((size_t)i * m + p) * C * T;
            for (int ch = 0; ch < C; ++ch)
                for (int t = 0; t < T; ++t) {
                    double a = 0.0;
                    for (int s = 0; s < ns; ++s) a += (double)g[(size_t)s * C + ch] * wn[s * T + t];
                    o[ch * T + t] = (float)a;
                }
        }
    free(wn);
    return 0;
}

/* get_uniform_loss, one percentage term — Common/loss_utils.py:238-269 given the ball-query idx
 * (b,npoint,nsample): per group g = pcd[idx]; knn_point(2, g, g) (tf_grouping.py:116-141: unfused squared
 * distances, top_k over ALL nsample entries so the first hit is the point itself and padding repeats count);
 * u = sqrt(|d2 + 1e-8|); mean over (b,npoint,nsample) of (u - e)^2 / (e + 1e-8). */
int orc_uniform_loss_term(int b, int n, int npoint, int nsample, float expect_len, const float *pcd, const int *idx,
                          float *loss) {
    double acc = 0.0;
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < npoint; ++j) {
            const int *row = idx + ((size_t)i * npoint + j) * nsample;
            for (int s = 0; s < nsample; ++s) {
                const float *c = pcd + ((size_t)i * n + row[s]) * 3;
                float d0 = INFINITY, d1 = INFINITY;
                for (int t = 0; t < nsample; ++t) {
                    const float *q = pcd + ((size_t)i * n + row[t]) * 3;
                    float d = d2_cpu(q[0] - c[0], q[1] - c[1], q[2] - c[2]);
                    if (d < d0) { float tf = d0; d0 = d; d = tf; }
                    if (d < d1) d1 = d;
                }
                if (nsample < 2) d1 = 0.0f;
                float u = sqrtf(fabsf(d1 + 1e-8f));
                float tt = u - expect_len;
                acc += (double)((tt * tt) / (expect_len + 1e-8f));
            }
        }
    *loss = (float)(acc / ((double)b * npoint * nsample));
    return 0;
}

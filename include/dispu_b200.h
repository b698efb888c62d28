/*
 * dispu_b200.h — C ABI of the B200-native (sm_100a) Dis-PU sampling / grouping /
 * interpolation / Chamfer hot path (libdispu_b200.so).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *     tensors are dense, row-major, contiguous; float32 / int32;
 *   - the callee never allocates or frees device memory and never synchronises the stream;
 *   - dims are int32 like the reference launchers, indexing inside is 64-bit;
 *   - every entry point returns 0 on success, a NEGATIVE DISPU_E* code for a rejected
 *     argument (the reference wrappers' OP_REQUIRES checks), or a POSITIVE cudaError_t;
 *   - `stream` is a cudaStream_t (0 = legacy default stream, which is what the reference
 *     launchers use: e.g. tf_ops/sampling/tf_sampling_g.cu:203-205);
 *   - there is no CPU fallback anywhere.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  The same object also exports the reference's C++-mangled launcher
 * symbols (csrc/ref_launchers.cu) so the reference's unmodified tf_*.cpp wrappers link
 * against it — see INTEGRATION.md.
 */
#ifndef DISPU_B200_H_
#define DISPU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *dispu_stream_t; /* == cudaStream_t */

#define DISPU_OK 0
#define DISPU_EINVAL (-1)      /* bad dimension / attribute (OP_REQUIRES InvalidArgument) */
#define DISPU_ENULL (-2)       /* required pointer is NULL */
#define DISPU_EUNSUPPORTED (-3) /* size outside what the kernels are instantiated for */
#define DISPU_EWORKSPACE (-4)  /* scratch buffer missing or too small */

#define DISPU_ABI_VERSION 1
int dispu_abi_version(void);
/* Static string for a return code of any entry point (negative: ours, positive: CUDA). */
const char *dispu_error_string(int rc);
/* Number of kernel launches issued by this library since load (bench.py "gpu_launches"). */
uint64_t dispu_launch_count(void);

/* ---- tf_ops/sampling ------------------------------------------------------------------- */

/* FarthestPointSample.  Replaces farthestpointsamplingLauncher (tf_sampling.cpp:94 ->
 * tf_sampling_g.cu:203) and its kernel (tf_sampling_g.cu:105-170).  inp (b,n,3) -> out (b,m),
 * out[:,0]=0, bit-exact tie order of the reference.  `temp` is the
 * reference's (32,n) float scratch (tf_sampling.cpp:115); it is only touched when n exceeds
 * what fits on chip (n > dispu_fps_onchip_capacity()) and may be NULL otherwise.  The capacity is
 * 131072 points where 16-CTA (non-portable) clusters can be launched and 65536 where only the portable
 * 8-CTA clusters can (queried once per process with cudaOccupancyMaxPotentialClusterSize). */
int dispu_fps(int b, int n, int m, const float *inp, float *temp, int *out, dispu_stream_t stream);
int dispu_fps_onchip_capacity(void);

/* GatherPoint / GatherPointGrad.  gatherpointLauncher tf_sampling.cpp:125 (kernel
 * tf_sampling_g.cu:172) and scatteraddpointLauncher :150 (kernel :183; the zero fill the
 * wrapper does at :174 is done here). */
int dispu_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, dispu_stream_t stream);
int dispu_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g,
                            dispu_stream_t stream);

/* ProbSample.  probsampleLauncher tf_sampling.cpp:65 (kernels tf_sampling_g.cu:7,90): inclusive
 * prefix sum of inp_p (b,n) into temp (b,n), then for every inp_r (b,m) in [0,1) the first
 * bucket whose cumulative mass reaches r*total.  Registered for API completeness (its only
 * call site, tf_sampling.py:87, is commented out). */
int dispu_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out,
                      dispu_stream_t stream);

/* ---- tf_ops/grouping ------------------------------------------------------------------- */

/* QueryBallPoint.  queryBallPointLauncher tf_grouping.cpp:67 (kernel tf_grouping_g.cu:3-36).
 * xyz1 (b,n,3) dataset, xyz2 (b,m,3) queries, radius: DEVICE tensor (b,) of which only
 * element 0 is read (tf_grouping_g.cu:25) -> idx (b,m,nsample), pts_cnt (b,m).
 * One documented difference: a row with pts_cnt==0 is written as zeros (the reference leaves
 * it uninitialised). */
int dispu_query_ball_point(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                           const float *xyz2, int *idx, int *pts_cnt, dispu_stream_t stream);

/* Same op through a uniform grid (cell edge >= radius): identical idx / pts_cnt, O(points in the 27
 * neighbouring cells) work per query instead of O(n).  Needs scratch the reference interface has no
 * slot for: dispu_query_ball_point_workspace_bytes(b,n,m) bytes, any contents.  Falls back to the
 * brute-force kernel for n < 2048 or workspace == NULL. */
size_t dispu_query_ball_point_workspace_bytes(int b, int n, int m);
int dispu_query_ball_point_ws(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                              const float *xyz2, int *idx, int *pts_cnt, void *workspace, size_t workspace_bytes,
                              dispu_stream_t stream);

/* GroupPoint / GroupPointGrad.  groupPointLauncher tf_grouping.cpp:146 (kernel
 * tf_grouping_g.cu:40) and groupPointGradLauncher :177 (kernel :61; zero fill of
 * tf_grouping.cpp:208 done here). */
int dispu_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                      dispu_stream_t stream);
int dispu_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, dispu_stream_t stream);

/* GroupPoint fused with what the reference does to its output (no (b,m,nsample,c) intermediate):
 *   dispu_group_max:          out[b,i,:]   = max_s points[b,idx[b,i,s],:]        (group_point + reduce_max,
 *                             e.g. Common/ops.py:1049)
 *   dispu_group_affine_relu:  out[b,i,s,:] = relu(P[b,idx[b,i,s],:] + sign*Q[b,i,:])   — the first layer of a
 *                             grouped MLP hoisted to per-point GEMMs (Common/ops.py:1033-1058, :1897-1905);
 *                             Q (b,m,c) may be NULL. */
int dispu_group_max(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                    dispu_stream_t stream);
int dispu_group_affine_relu(int b, int n, int c, int m, int nsample, const float *P, const float *Q, float sign,
                            const int *idx, float *out, dispu_stream_t stream);
/* dispu_group_max with row strides (floats) on points and out: the result lands in its column slot of a wider buffer. */
int dispu_group_max_ld(int b, int n, int c, int m, int nsample, const float *points, int ldp, const int *idx, float *out,
                       int ldo, dispu_stream_t stream);

/* SelectionSort.  selectionSortLauncher tf_grouping.cpp:112 (kernel tf_grouping_g.cu:83). */
int dispu_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out,
                         dispu_stream_t stream);

/* kNN grouping in 3-D: k smallest squared L2 distances from every query xyz2[b,j] into
 * xyz1[b,:], ascending, ties to the lower index.  Replaces the three reference routes
 * (SURVEY.md §8 a5): nanoflann cpp_knn_batch_omp (libs/nearest_neighbors/knn_.cxx:104 via
 * Common/ops.py:110-118,165 — removing its device->host->device round trip), knn_point
 * (tf_grouping.py:116-141; it returns -dist) and the 3-D use of knn_point_2 (:95-114).
 * dist (b,m,k) may be NULL.  1 <= k <= 64, k <= n. */
int dispu_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *dist, int *idx,
              dispu_stream_t stream);

/* The same search for large k (64 < k <= 1024): the kNN patch extraction of the inference driver (DisPU/model.py:331 ->
 * Common/pc_util.py:83-91 extract_knn_patch, sklearn NearestNeighbors with k = patch_num_point = 256 on the host).
 * Radix select + ordered compaction + bitonic sort per query; ascending distance, ties to the lower index.
 * workspace: dispu_knn_large_workspace_bytes(b,n,m) bytes. */
size_t dispu_knn_large_workspace_bytes(int b, int n, int m);
int dispu_knn_large(int b, int n, int m, int k, const float *xyz1, const float *xyz2, float *dist, int *idx,
                    void *workspace, size_t workspace_bytes, dispu_stream_t stream);

/* ---- tf_ops/interpolation (the reference has CPU-only kernels here) --------------------- */

/* ThreeNN.  threenn_cpu tf_interpolate.cpp:60-103 (op :157-187): xyz1 (b,n,3) unknown,
 * xyz2 (b,m,3) known -> dist (b,n,3) SQUARED distances ascending, idx (b,n,3). */
int dispu_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                   dispu_stream_t stream);
/* ThreeInterpolate / Grad.  threeinterpolate_cpu :107-127, threeinterpolate_grad_cpu :131-153
 * (zero fill of :258 done here).  points (b,m,c), idx/weight (b,n,3) -> out (b,n,c). */
int dispu_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                            float *out, dispu_stream_t stream);
int dispu_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                 const float *weight, float *grad_points, dispu_stream_t stream);
/* Inverse-distance weights of pointnet_fp_module (Common/pointnet_util.py:205-208):
 * d=max(d,1e-10); w=(1/d)/sum_3(1/d).  dist (b,n,3) -> weight (b,n,3). */
int dispu_three_nn_weights(int b, int n, const float *dist, float *weight, dispu_stream_t stream);

/* ---- tf_ops/nn_distance + Common/loss_utils.py ------------------------------------------ */

/* NnDistance.  NmDistanceKernelLauncher tf_nndistance.cpp:168 (kernel tf_nndistance_g.cu:5-131).
 * Same argument order as the launcher. */
int dispu_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                      float *dist2, int *idx2, dispu_stream_t stream);
/* NnDistanceGrad.  NmDistanceGradKernelLauncher tf_nndistance.cpp:208 (tf_nndistance_g.cu:132-157). */
int dispu_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                           const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                           float *grad_xyz2, dispu_stream_t stream);

/* Fused NnDistance + chamfer() + hausdorff_loss() (Common/loss_utils.py:45-64, :67-84), called
 * like the trainer does (DisPU/model.py:75-79): chamfer(pred, gt) -> nn_distance(gt, pred).
 *   per_cloud (b,2): [cd_b, hd_b] = [(mean fwd + mean bwd)/radius_b, (max fwd + max bwd)/radius_b]
 *   out (4):  [mean_b cd_b, max_b hd_b, sum_b cd_b, (float)b]
 * out[2] is the partial a data-parallel rank contributes to the one all-reduce of the path.
 * radius (b,) may be NULL (= ones, DisPU/dataset.py:68).  per_cloud may be NULL.
 * workspace: dispu_chamfer_workspace_bytes(b, n_gt, n_pred) bytes, any contents.
 * Deterministic: fixed-order reductions, no float atomics. */
size_t dispu_chamfer_workspace_bytes(int b, int n_gt, int n_pred);
int dispu_chamfer(int b, int n_gt, const float *gt, int n_pred, const float *pred, const float *radius,
                  float *out, float *per_cloud, void *workspace, size_t workspace_bytes, dispu_stream_t stream);
/* The same, additionally storing this call's batch SUM of per-cloud cd (= out[2], the per-rank partial that the
 * multi-GPU all-reduce sums) to *sum_out and the batch max of hd (= out[1]) to *max_out — arbitrary
 * device addresses (either may be NULL), so several losses can deposit their partials side by side in ONE buffer
 * that a single collective reduces. */
int dispu_chamfer_ex(int b, int n_gt, const float *gt, int n_pred, const float *pred, const float *radius, float *out,
                     float *per_cloud, float *sum_out, float *max_out, void *workspace, size_t workspace_bytes,
                     dispu_stream_t stream);

/* Repulsion loss (Common/loss_utils.py:271-298, DisPU/model.py:83), fused after the ball query:
 * idx (b,n,nsample) = QueryBallPoint(radius, nsample, pred, pred); per point the five smallest squared
 * distances to its grouped points, drop the smallest, mean over (b,n,4) of max(0, h - d2).
 * out[0] = mean, out[1] = sum (multi-GPU partial).  workspace: dispu_repulsion_workspace_bytes(b,n). */
size_t dispu_repulsion_workspace_bytes(int b, int n);
int dispu_repulsion_loss(int b, int n, int nsample, float h, const float *pred, const int *idx, float *out,
                         void *workspace, size_t workspace_bytes, dispu_stream_t stream);

/* One percentage term of get_uniform_loss (Common/loss_utils.py:238-269), fused after FPS + ball query:
 * idx (b,npoint,nsample) = QueryBallPoint(sqrt(p), nsample, pcd, gather(pcd, FPS(npoint))); per grouped point the
 * distance u = sqrt(|d2 + 1e-8|) to its nearest other entry of the same group (knn_point(2, g, g)[:, :, 1]);
 * out[0] = mean over (b,npoint,nsample) of (u - expect_len)^2 / (expect_len + 1e-8), out[1] = the sum.
 * The caller scales by (100 p)^2 and averages the terms (loss_utils.py:264-268).
 * workspace: dispu_uniform_workspace_bytes(b,npoint,nsample). */
size_t dispu_uniform_workspace_bytes(int b, int npoint, int nsample);
int dispu_uniform_loss_term(int b, int n, int npoint, int nsample, float expect_len, const float *pcd, const int *idx,
                            float *out, void *workspace, size_t workspace_bytes, dispu_stream_t stream);

/* ---- Common/pointnet_util.py set abstraction --------------------------------------------- */

/* Fused group -> centre-subtract -> concat -> 1x1-conv MLP (bias, ReLU) -> max over nsample:
 * pointnet_sa_module / sample_and_group (pointnet_util.py:41-59, :117-133) and the _msg
 * channel order (:179), with tf_util.conv2d (Common/tf_util.py:120-185) as the layer.  The
 * (b,m,nsample,C) tensors of the reference are never materialised.
 *   xyz (b,n,3); points (b,n,c_pts) or NULL (c_pts=0); new_xyz (b,m,3); idx (b,m,nsample)
 *   layout 0: channels = [xyz-centre (3), points]   (sample_and_group, :53)
 *   layout 1: channels = [points, xyz-centre (3)]   (pointnet_sa_module_msg, :179)
 *   n_layers in 1..4; widths[0] = 3+c_pts, widths[l+1] = C_out of layer l;
 *   weights[l]: DEVICE (widths[l], widths[l+1]) row-major (= the (1,1,C_in,C_out) kernel);
 *   biases[l]: DEVICE (widths[l+1]) or NULL;  the arrays `widths/weights/biases` are HOST arrays.
 *   out (b,m,widths[n_layers]).
 *   precision: DISPU_MLP_FP32 = exact fp32 FMA chain; DISPU_MLP_TF32X3 = tcgen05 tensor cores
 *   with 3xTF32 error-compensated splits (fp32-level accuracy). */
#define DISPU_MLP_FP32 0
#define DISPU_MLP_TF32X3 1
#define DISPU_MLP_TF32 2 /* single-pass TF32 tensor cores: ~1e-3 relative, NOT fp32-level; opt-in only */
int dispu_sa_mlp(int b, int n, int m, int nsample, int c_pts, int layout, const float *xyz, const float *points,
                 const float *new_xyz, const int *idx, int n_layers, const int *widths_host,
                 const float *const *weights_host_array, const float *const *biases_host_array, float *out,
                 int precision, dispu_stream_t stream);

/* The same with the remaining options of pointnet_sa_module (Common/pointnet_util.py:91-149):
 *   use_xyz 0: channels = points only (:54-55; widths[0] = c_pts, c_pts > 0);
 *   pooling (:122-141): MAX; AVG = sum/nsample; WEIGHTED_AVG = sum_s x_s * w_s / sum_s w_s with
 *   w_s = exp(-5 |xyz_s - centre|); MIN = max(-x) (the reference pools -1*x and does not negate back);
 *   MAX_AND_AVG = [max | avg], out is (b,m,2*widths[n_layers]);
 *   bn_scale[l] / bn_shift[l]: DEVICE (widths[l+1]) or NULL — inference batch norm of tf_util.conv2d(bn=True)
 *   (Common/tf_util.py:178-181) folded to y = relu((x.W + b) * scale + shift), scale = gamma / sqrt(var + 1e-3),
 *   shift = beta - mean * scale; the arrays themselves are HOST arrays (may be NULL = no BN anywhere).
 * The tensor-core precisions support the hot-path form only (MAX, use_xyz, no BN): anything else with
 * precision != DISPU_MLP_FP32 returns DISPU_EUNSUPPORTED. */
#define DISPU_POOL_MAX 0
#define DISPU_POOL_AVG 1
#define DISPU_POOL_WEIGHTED_AVG 2
#define DISPU_POOL_MIN 3
#define DISPU_POOL_MAX_AND_AVG 4
int dispu_sa_mlp_ex(int b, int n, int m, int nsample, int c_pts, int layout, int use_xyz, int pooling, const float *xyz,
                    const float *points, const float *new_xyz, const int *idx, int n_layers, const int *widths_host,
                    const float *const *weights_host_array, const float *const *biases_host_array,
                    const float *const *bn_scale_host_array, const float *const *bn_shift_host_array, float *out,
                    int precision, dispu_stream_t stream);

/* Dense layer on tcgen05 with FP32-level accuracy (3xTF32, FP32 accumulate in TMEM), bias + ReLU fused:
 *     Y[M,N] = act(X[M,K] . W[K,N] + bias),  N <= 256, any K.
 * The unit the live generator's grouped MLPs reduce to after hoisting their first layer (Common/ops.py:1055-1072:
 * conv1 over (b,4N,16,128) rows, the [1,128]-conv = 2048->256 per point; tf_util.conv2d, Common/tf_util.py:120-185).
 * W is packed once (TF32 hi/lo planes in UMMA layout) into dispu_dense_packed_bytes(K,N) bytes by dispu_dense_pack. */
size_t dispu_dense_packed_bytes(int K, int N);
int dispu_dense_pack(int K, int N, const float *W, void *packed, dispu_stream_t stream);
int dispu_dense_tf32x3(long long M, int K, int N, const float *X, const void *packed, const float *bias, int relu,
                       float *Y, dispu_stream_t stream);

/* The same kernel with everything the generator's layers need to run without concat / slice / add passes
 * (HOST struct; zero-initialise, unset fields keep the plain behaviour):
 *   - ldx / ldy: row strides in floats (0 = dense), so a layer reads a column slice of a wider buffer and writes
 *     straight into its slot of the tensor the reference builds with tf.concat (Common/ops.py:1455-1486);
 *   - gather front-end (GroupPoint + centre term + ReLU fused into the operand load, Common/ops.py:1033-1058):
 *       A[r,:] = relu(X[cloud(r), gather_idx[r], :] + sign * Q[r / gather_ns, :]),  cloud(r) = r / gather_rows_per_cloud,
 *     X is then (clouds, gather_n, K) with row stride ldx and Q (M / gather_ns, K) with row stride ldq (NULL: no Q);
 *   - R1 / R2: residuals added AFTER the activation, Y = act(X.W + b) + R1 + R2 (tf.add at Common/ops.py:1074-1077);
 *   - batched weights: rows [i*rows_per_batch, (i+1)*rows_per_batch) multiply the packed weight at
 *     packed + i*packed_batch_stride bytes (per-cloud operands such as q.k^T and att.v of PointNonLocalCell,
 *     Common/ops.py:326-339); rows_per_batch must be a multiple of 128;
 *   - column blocks (n_blocks > 1): Y[:, j*N:(j+1)*N] = act(X . W_j + bias[j*N:(j+1)*N]) with W_j at
 *     packed + j*packed_block_stride (after the batch offset): one launch for an output wider than 256 columns
 *     (q.k^T: 1024 keys = 4 blocks) or for several layers that share their input (n_blocks * N <= 1024 with a bias);
 *     with Y2 != NULL the blocks j >= y2_block are written to a SECOND tensor Y2 (row stride ldy2, 0 = dense) at columns
 *     (j - y2_block)*N, so layers that only share their input keep separate compact outputs (N % 4 == 0, 16-byte aligned);
 *   - softmax without a softmax pass (PointNonLocalCell, Common/ops.py:326-339: softmax(q.k^T) . v): the launch that
 *     writes the scores also writes, per row and per (column block, warp half), the maximum of the values it stored:
 *     stat_out is (M, 2*n_blocks) float; the launch that consumes the scores as its X operand gets stat_in = that
 *     array and n_stat = 2*n_blocks: its producers turn every element into exp(x - max_row) while loading and sum
 *     them, its epilogue divides the row by the sum (K % 32 == 0).  The normalised attention matrix is never written.
 *   - weight-net epilogue (wn_xyz != NULL; needs the gather front-end with gather_ns == 16): the layer output G is
 *     never written; instead Y[r/16][ch*16+t] = sum_s G[16*(r/16)+s][ch] * wn[s][t] with
 *     wn[s][t] = relu(((xyz[cloud, gather_idx[r]] - xyz[r/16]) . wn_w[:,t] + wn_b[t]) * wn_scale): conv1 of PointShuffle2
 *     fused with weight_net_hidden and the transposed product (Common/ops.py:1055-1067); the 16-sample contraction runs on
 *     f16 (hi, scaled lo) pairs with FP32 accumulation — fp32-level accuracy for |activation| < 65504, inf/NaN beyond.
 *     Y is (M/16, 16*N), ldy its
 *     row stride (0 = dense), wn_w (3,16), wn_b (16) or NULL. */
typedef struct dispu_dense_args {
    long long M;
    int K, N;
    const float *X;
    long long ldx;
    const void *packed;
    const float *bias;
    int relu;
    float *Y;
    long long ldy;
    const int *gather_idx;
    int gather_n, gather_ns;
    long long gather_rows_per_cloud;
    const float *Q;
    long long ldq;
    float sign;
    const float *R1;
    long long ldr1;
    const float *R2;
    long long ldr2;
    long long rows_per_batch;
    long long packed_batch_stride;
    int n_blocks;                   /* > 1: column blocks, see below */
    long long packed_block_stride;
    const float *wn_xyz;            /* non-NULL: weight-net epilogue, see below */
    const float *wn_w, *wn_b;
    float wn_scale;
    void *stat_out;                 /* softmax statistics written by the epilogue, see below */
    const void *stat_in;            /* softmax applied to the operand rows by the producers, see below */
    int n_stat;
    float *Y2;                      /* split output: column blocks j >= y2_block go to Y2[:, (j-y2_block)*N ...], row stride ldy2 */
    long long ldy2;
    int y2_block;
} dispu_dense_args;
int dispu_dense_tf32x3_ex(const dispu_dense_args *args, dispu_stream_t stream);
/* sizeof(dispu_dense_args) as this library was compiled: lets a foreign-language binding verify its struct layout. */
size_t dispu_dense_args_size(void);
/* Pack W[k][n] = src[k*stride_k + n*stride_n] for `batch` operands (source advances by w_batch_stride floats,
 * destination by packed_batch_stride bytes, a multiple of 16 and >= dispu_dense_packed_bytes(K,N)). */
int dispu_dense_pack_strided(int K, int N, const float *W, long long stride_k, long long stride_n, int batch,
                             long long w_batch_stride, void *packed, long long packed_batch_stride,
                             dispu_stream_t stream);

/* A chain of two or three dense layers in ONE launch, intermediate activations kept in tensor memory (3xTF32):
 *     H1 = act1(X . W1 + b1),  H2 = act2(H1 . W2 + b2)  [,  H3 = act3(H2 . W3 + b3)]
 * — the coordinate regressor of the generator (Common/ops.py:1089-1110) behind its input layer.  HOST struct,
 * zero-initialise.  X (M,K0) row stride ldx (0 = dense); N[0] <= 256, N[1] <= 64, N[2] <= 256; W[l] packed by
 * dispu_dense_pack(K_l, N[l]) with K_0 = K0, K_l = N[l-1]; bias[l] (N[l]) or NULL; relu[l] 0/1.
 * Result of the last layer: with head == 0 all N_last columns go to Yt (row stride ldt, 0 = dense).  With head in 1..4
 * the last layer is laid out as [16-column head block | tail]: its first `head` columns are written to every non-NULL
 * Yh[d] (row stride ldh[d], 0 = head) — optionally as head_base[row*ldb + j] + sigmoid(x) - 0.5, the offset regressor
 * of Common/ops.py:1106-1108 + generator.py:78-81 — and columns 16.. (N_last - 16 of them) go to Yt. */
typedef struct dispu_mlp_chain_args {
    long long M;
    int K0;
    const float *X;
    long long ldx;
    int n_layers;
    int N[3];
    const void *W[3];
    const float *bias[3];
    int relu[3];
    int head;
    float *Yh[3];
    long long ldh[3];
    const float *head_base;
    long long ldb;
    float *Yt;
    long long ldt;
} dispu_mlp_chain_args;
int dispu_mlp_chain_tf32x3(const dispu_mlp_chain_args *args, dispu_stream_t stream);
size_t dispu_mlp_chain_args_size(void);

/* Non-local attention of PointNonLocalCell (Common/ops.py:326-339) as ONE tcgen05 kernel, 3xTF32 on both products:
 *     out[i,:] = sum_j softmax_j(q_i . k_j) v_j     per cloud, channels d = 64, m keys (m % 128 == 0)
 * so the (b,m,m) score tensor is never written.  q (b,m,64) float, row stride ldq, ALREADY scaled by 1/sqrt(d);
 * k_packed / v_packed: b*(m/128) packed operands of 64 KB each from dispu_dense_pack_strided, one per (cloud, block of
 * 128 keys): K as (K=64 channels, N=128 keys), V as (K=128 keys, N=64 channels); out (b,m,64), row stride ldo.
 * q, out and the packed operands must be 16-byte aligned, ldq and ldo multiples of 4. */
int dispu_attention_tf32x3(int b, int m, const float *q, long long ldq, const void *k_packed, const void *v_packed,
                           float *out, long long ldo, dispu_stream_t stream);

/* The K and V operands of dispu_attention_tf32x3 in ONE pass from the projected features: kv (b,m,>=128) row stride ld,
 * columns [0,64) = key, [64,128) = value; k_packed / v_packed receive b*(m/128) operands of 64 KB each, bit-identical
 * to dispu_dense_pack_strided's.  m % 128 == 0, ld % 4 == 0, 16-byte aligned pointers. */
int dispu_attention_pack(int b, int m, const float *kv, long long ld, void *k_packed, void *v_packed,
                         dispu_stream_t stream);

/* y = act(x . W + b) for K <= 8 input channels on CUDA cores (FP32 FMA chain in ascending k): the 3 -> 24 layer0 of
 * the feature extractor (Common/ops.py:1449-1452).  X (M,K) row stride ldx, W (K,N) row-major, Y (M,N) row stride ldy
 * (0 = dense). */
int dispu_linear_small(long long M, int K, int N, const float *X, long long ldx, const float *W, const float *bias,
                       int relu, float *Y, long long ldy, dispu_stream_t stream);

/* ---- the path's one collective over NVLink peer memory ------------------------------------------ */

/* Sum over the ranks of ONE node of a handful of floats (the per-rank Chamfer partials) as a single small kernel per rank:
 * every rank stores its values into its slot of every peer's "mailbox" (device memory mapped through CUDA IPC), waits
 * for the peers' slots of its own mailbox, and sums in rank order (bit-identical on every rank).  Setup, once, outside
 * the timed path: dispu_peer_alloc -> exchange the 64-byte handles through the host process group -> dispu_peer_open for
 * every other rank -> a DEVICE array `peers` of the `world` mailbox pointers (entry `rank` = the own mailbox) and one
 * zeroed device uint32 `epoch_ctr`.  dispu_peer_allreduce_sum is CUDA-graph capturable (the epoch advances on the device).
 * world <= 32, n <= 6.  A missing peer makes the call return NaNs after ~2 s instead of hanging. */
size_t dispu_peer_mailbox_bytes(void);
int dispu_peer_alloc(void **mailbox, unsigned char handle_out[64]);
int dispu_peer_open(const unsigned char handle[64], void **mailbox);
int dispu_peer_close(void *mailbox);
int dispu_peer_free(void *mailbox);
int dispu_peer_allreduce_sum(int world, int rank, void *const *peers, unsigned *epoch_ctr, const float *in, float *out, int n,
                             dispu_stream_t stream);

/* ---- feature-space kNN + EdgeConv dense block --------------------------------------------- */

/* knn_point_2 (tf_ops/grouping/tf_grouping.py:95-114) as get_edge_feature calls it (Common/ops.py:1866-1868):
 * k nearest neighbours in a c-dimensional feature space on D = (r_q - 2 q.p) + r_p (batch_distance_matrix_general,
 * tf_grouping.py:61-66), ascending, ties to the lower index (tf.nn.top_k); the first `skip` neighbours are dropped
 * (the reference drops the self match with idx[:, :, 1:]).  points (b,n,c) row stride ldp, queries (b,m,c) row
 * stride ldq -> dist / idx (b,m,k-skip); dist may be NULL.  Accumulation order is pinned (sequential fmaf over the
 * channels) and bit-identical to oracle/dispu_oracle.c:orc_feature_knn.  k <= 32, c <= 128. */
int dispu_feature_knn(int b, int n, int m, int c, int k, int skip, const float *points, int ldp, const float *queries,
                      int ldq, float *dist, int *idx, dispu_stream_t stream);

/* dense_conv (Common/ops.py:1897-1915) on the edge feature of get_edge_feature (:1856-1877), 3 layers of `growth`
 * channels, fused with the max over the k neighbours; no (b,n,k,*) tensor is built.
 *   feat (b,n,c) row stride ldf; idx (b,n,k) neighbour indices inside the cloud;
 *   w0 (2c,g) b0 (g) | w1 (g+c,g) b1 | w2 (2g+c,g) b2  row-major 1x1-conv kernels, biases may be NULL;
 *   out (b,n,3g+c) row stride ldo = [max y2 | max y1 | max y0 | feat]  (the concat order of :1901-1913).
 * workspace: dispu_dense_conv_workspace_bytes(b,n) bytes.  growth == 24, c % 4 == 0, c <= 128, ldf/ldo % 4 == 0. */
size_t dispu_dense_conv_workspace_bytes(int b, int n);
int dispu_dense_conv(int b, int n, int c, int k, int growth, const float *feat, int ldf, const int *idx, const float *w0,
                     const float *b0, const float *w1, const float *b1, const float *w2, const float *b2, float *out,
                     int ldo, void *workspace, size_t workspace_bytes, dispu_stream_t stream);
/* The same block INCLUDING get_edge_feature's neighbour search (k+1 nearest in feature space, self dropped): the kNN
 * kernel also evaluates the centre-point projections of its own queries, so the block is two launches in total.
 * idx_out (b,n,k) receives the neighbour indices the block used. */
int dispu_dense_conv_knn(int b, int n, int c, int k, int growth, const float *feat, int ldf, int *idx_out, const float *w0,
                         const float *b0, const float *w1, const float *b1, const float *w2, const float *b2, float *out,
                         int ldo, void *workspace, size_t workspace_bytes, dispu_stream_t stream);

/* Weight net + transposed product of PointShuffle2 (Common/ops.py:1062-1067, weight_net_hidden :181-191):
 *   wn[s][t]   = relu(((xyz[idx[i][s]] - xyz[i]) . Ww[:,t] + bw[t]) * bn_scale)
 *   out[i][ch*T + t] = sum_s G[i][s][ch] * wn[s][t]
 * xyz (b,m,3), idx (b,m,nsample), G (b,m,nsample,C), Ww (3,T), bw (T) or NULL -> out (b,m,C*T), the operand of the
 * [1,C] after_conv (:1068).  Instantiated for nsample = 16, C = 128, T = 16 (the live generator). */
int dispu_weightnet_aggregate(int b, int m, int nsample, int C, int T, const float *xyz, const int *idx, const float *G,
                              const float *Ww, const float *bw, float bn_scale, float *out, dispu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DISPU_B200_H_ */

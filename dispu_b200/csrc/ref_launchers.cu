// Link seam: the reference's own launcher prototypes (C++-mangled, exactly as its unmodified
// TF wrappers declare them) forwarding to the C ABI on the legacy default stream the reference
// launches on.  With these symbols exported, tf_ops/*/tf_*.cpp links against libdispu_b200.so in
// place of tf_*_g.cu.o (tf_ops/compile_ops.sh:11,14,20) and re-registers the same op names —
// see INTEGRATION.md.
//
//   tf_ops/sampling/tf_sampling.cpp:65,94,125,150
//   tf_ops/grouping/tf_grouping.cpp:67,112,146,177
//   tf_ops/nn_distance/tf_nndistance.cpp:168,208
//
// The reference launchers return void and never check errors; failures here are reported on
// stderr once per call site instead of being swallowed.
#include <cstdio>

#include "common.cuh"

static void report(const char *what, int rc) {
    if (rc != DISPU_OK) std::fprintf(stderr, "dispu_b200: %s failed: %s (%d)\n", what, dispu_error_string(rc), rc);
}

void farthestpointsamplingLauncher(int b, int n, int m, const float *inp, float *temp, int *out) {
    report("FarthestPointSample", dispu_fps(b, n, m, inp, temp, out, nullptr));
}
void gatherpointLauncher(int b, int n, int m, const float *inp, const int *idx, float *out) {
    report("GatherPoint", dispu_gather_point(b, n, m, inp, idx, out, nullptr));
}
// The reference wrapper zero-fills inp_g itself (tf_sampling.cpp:174) and this launcher only
// accumulates; dispu_gather_point_grad zero-fills again, which is idempotent.
void scatteraddpointLauncher(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    report("GatherPointGrad", dispu_gather_point_grad(b, n, m, out_g, idx, inp_g, nullptr));
}
void probsampleLauncher(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
    report("ProbSample", dispu_prob_sample(b, n, m, inp_p, inp_r, temp, out, nullptr));
}
void queryBallPointLauncher(int b, int n, int m, const float *radius, int nsample, const float *xyz1,
                            const float *xyz2, int *idx, int *pts_cnt) {
    report("QueryBallPoint", dispu_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, nullptr));
}
void selectionSortLauncher(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
    report("SelectionSort", dispu_selection_sort(b, n, m, k, dist, outi, out, nullptr));
}
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
    report("GroupPoint", dispu_group_point(b, n, c, m, nsample, points, idx, out, nullptr));
}
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                            float *grad_points) {
    report("GroupPointGrad", dispu_group_point_grad(b, n, c, m, nsample, grad_out, idx, grad_points, nullptr));
}
void NmDistanceKernelLauncher(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                              float *result2, int *result2_i) {
    report("NnDistance", dispu_nn_distance(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, nullptr));
}
void NmDistanceGradKernelLauncher(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                                  const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                                  float *grad_xyz2) {
    report("NnDistanceGrad", dispu_nn_distance_grad(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1,
                                                    grad_xyz2, nullptr));
}

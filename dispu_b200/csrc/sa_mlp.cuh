// Shared declarations of the fused set-abstraction MLP kernels (sa_mlp.cu: exact FP32 path,
// sa_mlp_tc.cu: tcgen05 tensor-core path).
#pragma once
#include "common.cuh"

namespace dispu {

constexpr int kMlpMaxLayers = 4;

struct MlpParams {
    int n_layers;
    int widths[kMlpMaxLayers + 1];
    const float *w[kMlpMaxLayers];
    const float *bias[kMlpMaxLayers];
    // inference batch norm folded to a per-channel affine applied after the bias, before the ReLU
    // (tf_util.conv2d bn=True, Common/tf_util.py:178-181): y = relu((x.W + b) * bn_scale + bn_shift); NULL = no BN
    const float *bn_scale[kMlpMaxLayers];
    const float *bn_shift[kMlpMaxLayers];
};

// Launches the tensor-core kernel when the shape is one it is built for; *handled tells.
// precision: DISPU_MLP_TF32X3 (fp32-level) or DISPU_MLP_TF32 (single pass).
int sa_mlp_tc_launch(int b, int n, int m, int ns, int c_pts, int layout, const float *xyz, const float *points,
                     const float *new_xyz, const int *idx, const MlpParams &prm, float *out, int precision,
                     cudaStream_t stream, bool *handled);

}  // namespace dispu

// Tensor-core (tcgen05) path of the fused set-abstraction MLP — placeholder until the kernel lands.
#include "common.cuh"

namespace dispu {
struct MlpParams;
int sa_mlp_tc_launch(int, int, int, int, int, int, const float *, const float *, const float *, const int *,
                     const MlpParams &, float *, cudaStream_t, bool *handled) {
    *handled = false;
    return DISPU_EUNSUPPORTED;
}
}  // namespace dispu

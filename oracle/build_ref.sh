#!/usr/bin/env bash
# Build the REAL reference code for this path into oracle/_ref/ (git-ignored, travels to the
# GPU box as prebuilt .so files).  Sources are compiled where they lie under $REF (read-only);
# nothing is copied into the repository — only shared objects are written.
#
#   libref_gpu.so  tf_sampling_g.cu + tf_grouping_g.cu + tf_nndistance_g.cu, unmodified, with the
#                  reference's own flags (tf_ops/compile_ops.sh:10,13,19: -O2 -DGOOGLE_CUDA=1,
#                  default -fmad) plus the sm_100a target.  Exports the reference's C++-mangled
#                  launchers (tf_sampling.cpp:65,94,125,150; tf_grouping.cpp:67,112,146,177;
#                  tf_nndistance.cpp:168,208).  Needs a GPU to run -> used by the -m gpu tests.
#   libref_cpu.so  the host loops that live inside TF-dependent / non-compiling files, taken by
#                  line range straight into the compiler's stdin (compile_ops.sh:8 flags:
#                  g++ -std=c++11 -O2): nnsearch tf_nndistance.cpp:21-43 (static -> wrapped),
#                  threenn_cpu/threeinterpolate_cpu/_grad_cpu tf_interpolate.cpp:60-153,
#                  query_ball_point_cpu/group_point_cpu/_grad_cpu query_ball_point.cpp:19-84,
#                  selection_sort_cpu selection_sort.cpp:20-63.
#   libref_knn.so  nanoflann 1.2.3 kNN: libs/nearest_neighbors/knn_.cxx compiled whole.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/tf_ops" ]; then
  echo "build_ref: $REF not present (GPU box?) — keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"

"$NVCC" -O2 -DGOOGLE_CUDA=1 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
  -o "$OUT/libref_gpu.so" \
  "$REF/tf_ops/sampling/tf_sampling_g.cu" \
  "$REF/tf_ops/grouping/tf_grouping_g.cu" \
  "$REF/tf_ops/nn_distance/tf_nndistance_g.cu"

{
  cat <<'EOF'
#include <cstdio>
#include <ctime>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
using namespace std;
EOF
  sed -n '21,43p' "$REF/tf_ops/nn_distance/tf_nndistance.cpp"
  cat <<'EOF'
extern "C" void ref_nnsearch(int b,int n,int m,const float*xyz1,const float*xyz2,float*dist,int*idx){
  nnsearch(b,n,m,xyz1,xyz2,dist,idx);
}
EOF
  sed -n '60,153p' "$REF/tf_ops/interpolation/tf_interpolate.cpp"
  sed -n '19,84p' "$REF/tf_ops/grouping/query_ball_point.cpp"
  sed -n '20,63p' "$REF/tf_ops/grouping/selection_sort.cpp"
} | g++ -std=c++11 -O2 -shared -fPIC -x c++ - -o "$OUT/libref_cpu.so"

g++ -std=c++11 -O2 -fopenmp -shared -fPIC -I "$REF/libs/nearest_neighbors" \
  "$REF/libs/nearest_neighbors/knn_.cxx" -o "$OUT/libref_knn.so"

echo "build_ref: wrote $(ls "$OUT" | tr '\n' ' ')"

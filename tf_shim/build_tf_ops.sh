#!/usr/bin/env bash
# Builds the four op libraries the reference's Python wrappers load (file names and directories are
# part of the contract: tf_sampling.py:12, tf_grouping.py:8, tf_interpolate.py:7, tf_nndistance.py:12)
# from tf_shim/dispu_tf_ops.cc + libdispu_b200.so.  Needs the TensorFlow of the reference install;
# NOT runnable in the build container (no TensorFlow) — see INTEGRATION.md.
#   usage: tf_shim/build_tf_ops.sh /path/to/Dis-PU/tf_ops
set -euo pipefail
DST="${1:?usage: build_tf_ops.sh <reference tf_ops dir>}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
LIBDIR="$HERE/../dispu_b200"
TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
for pair in sampling:DISPU_TF_SAMPLING grouping:DISPU_TF_GROUPING interpolation:DISPU_TF_INTERPOLATE nn_distance:DISPU_TF_NNDISTANCE; do
  dir="${pair%%:*}"; def="${pair##*:}"
  name="tf_${dir}_so.so"; [ "$dir" = interpolation ] && name="tf_interpolate_so.so"; [ "$dir" = nn_distance ] && name="tf_nndistance_so.so"
  g++ -std=c++14 -shared -fPIC -O2 -D"$def" "$HERE/dispu_tf_ops.cc" -o "$DST/$dir/$name" \
      ${TF_CFLAGS} ${TF_LFLAGS} -L"$LIBDIR" -ldispu_b200 -Wl,-rpath,"$LIBDIR"
  echo "built $DST/$dir/$name"
done

"""dispu_b200 — B200-native (sm_100a) implementation of the Dis-PU point-sampling / grouping /
interpolation / Chamfer hot path, behind the reference's own op interface.

Modules mirror the reference files they stand in for:
    tf_sampling, tf_grouping, tf_interpolate, tf_nndistance   <- tf_ops/*/tf_*.py
    pointnet_util                                             <- Common/pointnet_util.py
    loss_utils                                                <- Common/loss_utils.py (chamfer, hausdorff)
    ops                                                       <- Common/ops.py:110-179 (knn_query, sampling, grouping)
Everything computes through the C ABI of libdispu_b200.so (include/dispu_b200.h); there is no
CPU or PyTorch fallback.
"""
__all__ = ["tf_sampling", "tf_grouping", "tf_interpolate", "tf_nndistance", "pointnet_util", "loss_utils", "ops"]

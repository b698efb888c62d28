"""ctypes binding of the C ABI declared in include/dispu_b200.h.

There is no fallback: if libdispu_b200.so is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdispu_b200.so")

EINVAL, ENULL, EUNSUPPORTED, EWORKSPACE = -1, -2, -3, -4

_i, _p, _sz, _ll = C.c_int, C.c_void_p, C.c_size_t, C.c_longlong


class DenseArgs(C.Structure):
    """dispu_dense_args of include/dispu_b200.h (host struct)."""
    _fields_ = [
        ("M", _ll), ("K", _i), ("N", _i), ("X", _p), ("ldx", _ll), ("packed", _p), ("bias", _p), ("relu", _i),
        ("Y", _p), ("ldy", _ll), ("gather_idx", _p), ("gather_n", _i), ("gather_ns", _i), ("gather_rows_per_cloud", _ll),
        ("Q", _p), ("ldq", _ll), ("sign", C.c_float), ("R1", _p), ("ldr1", _ll), ("R2", _p), ("ldr2", _ll),
        ("rows_per_batch", _ll), ("packed_batch_stride", _ll), ("n_blocks", _i), ("packed_block_stride", _ll),
        ("wn_xyz", _p), ("wn_w", _p), ("wn_b", _p), ("wn_scale", C.c_float),
        ("stat_out", _p), ("stat_in", _p), ("n_stat", _i),
        ("Y2", _p), ("ldy2", _ll), ("y2_block", _i),
    ]


class MlpChainArgs(C.Structure):
    """dispu_mlp_chain_args of include/dispu_b200.h (host struct)."""
    _fields_ = [
        ("M", _ll), ("K0", _i), ("X", _p), ("ldx", _ll), ("n_layers", _i), ("N", _i * 3), ("W", _p * 3), ("bias", _p * 3),
        ("relu", _i * 3), ("head", _i), ("Yh", _p * 3), ("ldh", _ll * 3), ("head_base", _p), ("ldb", _ll), ("Yt", _p), ("ldt", _ll),
    ]


# name -> argtypes, exactly the prototypes of include/dispu_b200.h
SIGNATURES = {
    "dispu_abi_version": ([], C.c_int),
    "dispu_error_string": ([_i], C.c_char_p),
    "dispu_launch_count": ([], C.c_uint64),
    "dispu_fps": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_fps_onchip_capacity": ([], _i),
    "dispu_gather_point": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_gather_point_grad": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_prob_sample": ([_i, _i, _i, _p, _p, _p, _p, _p], _i),
    "dispu_query_ball_point": ([_i, _i, _i, _p, _i, _p, _p, _p, _p, _p], _i),
    "dispu_query_ball_point_workspace_bytes": ([_i, _i, _i], _sz),
    "dispu_query_ball_point_ws": ([_i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_group_point": ([_i, _i, _i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_group_point_grad": ([_i, _i, _i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_group_max": ([_i, _i, _i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_group_max_ld": ([_i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p], _i),
    "dispu_group_affine_relu": ([_i, _i, _i, _i, _i, _p, _p, C.c_float, _p, _p, _p], _i),
    "dispu_selection_sort": ([_i, _i, _i, _i, _p, _p, _p, _p], _i),
    "dispu_knn": ([_i, _i, _i, _i, _p, _p, _p, _p, _p], _i),
    "dispu_knn_large_workspace_bytes": ([_i, _i, _i], _sz),
    "dispu_knn_large": ([_i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_three_nn": ([_i, _i, _i, _p, _p, _p, _p, _p], _i),
    "dispu_three_interpolate": ([_i, _i, _i, _i, _p, _p, _p, _p, _p], _i),
    "dispu_three_interpolate_grad": ([_i, _i, _i, _i, _p, _p, _p, _p, _p], _i),
    "dispu_three_nn_weights": ([_i, _i, _p, _p, _p], _i),
    "dispu_nn_distance": ([_i, _i, _p, _i, _p, _p, _p, _p, _p, _p], _i),
    "dispu_nn_distance_grad": ([_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p], _i),
    "dispu_chamfer_workspace_bytes": ([_i, _i, _i], _sz),
    "dispu_chamfer": ([_i, _i, _p, _i, _p, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_chamfer_ex": ([_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_repulsion_workspace_bytes": ([_i, _i], _sz),
    "dispu_repulsion_loss": ([_i, _i, _i, C.c_float, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_uniform_workspace_bytes": ([_i, _i, _i], _sz),
    "dispu_uniform_loss_term": ([_i, _i, _i, _i, C.c_float, _p, _p, _p, _p, _sz, _p], _i),
    "dispu_dense_packed_bytes": ([_i, _i], _sz),
    "dispu_dense_pack": ([_i, _i, _p, _p, _p], _i),
    "dispu_dense_tf32x3": ([C.c_longlong, _i, _i, _p, _p, _p, _i, _p, _p], _i),
    "dispu_dense_tf32x3_ex": ([C.POINTER(DenseArgs), _p], _i),
    "dispu_dense_args_size": ([], _sz),
    "dispu_dense_pack_strided": ([_i, _i, _p, _ll, _ll, _i, _ll, _p, _ll, _p], _i),
    "dispu_mlp_chain_tf32x3": ([C.POINTER(MlpChainArgs), _p], _i),
    "dispu_mlp_chain_args_size": ([], _sz),
    "dispu_peer_mailbox_bytes": ([], _sz),
    "dispu_peer_alloc": ([C.POINTER(_p), C.c_char_p], _i),
    "dispu_peer_open": ([C.c_char_p, C.POINTER(_p)], _i),
    "dispu_peer_close": ([_p], _i),
    "dispu_peer_free": ([_p], _i),
    "dispu_peer_allreduce_sum": ([_i, _i, _p, _p, _p, _p, _i, _p], _i),
    "dispu_attention_pack": ([_i, _i, _p, _ll, _p, _p, _p], _i),
    "dispu_linear_small": ([_ll, _i, _i, _p, _ll, _p, _p, _i, _p, _ll, _p], _i),
    "dispu_attention_tf32x3": ([_i, _i, _p, _ll, _p, _p, _p, _ll, _p], _i),
    "dispu_feature_knn": ([_i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p], _i),
    "dispu_dense_conv_workspace_bytes": ([_i, _i], _sz),
    "dispu_dense_conv": ([_i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p], _i),
    "dispu_dense_conv_knn": ([_i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p], _i),
    "dispu_weightnet_aggregate": ([_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, C.c_float, _p, _p], _i),
    "dispu_sa_mlp": ([_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p], _i),
    "dispu_sa_mlp_ex": ([_i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p], _i),
}

_lib = None


class DispuError(RuntimeError):
    def __init__(self, fn, rc, msg):
        super().__init__(f"{fn} failed: {msg} (rc={rc})")
        self.rc = rc


def lib():
    """Load libdispu_b200.so.  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RuntimeError(
                f"{SO} is missing: build it with `python -m dispu_b200.build` (nvcc, sm_100a). "
                "dispu_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(SO)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = res
        if l.dispu_abi_version() != 1:
            raise RuntimeError("libdispu_b200.so ABI version mismatch")
        if l.dispu_dense_args_size() != C.sizeof(DenseArgs):
            raise RuntimeError("dispu_dense_args layout mismatch between include/dispu_b200.h and dispu_b200/_lib.py")
        if l.dispu_mlp_chain_args_size() != C.sizeof(MlpChainArgs):
            raise RuntimeError("dispu_mlp_chain_args layout mismatch between include/dispu_b200.h and dispu_b200/_lib.py")
        _lib = l
    return _lib


def check(fn: str, rc: int):
    if rc != 0:
        msg = lib().dispu_error_string(rc).decode()
        if rc in (EINVAL, ENULL):
            raise ValueError(f"{fn}: {msg} (rc={rc})")
        raise DispuError(fn, rc, msg)


def launch_count() -> int:
    return int(lib().dispu_launch_count())

"""ctypes bindings to the REAL reference code built by oracle/build_ref.sh into oracle/_ref/.
TEST INFRASTRUCTURE ONLY (same rule as oracle.py).

cpu_*  : the reference host loops (libref_cpu.so, libref_knn.so) — run anywhere.
gpu_*  : the reference CUDA kernels recompiled for sm_100a (libref_gpu.so) — need a GPU; they
         take raw device pointers (ints) and launch on the legacy default stream exactly as
         the reference launchers do, so callers synchronise around them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def available(which: str) -> bool:
    return os.path.exists(os.path.join(_REF, f"libref_{which}.so"))


_libs = {}


def _lib(which):
    if which not in _libs:
        _libs[which] = C.CDLL(os.path.join(_REF, f"libref_{which}.so"))
    return _libs[which]


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- host loops ---------------------------------------------------------------------------

def cpu_nnsearch(xyz1, xyz2):
    """tf_nndistance.cpp:21-43."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d = np.empty((b, n), np.float32)
    i = np.empty((b, n), np.int32)
    _lib("cpu").ref_nnsearch(b, n, m, _p(xyz1), _p(xyz2), _p(d), _p(i))
    return d, i


def cpu_nn_distance(xyz1, xyz2):
    """NnDistanceOp::Compute tf_nndistance.cpp:79-80: nnsearch both ways."""
    d1, i1 = cpu_nnsearch(xyz1, xyz2)
    d2, i2 = cpu_nnsearch(xyz2, xyz1)
    return d1, i1, d2, i2


def cpu_three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d = np.empty((b, n, 3), np.float32)
    i = np.empty((b, n, 3), np.int32)
    _lib("cpu")._Z11threenn_cpuiiiPKfS0_PfPi(b, n, m, _p(xyz1), _p(xyz2), _p(d), _p(i))
    return d, i


def cpu_three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _i(idx), _f(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    _lib("cpu")._Z20threeinterpolate_cpuiiiiPKfPKiS0_Pf(b, m, c, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def cpu_three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f(points), _i(idx), _f(weight), _f(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.zeros((b, m, c), np.float32)
    _lib("cpu")._Z25threeinterpolate_grad_cpuiiiiPKfPKiS0_Pf(b, n, c, m, _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


def cpu_query_ball_point(radius, nsample, xyz, new_xyz):
    """query_ball_point.cpp:19-47 (no pts_cnt; rows without a hit stay at the -1 prefill)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    r = np.array([radius], np.float32)
    idx = np.full((b, m, nsample), -1, np.int32)
    _lib("cpu")._Z20query_ball_point_cpuiiiPKfiS0_S0_Pi(b, n, m, _p(r), nsample, _p(xyz), _p(new_xyz), _p(idx))
    return idx


def cpu_group_point(points, idx):
    points, idx = _f(points), _i(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    _lib("cpu")._Z15group_point_cpuiiiiiPKfPKiPf(b, n, c, m, ns, _p(points), _p(idx), _p(out))
    return out


def cpu_group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f(points), _i(idx), _f(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.zeros((b, n, c), np.float32)
    _lib("cpu")._Z20group_point_grad_cpuiiiiiPKfPKiPf(b, n, c, m, ns, _p(grad_out), _p(idx), _p(g))
    return g


def cpu_selection_sort(dist, k):
    dist = _f(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    _lib("cpu")._Z18selection_sort_cpuiiiiPKfPiPf(b, n, m, k, _p(dist), _p(outi), _p(out))
    return outi, out


def cpu_knn_batch(support, query, k, omp=True):
    """libs/nearest_neighbors/knn_.cxx:104-135 (omp) / :75-102 — the call made by
    Common/ops.py:117 through knn.pyx:71; returns int64 like the Cython wrapper."""
    support, query = _f(support), _f(query)
    b, n, dim = support.shape
    m = query.shape[1]
    out = np.zeros((b, m, k), np.int64)
    fn = _lib("knn")._Z17cpp_knn_batch_ompPKfmmmS0_mmPl if omp else _lib("knn")._Z13cpp_knn_batchPKfmmmS0_mmPl
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    fn(_p(support), b, n, dim, _p(query), m, k, _p(out))
    return out


# ---- CUDA kernels (device pointers as ints) -------------------------------------------------

def _vp(x):
    return C.c_void_p(int(x))


def gpu_fps(b, n, m, inp, temp, out):
    """tf_sampling.cpp:94 farthestpointsamplingLauncher(b,n,m,inp,temp,out); temp is (32,n) f32."""
    _lib("gpu")._Z29farthestpointsamplingLauncheriiiPKfPfPi(b, n, m, _vp(inp), _vp(temp), _vp(out))


def gpu_gather_point(b, n, m, inp, idx, out):
    _lib("gpu")._Z19gatherpointLauncheriiiPKfPKiPf(b, n, m, _vp(inp), _vp(idx), _vp(out))


def gpu_scatter_add_point(b, n, m, out_g, idx, inp_g):
    _lib("gpu")._Z23scatteraddpointLauncheriiiPKfPKiPf(b, n, m, _vp(out_g), _vp(idx), _vp(inp_g))


def gpu_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt):
    _lib("gpu")._Z22queryBallPointLauncheriiiPKfiS0_S0_PiS1_(b, n, m, _vp(radius), nsample, _vp(xyz1), _vp(xyz2),
                                                             _vp(idx), _vp(pts_cnt))


def gpu_group_point(b, n, c, m, nsample, points, idx, out):
    _lib("gpu")._Z18groupPointLauncheriiiiiPKfPKiPf(b, n, c, m, nsample, _vp(points), _vp(idx), _vp(out))


def gpu_group_point_grad(b, n, c, m, nsample, grad_out, idx, grad_points):
    _lib("gpu")._Z22groupPointGradLauncheriiiiiPKfPKiPf(b, n, c, m, nsample, _vp(grad_out), _vp(idx), _vp(grad_points))


def gpu_selection_sort(b, n, m, k, dist, outi, out):
    _lib("gpu")._Z21selectionSortLauncheriiiiPKfPiPf(b, n, m, k, _vp(dist), _vp(outi), _vp(out))


def gpu_nn_distance(b, n, xyz, m, xyz2, result, result_i, result2, result2_i):
    _lib("gpu")._Z24NmDistanceKernelLauncheriiPKfiS0_PfPiS1_S2_(b, n, _vp(xyz), m, _vp(xyz2), _vp(result), _vp(result_i),
                                                                _vp(result2), _vp(result2_i))


def gpu_nn_distance_grad(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2):
    _lib("gpu")._Z28NmDistanceGradKernelLauncheriiPKfiS0_S0_PKiS0_S2_PfS3_(
        b, n, _vp(xyz1), m, _vp(xyz2), _vp(grad_dist1), _vp(idx1), _vp(grad_dist2), _vp(idx2), _vp(grad_xyz1),
        _vp(grad_xyz2))

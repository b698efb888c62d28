"""The C-ABI library builds, loads and exports every symbol include/dispu_b200.h declares, plus the
reference's C++-mangled launcher symbols (the link seam).  No compute calls (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dispu_b200.h")


@pytest.fixture(scope="module")
def so():
    from dispu_b200 import build
    return build.build()


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dispu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(so):
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(so)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dispu_b200.h but not exported"


def test_binding_covers_header(so):
    from dispu_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    l = _lib.lib()
    assert l.dispu_abi_version() == 1
    assert l.dispu_fps_onchip_capacity() in (65536, 131072)  # 131072 where 16-CTA clusters launch (B200); 65536 otherwise / without a device
    assert b"invalid" in l.dispu_error_string(-1)
    assert l.dispu_chamfer_workspace_bytes(32, 1024, 1024) == 8 * 32 * 2 * 8 + 8 * 32  # float2 partial per (cloud, direction, 128-query chunk)


def test_dense_args_layout_and_validation(so):
    """The ctypes mirror of dispu_dense_args has the compiled struct's size; bad arguments are rejected before any CUDA call."""
    import ctypes as C
    from dispu_b200 import _lib
    l = _lib.lib()
    assert l.dispu_dense_args_size() == C.sizeof(_lib.DenseArgs)
    a = _lib.DenseArgs()
    assert l.dispu_dense_tf32x3_ex(None, None) == _lib.ENULL
    a.M, a.K, a.N = 128, 32, 300
    assert l.dispu_dense_tf32x3_ex(C.byref(a), None) == _lib.EUNSUPPORTED  # N > 256 per column block
    a.N = 64
    assert l.dispu_dense_tf32x3_ex(C.byref(a), None) == _lib.ENULL         # X / packed / Y missing
    a.M = 0
    assert l.dispu_dense_tf32x3_ex(C.byref(a), None) == 0                  # empty problem: nothing to do
    assert l.dispu_dense_packed_bytes(100, 64) == 4 * 2 * 64 * 32 * 4      # 4 K-chunks x [hi|lo] x 64 rows x 32 k
    assert l.dispu_feature_knn(1, 8, 8, 4, 9, 0, None, 4, None, 4, None, None, None) == _lib.EINVAL   # k > n
    assert l.dispu_feature_knn(1, 64, 8, 4, 33, 0, None, 4, None, 4, None, None, None) == _lib.EUNSUPPORTED
    assert l.dispu_dense_conv(1, 8, 6, 4, 24, None, 6, None, None, None, None, None, None, None, None, 80, None, 0, None) == _lib.ENULL
    assert l.dispu_knn_large(1, 100, 4, 2000, None, None, None, None, None, 0, None) == _lib.EINVAL
    assert l.dispu_knn_large_workspace_bytes(2, 100, 4) == 2 * 4 * 100 * 4


def test_reference_launcher_symbols_exported(so):
    """Mangled names of the launchers the reference wrappers call (tf_sampling.cpp:65,94,125,150;
    tf_grouping.cpp:67,112,146,177; tf_nndistance.cpp:168,208)."""
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    for sym in ["_Z29farthestpointsamplingLauncheriiiPKfPfPi", "_Z19gatherpointLauncheriiiPKfPKiPf",
                "_Z23scatteraddpointLauncheriiiPKfPKiPf", "_Z18probsampleLauncheriiiPKfS0_PfPi",
                "_Z22queryBallPointLauncheriiiPKfiS0_S0_PiS1_", "_Z21selectionSortLauncheriiiiPKfPiPf",
                "_Z18groupPointLauncheriiiiiPKfPKiPf", "_Z22groupPointGradLauncheriiiiiPKfPKiPf",
                "_Z24NmDistanceKernelLauncheriiPKfiS0_PfPiS1_S2_",
                "_Z28NmDistanceGradKernelLauncheriiPKfiS0_S0_PKiS0_S2_PfS3_"]:
        assert sym in out, sym


def test_argument_validation_needs_no_gpu(so):
    """Rejected arguments return the negative ABI codes before any CUDA call."""
    from dispu_b200 import _lib
    l = _lib.lib()
    assert l.dispu_fps(1, 0, 4, None, None, None, None) == _lib.EINVAL
    assert l.dispu_fps(1, 8, 0, None, None, None, None) == _lib.EINVAL
    assert l.dispu_fps(1, 8, 4, None, None, None, None) == _lib.ENULL
    assert l.dispu_query_ball_point(1, 8, 4, None, 0, None, None, None, None, None) == _lib.EINVAL
    assert l.dispu_knn(1, 8, 4, 9, None, None, None, None, None) == _lib.EINVAL
    assert l.dispu_knn(1, 100, 4, 65, None, None, None, None, None) == _lib.EUNSUPPORTED
    assert l.dispu_chamfer(2, 8, None, 8, None, None, None, None, None, 0, None) == _lib.ENULL
    assert l.dispu_nn_distance(0, 8, None, 8, None, None, None, None, None, None) == 0


def test_product_has_no_oracle_dependency():
    """The shipped package must not import or reference oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "dispu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("SURVEY", ""), f"{f} mentions oracle"


def test_ops_refuse_cpu_tensors():
    import torch
    from dispu_b200 import tf_sampling
    with pytest.raises(RuntimeError, match="GPU-only"):
        tf_sampling.farthest_point_sample(4, torch.zeros((1, 8, 3)))

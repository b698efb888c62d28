"""CPU tests of the oracle (oracle/dispu_oracle.c): pinned against the committed golden vectors
that the REAL reference code produced (tests/golden/ref_cpu.npz from the reference host loops,
tests/golden/ref_gpu.npz from the reference CUDA kernels run on a B200), against the live
oracle/_ref libraries when they are present, and against independent numpy restatements."""
import os

import numpy as np
import pytest

from conftest import lattice, unit_ball
from oracle import oracle as O
from oracle import ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gc():
    return np.load(os.path.join(GOLD, "ref_cpu.npz"))


@pytest.fixture(scope="module")
def gg():
    p = os.path.join(GOLD, "ref_gpu.npz")
    if not os.path.exists(p):
        pytest.skip("ref_gpu.npz not generated yet")
    return np.load(p)


# ---- golden vectors from the reference host loops -------------------------------------------
@pytest.mark.parametrize("tag", ["ball", "lat"])
def test_golden_cpu_nn_distance(gc, tag):
    a, b = gc[f"{tag}_a"], gc[f"{tag}_b"]
    d1, i1, d2, i2 = O.nn_distance(a, b, mode=O.CPU)
    assert np.array_equal(d1, gc[f"{tag}_nn_d1"]) and np.array_equal(i1, gc[f"{tag}_nn_i1"])
    assert np.array_equal(d2, gc[f"{tag}_nn_d2"]) and np.array_equal(i2, gc[f"{tag}_nn_i2"])


@pytest.mark.parametrize("tag", ["ball", "lat"])
def test_golden_cpu_three_nn(gc, tag):
    d, i = O.three_nn(gc[f"{tag}_a"], gc[f"{tag}_b"])
    assert np.array_equal(d, gc[f"{tag}_3nn_d"]) and np.array_equal(i, gc[f"{tag}_3nn_i"])


def test_golden_cpu_interpolate_and_weights(gc):
    w = O.three_nn_weights(gc["ball_3nn_d"])
    np.testing.assert_allclose(w, gc["ball_3nn_w"], rtol=1e-6)
    out = O.three_interpolate(gc["feat_b"], gc["ball_3nn_i"], gc["ball_3nn_w"])
    assert np.array_equal(out, gc["ball_interp"])


@pytest.mark.parametrize("tag", ["ball", "lat"])
def test_golden_cpu_query_ball_point(gc, tag):
    idx, cnt = O.query_ball_point(0.3, 16, gc[f"{tag}_a"], gc[f"{tag}_b"], mode=O.CPU, fill_empty=False)
    assert np.array_equal(idx, gc[f"{tag}_qbp_r0.3_ns16"])  # rows without a hit keep the -1 prefill in both


def test_golden_cpu_group_point(gc):
    assert np.array_equal(O.group_point(gc["feat_a"], gc["ball_group_idx"]), gc["ball_group"])


def test_golden_cpu_knn_nanoflann(gc):
    d, i = O.knn(16, gc["ball_a"], gc["ball_a"])
    assert np.array_equal(i, gc["ball_knn16"])  # distinct distances: identical order, self at slot 0
    assert (i[:, :, 0] == np.arange(600)[None]).all()
    # lattice: equal distances are ordered by KD-tree traversal in nanoflann -> compare distances
    a = gc["lat_a"]
    d, i = O.knn(16, a, a)
    gi = gc["lat_knn16"]
    gd = ((a[:, :, None, :] - np.take_along_axis(a[:, None], gi[..., None].astype(np.int64), 2)) ** 2)
    gd = (gd[..., 0] + gd[..., 1]) + gd[..., 2]
    assert np.array_equal(d, gd.astype(np.float32))


# ---- golden vectors from the reference CUDA kernels (generated on a B200) ---------------------
def test_golden_gpu_fps(gc, gg):
    assert np.array_equal(O.farthest_point_sample(150, gc["ball_a"]), gg["ball_fps150"])
    assert np.array_equal(O.farthest_point_sample(200, gc["lat_a"]), gg["lat_fps200"])
    assert np.array_equal(O.farthest_point_sample(64, gc["big"]), gg["big_fps64"])
    assert np.array_equal(O.farthest_point_sample(200, gc["ball_b"]), gg["ballb_fps200"])


@pytest.mark.parametrize("tag", ["ball", "lat"])
def test_golden_gpu_query_ball_and_nn(gc, gg, tag):
    a, q = gc[f"{tag}_a"], gc[f"{tag}_b"]
    for r, ns in ((0.3, 16), (0.12, 8)):
        idx, cnt = O.query_ball_point(r, ns, a, q, mode=O.GPU, fill_empty=True)
        assert np.array_equal(cnt, gg[f"{tag}_qbp_r{r}_ns{ns}_cnt"])
        assert np.array_equal(idx, gg[f"{tag}_qbp_r{r}_ns{ns}_idx"])  # generator pre-zeroed empty rows
    d1, i1, d2, i2 = O.nn_distance(a, q, mode=O.GPU)
    assert np.array_equal(d1, gg[f"{tag}_nn_d1"]) and np.array_equal(i1, gg[f"{tag}_nn_i1"])
    assert np.array_equal(d2, gg[f"{tag}_nn_d2"]) and np.array_equal(i2, gg[f"{tag}_nn_i2"])


# ---- real clouds shipped with the reference (data/test/*.xyz, output/*_X4.xyz): goldens from the reference host code ----
@pytest.fixture(scope="module")
def real():
    return np.load(os.path.join(GOLD, "real_clouds.npz"))


def knn_dist(a, idx):
    g = np.take_along_axis(a[:, None], idx[..., None].astype(np.int64), 2)
    d = a[:, :, None, :] - g
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).astype(np.float32)


@pytest.mark.parametrize("name", ["Icosahedron", "fandisk"])
def test_real_clouds_oracle_vs_reference_host_code(real, name):
    """Icosahedron.xyz has exact distance ties (symmetry) — the tie-break stressor SURVEY.md §8c names."""
    x, y = real[name], real[name + "_X4"]
    d1, i1, d2, i2 = O.nn_distance(x, y, mode=O.CPU)
    assert np.array_equal(d1, real[name + "_nn_d1"]) and np.array_equal(i1, real[name + "_nn_i1"])
    assert np.array_equal(d2, real[name + "_nn_d2"]) and np.array_equal(i2, real[name + "_nn_i2"])
    d, i = O.three_nn(y, x)
    assert np.array_equal(d, real[name + "_3nn_d"]) and np.array_equal(i, real[name + "_3nn_i"])
    idx, _ = O.query_ball_point(0.2, 32, x, x[:, ::8].copy(), mode=O.CPU, fill_empty=False)
    assert np.array_equal(idx, real[name + "_qbp_r0.2_ns32"])
    # nanoflann orders exactly equal distances by tree traversal: distances must agree everywhere, indices wherever
    # a neighbour's distance is distinct from its rank neighbours'
    d, i = O.knn(16, x, x)
    gi = real[name + "_knn16"]
    assert np.array_equal(d, knn_dist(x, gi))
    uniq = np.ones_like(d, bool)
    uniq[..., 1:] &= d[..., 1:] != d[..., :-1]
    uniq[..., :-1] &= d[..., :-1] != d[..., 1:]
    uniq[..., -1] = False  # the last slot may tie with the first neighbour left out
    assert np.array_equal(i[uniq], gi[uniq]) and uniq.mean() > 0.5


@pytest.mark.parametrize("name", ["Icosahedron", "fandisk"])
def test_real_clouds_evaluate_metrics_golden(real, name):
    """evaluate.py:34-39,153-165 (CD / HD of the reference's own 4x output against its input cloud) through the oracle."""
    def norm(pc):
        c = pc.mean(axis=1, keepdims=True, dtype=np.float32)
        pc = (pc - c).astype(np.float32)
        return (pc / np.sqrt((pc * pc).sum(-1, keepdims=True, dtype=np.float32)).max(axis=1, keepdims=True)).astype(np.float32)
    p, g = norm(real[name + "_X4"]), norm(real[name])
    d_f, _, d_b, _ = O.nn_distance(p, g, mode=O.CPU)
    assert float(d_f.astype(np.float64).mean() + d_b.astype(np.float64).mean()) == pytest.approx(float(real[name + "_eval_cd"]), rel=1e-12)
    assert float(d_f.max() + d_b.max()) == pytest.approx(float(real[name + "_eval_hd"]), rel=1e-7)
    assert 0 < float(real[name + "_eval_cd"]) < 1e-2  # a 4x upsampling of the same surface sits close to it


# ---- live reference host loops (present where oracle/build_ref.sh ran) -----------------------
needs_ref = pytest.mark.skipif(not (R.available("cpu") and R.available("knn")), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("n,m", [(1, 1), (5, 6), (257, 33), (64, 700)])
def test_live_ref_host_loops(rng, n, m):
    a, b = unit_ball(rng, 2, n), unit_ball(rng, 2, m)
    for got, want in zip(O.nn_distance(a, b, mode=O.CPU), R.cpu_nn_distance(a, b)):
        assert np.array_equal(got, want)
    for got, want in zip(O.three_nn(a, b), R.cpu_three_nn(a, b)):
        assert np.array_equal(got, want)
    idx, _ = O.query_ball_point(0.4, 8, a, b, mode=O.CPU, fill_empty=False)
    assert np.array_equal(idx, R.cpu_query_ball_point(0.4, 8, a, b))
    feat = rng.standard_normal((2, m, 7)).astype(np.float32)
    d, i = O.three_nn(a, b)
    w = O.three_nn_weights(d) if m >= 3 else np.full((2, n, 3), 1 / 3, np.float32)
    assert np.array_equal(O.three_interpolate(feat, i, w), R.cpu_three_interpolate(feat, i, w))
    go = rng.standard_normal((2, n, 7)).astype(np.float32)
    assert np.array_equal(O.three_interpolate_grad(feat, i, w, go), R.cpu_three_interpolate_grad(feat, i, w, go))
    if n >= 16:
        assert np.array_equal(O.knn(16, a, a)[1], R.cpu_knn_batch(a, a, 16).astype(np.int32))


# ---- independent numpy restatements -----------------------------------------------------------
def fps_numpy(xyz, m):
    """A.1 of SURVEY.md evaluated with numpy float32 ops (FMA emulated in float64 then rounded once)."""
    b, n, _ = xyz.shape
    out = np.zeros((b, m), np.int32)
    for i in range(b):
        p = xyz[i].astype(np.float32)
        mind = np.full(n, np.float32(1e38))
        last = 0
        lane = np.arange(n) & 511
        for j in range(1, m):
            dx, dy, dz = (p[:, c] - p[last, c] for c in range(3))
            t = (dy * dy).astype(np.float32)
            t = (dx.astype(np.float64) * dx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            d = (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            mind = np.minimum(d, mind)
            cand = np.flatnonzero(mind == mind.max())
            last = int(cand[np.lexsort((cand, lane[cand]))[0]])
            out[i, j] = last
    return out


@pytest.mark.parametrize("n,m,kind", [(300, 40, "ball"), (1300, 50, "lat"), (700, 30, "lat"), (20, 30, "ball")])
def test_fps_against_numpy_restatement(rng, n, m, kind):
    xyz = unit_ball(rng, 2, n) if kind == "ball" else lattice(rng, 2, n, q=5)
    assert np.array_equal(O.farthest_point_sample(m, xyz), fps_numpy(xyz, m))


def test_nn_distance_matches_tf_spec(rng):
    """tf_ops/nn_distance/tf_nndistance_cpu.py:4-25: tile/sub/square/reduce_min/argmin."""
    a, b = unit_ball(rng, 2, 40), unit_ball(rng, 2, 55)
    diff = a[:, :, None, :] - b[:, None, :, :]
    d = (diff ** 2).sum(-1)
    d1, i1, d2, i2 = O.nn_distance(a, b, mode=O.CPU)
    np.testing.assert_allclose(d1, d.min(2), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(d2, d.min(1), rtol=1e-5, atol=1e-7)
    assert np.array_equal(i1, d.argmin(2)) and np.array_equal(i2, d.argmin(1))


def test_chamfer_definition(rng):
    pred, gt = unit_ball(rng, 3, 64), unit_ball(rng, 3, 128)
    radius = np.array([1.0, 2.0, 0.5], np.float32)
    o = O.chamfer(pred, gt, radius)
    df, _, db, _ = O.nn_distance(gt, pred)
    cd = ((df.mean(1) + db.mean(1)) / radius).mean()
    hd = ((df.max(1) + db.max(1)) / radius).max()
    np.testing.assert_allclose(o["cd"], cd, rtol=1e-6)
    np.testing.assert_allclose(o["hd"], hd, rtol=1e-6)


def test_sa_mlp_against_numpy(rng):
    b, n, m, ns, c = 2, 200, 20, 16, 6
    xyz = unit_ball(rng, b, n)
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    new_xyz, new_points, idx, _ = O.sample_and_group(m, 0.5, ns, xyz, pts)
    ws = [rng.standard_normal((3 + c, 12)).astype(np.float32), rng.standard_normal((12, 7)).astype(np.float32)]
    bs = [rng.standard_normal(12).astype(np.float32), None]
    x = new_points.astype(np.float64)
    for w, bb in zip(ws, bs):
        x = np.maximum(x @ w.astype(np.float64) + (0 if bb is None else bb), 0)
    np.testing.assert_allclose(O.sa_mlp(xyz, pts, new_xyz, idx, ws, bs, 0), x.max(2), rtol=1e-4, atol=1e-5)


def test_sa_module_numpy_restatement_matches_c_oracle(rng):
    """The float64 numpy restatement of pointnet_sa_module's body (pooling variants, mlp2, BN: oracle.sa_module) agrees with
    the C oracle's orc_sa_mlp on the form both cover (max pooling, [xyz | points] and the _msg channel order), and its
    variants satisfy the identities the reference's definitions imply (pointnet_util.py:122-141)."""
    b, n, m, ns, c = 2, 200, 16, 8, 5
    xyz = unit_ball(rng, b, n)
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    new_xyz = O.gather_point(xyz, O.farthest_point_sample(m, xyz))
    idx, _ = O.query_ball_point(0.5, ns, xyz, new_xyz)
    ws = [(rng.standard_normal((3 + c, 16)) / 3).astype(np.float32), (rng.standard_normal((16, 12)) / 4).astype(np.float32)]
    bs = [rng.standard_normal(16).astype(np.float32) * 0.1, None]
    layers = list(zip(ws, bs))
    for layout in (0, 1):
        np.testing.assert_allclose(O.sa_module(xyz, pts, new_xyz, idx, layers, layout=layout),
                                   O.sa_mlp(xyz, pts, new_xyz, idx, ws, bs, layout), rtol=1e-5, atol=1e-5)
    mx = O.sa_module(xyz, pts, new_xyz, idx, layers, pooling="max")
    av = O.sa_module(xyz, pts, new_xyz, idx, layers, pooling="avg")
    both = O.sa_module(xyz, pts, new_xyz, idx, layers, pooling="max_and_avg")
    np.testing.assert_array_equal(both, np.concatenate([mx, av], -1))
    mn = O.sa_module(xyz, pts, new_xyz, idx, layers, pooling="min")
    assert (mn <= 0).all() and (-mn <= av + 1e-12).all() and (av <= mx + 1e-12).all()  # post-ReLU: 0 <= min <= avg <= max
    wa = O.sa_module(xyz, pts, new_xyz, idx, layers, pooling="weighted_avg")
    assert (wa <= mx + 1e-12).all() and (wa >= -mn - 1e-12).all()
    # identity BN statistics leave the layer unchanged up to the epsilon
    ident = (np.ones(16, np.float32), np.zeros(16, np.float32), np.zeros(16, np.float32), np.ones(16, np.float32) - 1e-3)
    np.testing.assert_allclose(O.sa_module(xyz, pts, new_xyz, idx, [(ws[0], bs[0], ident), (ws[1], bs[1])]), mx, rtol=1e-6, atol=1e-7)  # float32(1 - 1e-3) + 1e-3 is 1 to ~2e-8


def test_selection_sort_known_answer():
    """tf_ops/grouping/selection_sort.cpp:65-94: k=3 of a 2x2x4 descending ramp -> idx 3 2 1 ..."""
    dist = np.arange(16, 0, -1, dtype=np.float32).reshape(2, 2, 4)
    outi, out = O.selection_sort(dist, 3)
    assert (outi[..., :3] == np.array([3, 2, 1])).all()
    assert np.array_equal(out[..., :3], np.sort(dist, -1)[..., :3])


def test_repulsion_loss_against_numpy(rng):
    pred = (unit_ball(rng, 2, 300) * np.float32(0.4)).astype(np.float32)
    idx, _ = O.query_ball_point(0.07, 20, pred, pred)
    g = O.group_point(pred, idx) - pred[:, :, None, :]
    d = np.sort((g.astype(np.float64) ** 2).sum(-1), axis=-1)[:, :, 1:5]
    want = np.maximum(0.0, 0.001 - d).mean()
    np.testing.assert_allclose(O.repulsion_loss(pred, idx, 0.001), want, rtol=1e-5)


# ---- feature-space kNN / EdgeConv dense block / weight net: oracle vs independent numpy restatements
def test_feature_knn_oracle_vs_numpy():
    """orc_feature_knn vs tf_grouping.py:61-66 + top_k restated in float64 (indices where the gaps are clear)."""
    rng = np.random.default_rng(5)
    f = rng.standard_normal((2, 96, 24)).astype(np.float32)
    q = rng.standard_normal((2, 40, 24)).astype(np.float32)
    d, i = O.feature_knn(9, f, q)
    r_q = (q.astype(np.float64) ** 2).sum(-1)
    r_p = (f.astype(np.float64) ** 2).sum(-1)
    D = r_q[:, :, None] - 2 * np.einsum("bic,bjc->bij", q.astype(np.float64), f.astype(np.float64)) + r_p[:, None, :]
    assert np.array_equal(i, np.argsort(D, axis=-1, kind="stable")[:, :, :9])
    np.testing.assert_allclose(d, np.sort(D, axis=-1)[:, :, :9], atol=2e-5)
    # exact ties go to the lower index; skip drops the leading neighbours (get_edge_feature, ops.py:1868)
    t = np.zeros((1, 8, 4), np.float32)
    assert np.array_equal(O.feature_knn(4, t, t)[1][0, 3], [0, 1, 2, 3])
    assert np.array_equal(O.feature_knn(4, t, t, skip=1)[1][0, 3], [1, 2, 3])


def test_dense_conv_oracle_vs_numpy():
    """orc_dense_conv vs the literal concat/conv2d sequence of Common/ops.py:1897-1915 in numpy float64."""
    rng = np.random.default_rng(6)
    b, n, c, k, g = 2, 50, 12, 7, 24
    f = rng.standard_normal((b, n, c)).astype(np.float32)
    _, idx = O.feature_knn(k + 1, f, f, skip=1)
    ws = [(rng.standard_normal((ci, g)) / 4).astype(np.float32) for ci in (2 * c, g + c, 2 * g + c)]
    bs = [(rng.standard_normal(g) / 4).astype(np.float32) for _ in range(3)]
    got = O.dense_conv(f, idx, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])
    fi = np.broadcast_to(f[:, :, None, :], (b, n, k, c)).astype(np.float64)
    fj = f[np.arange(b)[:, None, None], idx].astype(np.float64)
    y = np.concatenate([fi, fj - fi], -1)
    y = np.concatenate([np.maximum(y @ ws[0] + bs[0], 0), fi], -1)
    y = np.concatenate([np.maximum(y @ ws[1] + bs[1], 0), y], -1)
    y = np.concatenate([y @ ws[2] + bs[2], y], -1)
    np.testing.assert_allclose(got, y.max(2), rtol=1e-5, atol=1e-5)


def test_weightnet_aggregate_oracle_vs_numpy():
    rng = np.random.default_rng(7)
    b, m, ns, C, T = 2, 30, 16, 128, 16
    xyz = rng.uniform(-1, 1, (b, m, 3)).astype(np.float32)
    idx = rng.integers(0, m, (b, m, ns)).astype(np.int32)
    G = rng.standard_normal((b, m, ns, C)).astype(np.float32)
    Ww = rng.standard_normal((3, T)).astype(np.float32)
    bw = rng.standard_normal(T).astype(np.float32)
    H = O.weightnet_aggregate(xyz, idx, G, Ww, bw, 0.9995)
    gx = xyz[np.arange(b)[:, None, None], idx] - xyz[:, :, None, :]
    wn = np.maximum((gx.astype(np.float64) @ Ww + bw) * 0.9995, 0)
    want = np.einsum("bmsc,bmst->bmct", G.astype(np.float64), wn).reshape(b, m, -1)
    np.testing.assert_allclose(H, want, rtol=1e-5, atol=1e-5)


def test_f_rows_golden_fixture():
    """The oracle's §8(f) restatements against the committed golden vectors (tests/golden/f_rows.npz, generated by
    make_golden_f.py from independent numpy float64 restatements of the cited TensorFlow expressions)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f_rows.npz"))
    d, i = O.feature_knn(17, g["knn_feat"], g["knn_feat"])
    assert np.array_equal(i, g["knn_i"])
    np.testing.assert_allclose(d, g["knn_d"], atol=3e-5)
    out = O.dense_conv(g["knn_feat"], g["dc_idx"], g["dc_w0"], g["dc_b0"], g["dc_w1"], g["dc_b1"], g["dc_w2"], g["dc_b2"])
    np.testing.assert_allclose(out, g["dc_out"], rtol=1e-5, atol=1e-5)
    H = O.weightnet_aggregate(g["wn_xyz"], g["wn_idx"], g["wn_G"], g["wn_Ww"], g["wn_bw"], 1.0 / np.sqrt(1.0 + 1e-3))
    np.testing.assert_allclose(H, g["wn_out"], rtol=1e-5, atol=2e-5)
    t = O.uniform_loss_term(g["un_pcd"], g["un_idx"], float(g["un_expect"]))
    assert abs(float(t) - float(g["un_term"])) <= 1e-5 * max(1.0, abs(float(g["un_term"])))

"""Torch restatement of the LIVE Dis-PU generator forward (DisPU/generator.py:31-88) used only to
drive the hot path at the shapes of BASELINE.json configs 2 and 4 (the TF graph itself cannot
run here).  It is a caller of the path, not the product: the sampling / grouping / Chamfer ops
it calls are the sm_100a kernels of this package; the stock dense layers (1x1 convs, the
non-local attention) are plain torch matmuls exactly as the reference leaves them to stock TF.

Shape trace and literals follow Common/ops.py:
  feature_extraction_GCN ops.py:1437-1486, dense_conv :1897-1915, get_edge_feature :1856-1877,
  duplicate_up :1152-1199, gen_grid :60-76, coordinate_regressor :1089-1110,
  PointShuffle2 :1012-1087, grouping :154-179, weight_net_hidden :181-191,
  PointNonLocalCell :302-346; conv = tf_util.conv2d/conv1d (Common/tf_util.py:52-185): Xavier
  weights, zero bias, ReLU unless stated, no BN (use_bn=False, generator.py:33) except the
  weight net (bn=True hard-coded, ops.py:187) which runs with its initial moving statistics.

`ops` is the backend for the path's own operators:
    ops.knn3d(k, support (b,n,3), query (b,m,3)) -> idx (b,m,k)       Common/ops.py:110-118
    ops.group(feature (b,n,c), idx (b,m,k))      -> (b,m,k,c)         tf.gather_nd, ops.py:169,174
    ops.feature_knn(k, feat (b,n,c))             -> idx (b,n,k)       tf_grouping.py:95-114
The default backend is the CUDA one below; bench.py supplies a CPU backend for the reference arm.
"""
from __future__ import annotations

import math
import os

import torch

from . import tf_grouping

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (Common/tf_util.py:527)


class DispuOps:
    """The path's operators on the B200 kernels."""

    def knn3d(self, k, support, query):
        return tf_grouping.knn(k, support, query)[1]

    def group(self, feature, idx):
        return tf_grouping.group_point(feature, idx)

    def feature_knn(self, k, feat):
        # knn_point_2 (tf_grouping.py:95-114): D = |q|^2 - 2 q.p + |p|^2 ; top_k(-D, k)
        r = (feat * feat).sum(dim=2, keepdim=True)
        d = r - 2.0 * torch.bmm(feat, feat.transpose(1, 2)) + r.transpose(1, 2)
        return torch.topk(-d, k, dim=2, sorted=True)[1].to(torch.int32)

    # fused forms of GroupPoint (no (b,m,k,c) intermediates); their presence switches the generator to the
    # hoisted formulation of the first grouped layer (see Generator._dense_conv / point_shuffle)
    def group_max(self, feature, idx, out=None):
        return tf_grouping.group_max(feature, idx, out=out)

    def group_affine_relu(self, P, Q, idx, sign):
        return tf_grouping.group_affine_relu(P, Q, idx, sign)


def _xavier(gen, cin, cout, device, fan_in=None, fan_out=None):
    fi = cin if fan_in is None else fan_in
    fo = cout if fan_out is None else fan_out
    lim = math.sqrt(6.0 / (fi + fo))
    w = (torch.rand((cin, cout), generator=gen, dtype=torch.float32) * 2 - 1) * lim
    return w.to(device), torch.zeros((cout,), dtype=torch.float32, device=device)


class Generator:
    """coarse, fine = G(inputs (b,n,3)); up_ratio 4: (b,4n,3) each."""

    def __init__(self, up_ratio=4, device="cuda", seed=7, ops=None, hoist=True, tc=True, fused=None):
        self.up_ratio = up_ratio
        self.tc = tc  # dense layers on the tcgen05 3xTF32 kernel (GPU only)
        # fused: the whole forward on this library's kernels with no concat / slice / gather intermediates
        # (forward_fused); default = whenever the CUDA backend, hoisting and the tensor-core layers are all on
        self.fused = fused
        self.fuse_weightnet = os.environ.get("DISPU_FUSE_WEIGHTNET", "1") == "1"
        self.attention = os.environ.get("DISPU_ATTENTION", "fused")  # "fused" (one kernel) | "two_products"
        # coordinate regressors (fc0 -> fc1 -> fc2) as ONE launch each, activations in tensor memory (csrc/mlp_chain_tc.cu)
        self.chain = os.environ.get("DISPU_CHAIN", "1") == "1"
        self._bufs = {}
        self._streams = {}
        # independent branches (the 3-D kNN of the coarse cloud, the caller's coarse-loss hook) run on side streams so
        # they can overlap the non-local branch; inside a CUDA graph this becomes a fork/join.  Opt-in (DISPU_STREAMS=1): measured
        # neutral on B200 — the persistent tensor-core kernels fill every SM (registers, shared and tensor memory), so the side
        # stream's CTAs only run in the gaps they would have had anyway.
        self.use_streams = os.environ.get("DISPU_STREAMS", "0") == "1"
        self._dup_idx = {}
        self._packed = {}
        self.hoist = hoist  # use the fused / hoisted grouped-layer formulation when the backend offers it
        self.device = torch.device(device)
        self.ops = ops if ops is not None else DispuOps()
        g = torch.Generator().manual_seed(seed)
        d = self.device
        P = {}
        P["layer0"] = _xavier(g, 3, 24, d)
        c_in = 24
        for blk in range(1, 5):
            if blk > 1:
                P[f"layer{blk}_prep"] = _xavier(g, 120 * (blk - 1), 48, d)
                c_in = 48
            P[f"layer{blk}_l0"] = _xavier(g, 2 * c_in, 24, d)
            P[f"layer{blk}_l1"] = _xavier(g, 24 + c_in, 24, d)
            P[f"layer{blk}_l2"] = _xavier(g, 48 + c_in, 24, d)
        P["up_conv1"] = _xavier(g, 482, 256, d)
        P["up_conv2"] = _xavier(g, 256, 128, d)
        for name, cin in (("coarse", 128), ("fine", 256)):
            P[f"{name}_fc0"] = _xavier(g, cin, 256, d)
            P[f"{name}_fc1"] = _xavier(g, 256, 64, d)
            P[f"{name}_fc2"] = _xavier(g, 64, 3, d)
        P["nl_kv"] = _xavier(g, 128, 128, d)
        P["nl_query"] = _xavier(g, 128, 64, d)
        P["nl_back"] = _xavier(g, 64, 256, d)
        P["skip"] = _xavier(g, 134, 256, d)
        P["conv0"] = _xavier(g, 134, 128, d)
        P["conv1"] = _xavier(g, 128, 128, d)
        P["wconv0"] = _xavier(g, 3, 16, d)
        # after_conv kernel (1,128,16,256): rows ordered (c, s) to match the [B,N,C,S] operand
        P["after_conv"] = _xavier(g, 128 * 16, 256, d, fan_in=128 * 16, fan_out=128 * 256)
        P["aggregation"] = _xavier(g, 256, 256, d)
        self.P = P
        # gen_grid(4): 2x2 mesh of linspace(-0.2,0.2,2), rows = (x,y) pairs in meshgrid order
        sq = int(math.sqrt(up_ratio)) + 1
        for i in range(sq, 0, -1):
            if up_ratio % i == 0:
                num_x, num_y = i, up_ratio // i
                break
        gx = torch.linspace(-0.2, 0.2, num_x)
        gy = torch.linspace(-0.2, 0.2, num_y)
        yy, xx = torch.meshgrid(gy, gx, indexing="ij")
        self.grid = torch.stack([xx, yy], dim=-1).reshape(-1, 2).to(d)

    # -- layers ---------------------------------------------------------------------------------
    def _fc(self, x, name, act=True):
        """1x1 conv = dense layer over the last axis with bias (+ ReLU) fused.  On the GPU every such layer runs on
        the tcgen05 3xTF32 kernel (dense.PackedLinear, fp32-level accuracy); the CPU arm uses torch addmm."""
        w, b = self.P[name]
        if self.tc and x.is_cuda and w.shape[1] <= 256:
            lin = self._packed.get(name)
            if lin is None:
                from .dense import PackedLinear
                lin = self._packed[name] = PackedLinear(w, b)
            return lin(x, relu=act)
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.addmm(b, x2, w)
        if act:
            y = torch.relu_(y)
        return y.view(*x.shape[:-1], w.shape[1])

    def _lin(self, key, x, w, b=None, act=False):
        """Same as _fc for derived (hoisted) weights, cached under `key`."""
        if self.tc and x.is_cuda and w.shape[1] <= 256:
            lin = self._packed.get(key)
            if lin is None:
                from .dense import PackedLinear
                lin = self._packed[key] = PackedLinear(w, b)
            return lin(x, relu=act)
        y = torch.matmul(x, w)
        if b is not None:
            y = y + b
        return torch.relu_(y) if act else y

    def _dense_conv_hoisted(self, feature, blk, k=16):
        """Same function as _dense_conv with every term that does not depend on the neighbour hoisted to ONE
        per-point GEMM (exact reassociation of the 1x1 convs over concat[central, neigh-central | y | central]):
            y0 = relu(neigh.W0e + central.(W0c-W0e) + b0)         -> fused gather + centre term + ReLU kernel
            y1 = relu(y0.W1a + (central.W1c + b1))
            y2 =      y1.W2a + y0.W2b + (central.W2c + b2)
        so neither the tiled `central` nor any concatenated (b,n,k,*) tensor is materialised."""
        c = feature.shape[-1]
        (w0, b0), (w1, b1), (w2, b2) = (self.P[f"layer{blk}_l{i}"] for i in range(3))
        key = f"_hoist{blk}"
        if key not in self.P:
            wcat = torch.cat([w0[c:], w0[:c] - w0[c:], w1[24:], w2[48:]], dim=1).contiguous()  # (c, 96)
            bcat = torch.cat([torch.zeros_like(b0), b0, b1, b2]).contiguous()
            self.P[key] = (wcat, bcat, w1[:24].contiguous(), w2[:24].contiguous(), w2[24:48].contiguous())
        wcat, bcat, w1a, w2a, w2b = self.P[key]
        idx = self.ops.feature_knn(k + 1, feature)[:, :, 1:].contiguous()
        proj = self._lin(key, feature, wcat, bcat)
        pn, pc, t1, t2 = (proj[..., i * 24:(i + 1) * 24].contiguous() for i in range(4))
        y0 = self.ops.group_affine_relu(pn, pc, idx, 1.0)  # (b,n,k,24)
        y1 = torch.relu_(torch.matmul(y0, w1a) + t1.unsqueeze(2))
        y2 = torch.matmul(y1, w2a) + torch.matmul(y0, w2b) + t2.unsqueeze(2)
        return torch.cat([y2.max(dim=2)[0], y1.max(dim=2)[0], y0.max(dim=2)[0], feature], dim=-1)

    def _dense_conv(self, feature, blk, k=16):
        """ops.py:1897-1915 on (b,n,c) -> (b,n,72+c)."""
        if hasattr(self.ops, "group_affine_relu") and self.hoist:
            return self._dense_conv_hoisted(feature, blk, k)
        idx = self.ops.feature_knn(k + 1, feature)[:, :, 1:].contiguous()  # drop self, ops.py:1867-1868
        neigh = self.ops.group(feature, idx)  # (b,n,k,c)
        central = feature.unsqueeze(2).expand(-1, -1, k, -1)
        y = torch.cat([central, neigh - central], dim=-1)
        y = torch.cat([self._fc(y, f"layer{blk}_l0"), central], dim=-1)
        y = torch.cat([self._fc(y, f"layer{blk}_l1"), y], dim=-1)
        y = torch.cat([self._fc(y, f"layer{blk}_l2", act=False), y], dim=-1)
        return y.max(dim=2)[0]

    def feature_extraction(self, inputs):
        l0 = self._fc(inputs, "layer0", act=False)
        out = torch.cat([self._dense_conv(l0, 1), l0], dim=-1)  # 120
        for blk in (2, 3, 4):
            prep = self._fc(out, f"layer{blk}_prep")
            out = torch.cat([self._dense_conv(prep, blk), out], dim=-1)
        return out  # (b,n,480)

    def duplicate_up(self, feat):
        b, n, c = feat.shape
        r = self.up_ratio
        net = feat.repeat(1, r, 1)  # copy j occupies rows j*n..(j+1)*n-1
        grid = self.grid.view(1, r, 1, 2).expand(b, r, n, 2).reshape(b, r * n, 2)
        net = torch.cat([net, grid], dim=-1)
        net = self._fc(net, "up_conv1")
        return self._fc(net, "up_conv2")

    def regressor(self, feat, name, is_off=False):
        x = self._fc(self._fc(feat, f"{name}_fc0"), f"{name}_fc1")
        out = self._fc(x, f"{name}_fc2", act=False)
        if is_off:
            out = torch.sigmoid(out) * 1.0 - 0.5
        return out

    def point_shuffle(self, xyz, feature, nsample=16):
        """PointShuffle2 ops.py:1012-1087 with mlp=[128,128,256], kNN grouping, NL + Local."""
        b, m, c = feature.shape
        idx = self.ops.knn3d(nsample, xyz, xyz)  # ops.py:165 (nanoflann in the reference)
        hoist = hasattr(self.ops, "group_affine_relu") and self.hoist
        grouped_xyz_raw = self.ops.group(xyz, idx)
        grouped_xyz = grouped_xyz_raw - xyz.unsqueeze(2)
        if not hoist:
            grouped_feature = torch.cat([grouped_xyz_raw, self.ops.group(feature, idx)], dim=-1)  # 131
            grouped_feat = torch.cat([grouped_xyz, grouped_feature], dim=-1)  # 134
        # non-local cell
        kv = self._fc(feature, "nl_kv", act=False)
        key, val = kv[..., :64], kv[..., 64:]
        q = self._fc(feature, "nl_query", act=False)
        att = torch.softmax(torch.bmm(q, key.transpose(1, 2)) / math.sqrt(64.0), dim=-1)
        nonlocal_pt = self._fc(torch.bmm(att, val), "nl_back")
        if hoist:
            # grouped row = [x_j - x_i, x_j, f_j]; conv0 row.W + b = (x_j(Wa+Wb) + f_j Wc + b) - x_i Wa = P[j] - Q[i]:
            # one per-point GEMM, then the fused gather + centre term + ReLU kernel; the 131/134-channel grouped
            # tensors of ops.py:1033-1036 are never built.  max_s commutes with "- x_i" exactly.
            w0, b0 = self.P["conv0"]
            if "_hoist_ps" not in self.P:
                self.P["_hoist_ps"] = (torch.cat([w0[0:3] + w0[3:6], w0[6:]], dim=0).contiguous(), w0[0:3].contiguous())
            wp, wa = self.P["_hoist_ps"]
            xf = torch.cat([xyz, feature], dim=-1)  # (b,m,131) per point
            pm = self._lin("_hoist_ps_p", xf, wp, b0)
            qm = self._lin("_hoist_ps_q", xyz, wa)
            y0 = self.ops.group_affine_relu(pm, qm, idx, -1.0)  # relu(conv0): (b,m,s,128)
            gf = self._fc(y0, "conv1")
            mx = self.ops.group_max(xf, idx)  # [max_s x_j, max_s f_j]
            skip = self._fc(torch.cat([mx[..., :3] - xyz, mx], dim=-1), "skip")
        else:
            skip = self._fc(grouped_feat.max(dim=2)[0], "skip")
            gf = self._fc(self._fc(grouped_feat, "conv0"), "conv1")  # (b,m,s,128)
        w, bias = self.P["wconv0"]
        wn = torch.relu((torch.matmul(grouped_xyz, w) + bias) / math.sqrt(1.0 + BN_EPS))  # (b,m,s,16)
        gf = torch.matmul(gf.transpose(2, 3), wn)  # (b,m,128,16)
        gf = self._fc(gf.reshape(b, m, 128 * nsample), "after_conv")
        gf = gf + skip + nonlocal_pt
        return self._fc(gf, "aggregation")

    # -- fused forward ------------------------------------------------------------------------------
    def _prepare_fused(self):
        """Derived weights of the fused formulation (exact reassociations of the reference's 1x1 convs)."""
        if "_f_ready" in self.P:
            return
        P = self.P
        wq, bq = P["nl_query"]
        P["_f_query"] = ((wq * 0.125).contiguous(), (bq * 0.125).contiguous())  # softmax(q.k/sqrt(64)): 1/8 is exact
        w, b = P["up_conv1"]
        P["_f_up1"] = (w[:480].contiguous(), b)
        P["_f_grid"] = torch.matmul(self.grid, w[480:482]).contiguous()  # (r, 256): the grid columns of the tiled input
        w0, b0 = P["conv0"]  # rows: [x_j - x_i (3) | x_j (3) | f_j (128)]
        z1 = torch.zeros((1, w0.shape[1]), device=w0.device)
        P["_f_psp"] = (torch.cat([z1, w0[0:3] + w0[3:6], w0[6:]], dim=0).contiguous(), b0)  # on [0 | x | f] (132)
        P["_f_psq"] = (w0[0:3].contiguous(), None)
        ws, bs = P["skip"]  # rows as conv0; input = [max_s x_j - x_i | max_s x_j | max_s f_j]
        z = torch.zeros((1, ws.shape[1]), device=ws.device)
        P["_f_skip"] = (torch.cat([z, ws[0:3] + ws[3:6], ws[6:], -ws[0:3], z], dim=0).contiguous(), bs)  # on [0|mx|mf|x_i|0] (136)
        P["_f_ready"] = True

    def _buffers(self, b, n, dev):
        key = (b, n, str(dev))
        if key not in self._bufs:
            m = n * self.up_ratio
            self._bufs[key] = {
                "feat": torch.zeros((b, n, 480), dtype=torch.float32, device=dev),
                "xf": torch.zeros((b, m, 132), dtype=torch.float32, device=dev),   # [0 | coarse xyz | coarse_feat]
                "mxz": torch.zeros((b, m, 136), dtype=torch.float32, device=dev),  # [group_max(xf) | coarse xyz | 0]
            }
        return self._bufs[key]

    def _fx(self, x, name, act=True, **kw):
        """Dense layer `name` on the tcgen05 kernel with the extended options of dense.PackedLinear.__call__."""
        lin = self._packed.get(name)
        if lin is None:
            from .dense import PackedLinear
            w, b = self.P[name]
            lin = self._packed[name] = PackedLinear(w, b)
        return lin(x, relu=act, **kw)

    def _side(self, i, dev):
        key = (i, str(dev))
        st = self._streams.get(key)
        if st is None:
            st = self._streams[key] = torch.cuda.Stream(device=dev)
        return st

    def forward_fused(self, inputs, on_coarse=None):
        """`on_coarse(coarse)`: optional callback run on a SIDE stream as soon as the coarse cloud exists (e.g. its
        Chamfer loss), concurrently with PointShuffle2; it is joined before this function returns.
        Same function as forward() with every layer on this library's kernels: layers write into their slot of the
        tensors the reference builds with tf.concat, grouped layers gather inside the operand load, the EdgeConv
        dense blocks and the feature-space kNN are single kernels, the non-local attention is two batched tcgen05
        products around a softmax."""
        from . import ops as O
        from .dense import BatchedOperand
        self._prepare_fused()
        P = self.P
        b, n, _ = inputs.shape
        m = n * self.up_ratio
        dev = inputs.device
        buf = self._buffers(b, n, dev)
        feat, xf, mxz = buf["feat"], buf["xf"], buf["mxz"]
        # ---- feature_extraction_GCN (ops.py:1437-1486): feat = [blk4 | blk3 | blk2 | blk1 (96) | l0 (24)]
        l0 = feat[..., 456:480]
        from .dense import linear_small
        linear_small(inputs, *P["layer0"], relu=False, out=l0)  # 3 -> 24: a 3-term dot product stays on CUDA cores
        O.dense_conv(l0, [P[f"layer1_l{i}"] for i in range(3)], out=feat[..., 360:456])
        for blk, lo in ((2, 240), (3, 120), (4, 0)):
            prep = self._fx(feat[..., lo + 120:480], f"layer{blk}_prep")
            O.dense_conv(prep, [P[f"layer{blk}_l{i}"] for i in range(3)], out=feat[..., lo:lo + 120])
        # ---- duplicate_up (ops.py:1152-1199): the tiled features share one product; copy r adds its grid term
        h = self._fx(feat, "_f_up1", act=False)  # (b,n,256)
        # relu(h[i] + grid term of copy r) is exactly the gather front-end of the next layer (operand row (r, i) =
        # relu(h[i] + g[r])), so the tiled (b,4n,256) tensor is never built either
        key = (b, n, str(dev))
        if key not in self._dup_idx:
            self._dup_idx[key] = (torch.arange(n, dtype=torch.int32, device=dev).repeat(b, self.up_ratio).contiguous().view(b, self.up_ratio, n),
                                  P["_f_grid"].repeat(b, 1).contiguous().view(b, self.up_ratio, -1))
        dup_idx, dup_q = self._dup_idx[key]
        coarse_feat = xf[..., 4:132]
        self._fx(h, "up_conv2", gather=(dup_idx, dup_q, 1.0), out=coarse_feat.view(b, self.up_ratio, n, 128))
        # ---- coarse coordinate_regressor
        if self.chain:
            # fc0 -> fc1 -> [fc2 | fc2.Wa] in one launch: the 3 coordinates go straight to every place the next stage reads them
            # (the dense coarse tensor, columns 1:4 of xf, columns 132:135 of mxz), and the centre term of the hoisted conv0,
            # qm = coarse.Wa = h2.(W_fc2 Wa) + b_fc2 Wa (exact reassociation), leaves as the tail of the same last layer
            ch = self._packed.get("_chain_coarse")
            if ch is None:
                from .dense import ChainedMLP
                (w0_, b0_), (w1_, b1_), (w2_, b2_) = P["coarse_fc0"], P["coarse_fc1"], P["coarse_fc2"]
                wa = P["_f_psq"][0]
                w2q = torch.cat([w2_, (w2_.double() @ wa.double()).float()], dim=1).contiguous()
                b2q = torch.cat([b2_, (b2_.double() @ wa.double()).float()]).contiguous()
                ch = self._packed["_chain_coarse"] = ChainedMLP([(w0_, b0_, True), (w1_, b1_, True), (w2q, b2q, False)], head=3)
            coarse = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
            _, qm = ch(coarse_feat, head_out=[coarse, xf[..., 1:4], mxz[..., 132:135]])
        else:
            coarse = self._fx(self._fx(self._fx(coarse_feat, "coarse_fc0"), "coarse_fc1"), "coarse_fc2", act=False)
            xf[..., 1:4].copy_(coarse)
            qm = None
        # ---- PointShuffle2 (ops.py:1012-1087)
        main = torch.cuda.current_stream(dev)
        forks = []
        if self.use_streams:
            s_knn = self._side(0, dev)
            s_knn.wait_stream(main)
            with torch.cuda.stream(s_knn):
                idx = self.ops.knn3d(16, coarse, coarse)  # only conv1's gather and the skip branch need it
            forks.append(s_knn)
            if on_coarse is not None:
                s_cb = self._side(1, dev)
                s_cb.wait_stream(main)
                with torch.cuda.stream(s_cb):
                    on_coarse(coarse)
                forks.append(s_cb)
        else:
            idx = self.ops.knn3d(16, coarse, coarse)
            if on_coarse is not None:
                on_coarse(coarse)
        # non-local cell (PointNonLocalCell ops.py:302-346): q.k^T and att.v as batched tcgen05 products
        # key|value and query share their input [0 | xyz | coarse_feat]: ONE launch, two 128-column blocks (the 4 xyz
        # columns get zero weights, the query is zero-padded to 128).  The hoisted conv0 product pm keeps its own dense
        # (b,m,128) tensor: conv1 gathers its rows 16x each and a compact table stays L2-resident.
        # ... and the hoisted conv0 product pm = x_j(Wa+Wb) + f_j Wc + b shares that input too: it is a third column
        # block of the same launch, written to its OWN compact (b,m,128) tensor (conv1 gathers its rows 16x each, and a
        # compact table stays L2-resident)
        blk = self._packed.get("_f_kvq")
        if blk is None:
            from .dense import BlockedLinear
            wkv, bkv = P["nl_kv"]
            wq, bq = P["_f_query"]
            z4 = torch.zeros((4, 128), device=dev)
            wq128 = torch.cat([wq, torch.zeros((wq.shape[0], 64), device=dev)], dim=1)
            bq128 = torch.cat([bq, torch.zeros((64,), device=dev)])
            wp, bp = P["_f_psp"]
            blk = self._packed["_f_kvq"] = BlockedLinear([torch.cat([z4, wkv]), torch.cat([z4, wq128]), wp], [bkv, bq128, bp])
        kvq, pm = blk(xf, split=2)     # kvq (b,m,256): key | value | query (pre-scaled by 1/sqrt(64)) | 0 ; pm (b,m,128)
        q = kvq[..., 128:192]
        if m % 128:
            raise ValueError("forward_fused: 4*n must be a multiple of 128")
        if self.attention == "fused":
            # softmax(q.k^T).v as ONE tcgen05 kernel (csrc/attention_tc.cu): per 128-query tile two passes over the key
            # blocks (row maximum, then exp -> P ring in tensor memory -> O += P.V); the (b,m,m) score tensor of
            # Common/ops.py:330-336 is never written
            from .dense import attention_fused
            att_v = attention_fused(kvq)
        else:
            # two batched products with the softmax folded into them (round-1 form, kept for A/B: DISPU_ATTENTION=two_products)
            nb = m // 256
            if m % 256:
                raise ValueError("forward_fused (two_products attention): 4*n must be a multiple of 256")
            kop = BatchedOperand(b * nb, 64, 256, dev).pack(kvq, 0, 1, 256, 256 * 256)
            stats = torch.empty((b, m, 2 * nb), dtype=torch.float32, device=dev)
            scores = kop.matmul(q, n_blocks=nb, stats_out=stats)
            vop = BatchedOperand(b, m, 64, dev).pack(kvq, 64, 256, 1, m * 256)
            att_v = vop.matmul(scores, softmax_stats=stats)
        nonlocal_pt = self._fx(att_v, "nl_back")
        # local cell: conv0 hoisted to per-point products, gathered + ReLU'd inside conv1's operand load
        if qm is None:
            qm = self._fx(coarse, "_f_psq", act=False)   # (b,m,128) = x_i Wa
        # conv1 with the gather fused into its operand load AND the weight net + transposed product fused into its epilogue:
        # neither relu(conv0) nor conv1's (b,m,16,128) output is ever materialised.  (With the A operand in tensor memory the
        # fused epilogue no longer competes with the MMAs for shared-memory bandwidth: 0.23 ms vs 0.21 + 0.12 ms unfused;
        # DISPU_FUSE_WEIGHTNET=0 selects the two-kernel form.)
        ww, bw = P["wconv0"]
        if forks:
            main.wait_stream(forks[0])  # idx
        if self.fuse_weightnet:
            H = self._fx(pm, "conv1", gather=(idx, qm, -1.0), weightnet=(coarse, ww, bw, 1.0 / math.sqrt(1.0 + BN_EPS)))
        else:
            gf = self._fx(pm, "conv1", gather=(idx, qm, -1.0))  # (b,m,16,128)
            H = O.weightnet_aggregate(coarse, idx, gf, ww, bw, 1.0 / math.sqrt(1.0 + BN_EPS))  # (b,m,2048)
        # skip connection: max over the group commutes with the centre shift
        self.ops.group_max(xf, idx, out=mxz[..., 0:132])
        if not self.chain:
            mxz[..., 132:135].copy_(coarse)
        skip = self._fx(mxz, "_f_skip")
        g = self._fx(H, "after_conv", residuals=(skip, nonlocal_pt))
        fine_feat = self._fx(g, "aggregation")
        if self.chain:
            # fine regressor fc0 -> fc1 -> fc2 in one launch; its epilogue applies sigmoid(x) - 0.5 and adds the coarse cloud
            ch = self._packed.get("_chain_fine")
            if ch is None:
                from .dense import ChainedMLP
                ch = self._packed["_chain_fine"] = ChainedMLP([(*P["fine_fc0"], True), (*P["fine_fc1"], True), (*P["fine_fc2"], False)], head=3)
            fine, _ = ch(fine_feat, head_base=coarse)
        else:
            off = self._fx(self._fx(self._fx(fine_feat, "fine_fc0"), "fine_fc1"), "fine_fc2", act=False)
            fine = coarse + (torch.sigmoid(off) - 0.5)
        for st in forks[1:]:
            main.wait_stream(st)
        return coarse, fine

    def forward(self, inputs, on_coarse=None):
        fused = self.fused
        if fused is None:
            fused = self.tc and self.hoist and inputs.is_cuda and isinstance(self.ops, DispuOps)
        if fused:
            return self.forward_fused(inputs, on_coarse=on_coarse)
        if on_coarse is not None:
            raise ValueError("on_coarse is a feature of the fused forward")
        feat = self.feature_extraction(inputs)
        coarse_feat = self.duplicate_up(feat)
        coarse = self.regressor(coarse_feat, "coarse")
        fine_feat = self.point_shuffle(coarse, coarse_feat)
        fine = coarse + self.regressor(fine_feat, "fine", is_off=True)
        return coarse, fine

    __call__ = forward

    @staticmethod
    def flops_per_cloud(n_in=256, up_ratio=4, k=16):
        """Dense-math FLOPs of one forward per cloud (2*MACs), same accounting as SURVEY.md §8d."""
        n, m = n_in, n_in * up_ratio
        f = 2 * n * 3 * 24
        for blk in range(1, 5):
            c = 24 if blk == 1 else 48
            if blk > 1:
                f += 2 * n * 120 * (blk - 1) * 48
            f += 2 * n * n * c  # feature kNN matmul
            f += 2 * n * k * (2 * c * 24 + (24 + c) * 24 + (48 + c) * 24)
        f += 2 * m * (482 * 256 + 256 * 128)
        f += 2 * m * (128 * 256 + 256 * 64 + 64 * 3)
        f += 2 * m * (128 * 128 + 128 * 64 + 64 * 256) + 2 * 2 * m * m * 64
        f += 2 * m * 134 * 256
        f += 2 * m * k * (134 * 128 + 128 * 128 + 3 * 16) + 2 * m * 128 * k * k
        f += 2 * m * 2048 * 256 + 2 * m * 256 * 256
        f += 2 * m * (256 * 256 + 256 * 64 + 64 * 3)
        return f

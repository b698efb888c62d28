"""Dense layer on the tcgen05 tensor cores with FP32-level accuracy (3xTF32): y = act(x.W + b) [+ residuals].
The weight is packed once (TF32 hi/lo planes, UMMA layout); see dispu_dense_tf32x3(_ex) in include/dispu_b200.h."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._t import call, ptr, req, rows, stream_ptr


_SMS = {}


def _sm_count(device):
    key = str(device)
    if key not in _SMS:
        _SMS[key] = torch.cuda.get_device_properties(device).multi_processor_count
    return _SMS[key]


SPLIT_256 = os.environ.get("DISPU_SPLIT_256", "0") == "1"  # force the two-block form of every 256-column layer (see PackedLinear)


class PackedLinear:
    """W (K,N) float32 with N <= 256, optional bias (N,).  __call__(x (..., K), relu) -> (..., N).

    x may be a column slice of a wider buffer (x[..., a:a+K]); `out` may be one too — the layer then writes
    straight into its slot of a concatenated tensor.  `gather=(idx, Q, sign)`: the operand rows are
    relu(x[b, idx[b,i,s], :] + sign*Q[b,i,:]) (x is the per-point table (b,n,K), idx (b,m,ns), Q (b,m,K) or None),
    i.e. GroupPoint + centre term + ReLU fused into the load.  `residuals`: up to two (..., N) tensors added after
    the activation.  `weightnet=(xyz, Ww, bw, scale)` (with gather, ns = 16): the layer output G (b,m,16,N) is not
    written; the result is (b,m,N*16) = flatten(transpose(G) . relu(((xyz[idx]-xyz) . Ww + bw) * scale)), PointShuffle2's
    weight net + transposed product (Common/ops.py:1062-1067) fused into the epilogue."""

    def __init__(self, weight, bias=None):
        weight = req(weight, torch.float32, 2, "weight")
        self.K, self.N = weight.shape
        if self.N > 256:
            raise ValueError("PackedLinear supports N <= 256")
        self.bias = None if bias is None else req(bias, torch.float32, 1, "bias")
        # A 256-column layer can also run as two 128-column blocks (two TMEM accumulators fit, so the epilogue of one tile
        # overlaps the MMAs of the next, and there are twice as many work units).  That pays when the row tiles alone
        # cannot fill the SMs (M8192 K480 N256: 33 -> 25 us) and costs otherwise (the A operand is read twice, the MMAs are
        # smaller: after_conv 174 -> 203 us), so the form is chosen per call; DISPU_SPLIT_256=1 forces it everywhere.
        self._weight = weight
        self._packs = {}
        self._pack(1)

    def _pack(self, n_blocks):
        pk = self._packs.get(n_blocks)
        if pk is None:
            nblk = self.N // n_blocks
            nbytes = int(_lib.lib().dispu_dense_packed_bytes(self.K, nblk))
            stride = (nbytes + 15) // 16 * 16
            packed = torch.empty((n_blocks * stride,), dtype=torch.uint8, device=self._weight.device)
            # block j = columns [j*nblk, (j+1)*nblk): source element (k, n) at weight[k*N + j*nblk + n]
            call("dispu_dense_pack_strided", self.K, nblk, ptr(self._weight), self.N, 1, n_blocks, nblk, ptr(packed), stride, stream_ptr())
            pk = self._packs[n_blocks] = (packed, stride)
        return pk

    def __call__(self, x, relu=False, out=None, gather=None, residuals=(), weightnet=None):
        if x.shape[-1] != self.K:
            raise ValueError(f"PackedLinear expects last dim {self.K}, got {tuple(x.shape)}")
        a = _lib.DenseArgs()
        xp, xm, ldx = rows(x, "x")
        m_rows = xm if gather is None else gather[0].numel()
        n_blocks = 1
        if self.N == 256 and weightnet is None and (SPLIT_256 or 2 * ((m_rows + 127) // 128) <= _sm_count(x.device)):
            n_blocks = 2
        packed, block_stride = self._pack(n_blocks)
        a.K, a.N, a.X, a.ldx = self.K, self.N // n_blocks, xp, ldx
        a.packed, a.bias, a.relu = ptr(packed), ptr(self.bias), 1 if relu else 0
        a.n_blocks, a.packed_block_stride = n_blocks, block_stride
        if gather is not None:
            idx, Q, sign = gather
            idx = req(idx, torch.int32, 3, "idx")
            b, m, ns = idx.shape
            if x.dim() != 3 or x.shape[0] != b:
                raise ValueError("gather: x must be (b,n,K)")
            a.gather_idx, a.gather_n, a.gather_ns, a.gather_rows_per_cloud = ptr(idx), x.shape[1], ns, m * ns
            if Q is not None:
                qp, qm, ldq = rows(Q, "Q")
                if qm != b * m or Q.shape[-1] != self.K:
                    raise ValueError("gather: Q must be (b,m,K)")
                a.Q, a.ldq = qp, ldq
            a.sign = float(sign)
            M = b * m * ns
            oshape = (b, m, ns, self.N)
            if weightnet is not None:
                xyz, ww, bw, scale = weightnet
                xyz = req(xyz, torch.float32, 3, "xyz")
                ww = req(ww, torch.float32, 2, "Ww")
                if ns != 16 or tuple(ww.shape) != (3, 16) or tuple(xyz.shape) != (b, m, 3) or x.shape[1] != m:
                    raise ValueError("weightnet: needs ns == 16, Ww (3,16), xyz (b,m,3) and support == query points")
                self._wn_keep = (xyz, ww, bw)
                a.wn_xyz, a.wn_w, a.wn_b, a.wn_scale = ptr(xyz), ptr(ww), ptr(bw), float(scale)
                oshape = (b, m, self.N * 16)
        else:
            M = xm
            oshape = (*x.shape[:-1], self.N)
        a.M = M
        if out is None:
            out = torch.empty(oshape, dtype=torch.float32, device=x.device)
        yp, ym, ldy = rows(out, "out")
        if weightnet is not None:
            if ym * 16 != M or out.shape[-1] != self.N * 16:
                raise ValueError("weightnet: out must be (b,m,N*16)")
        elif ym != M or out.shape[-1] != self.N:
            raise ValueError(f"out: expected {M} rows of {self.N}, got {tuple(out.shape)}")
        a.Y, a.ldy = yp, ldy
        for i, r in enumerate(residuals):
            rp, rm, ldr = rows(r, "residual")
            if rm != M or r.shape[-1] != self.N:
                raise ValueError("residual shape mismatch")
            if i == 0:
                a.R1, a.ldr1 = rp, ldr
            elif i == 1:
                a.R2, a.ldr2 = rp, ldr
            else:
                raise ValueError("at most two residuals")
        call("dispu_dense_tf32x3_ex", C.byref(a), stream_ptr())
        return out


class BatchedOperand:
    """Per-batch right-hand operands W_i (K,N), N <= 256, packed every call from a strided source:
    W_i[k][n] = src[i*batch_stride + k*stride_k + n*stride_n] — e.g. k^T and v of PointNonLocalCell
    (Common/ops.py:326-339), which are activations, not weights."""

    def __init__(self, batch, K, N, device):
        self.batch, self.K, self.N = batch, K, N
        nb = int(_lib.lib().dispu_dense_packed_bytes(K, N))
        self.stride = (nb + 15) // 16 * 16
        self.packed = torch.empty((batch * self.stride,), dtype=torch.uint8, device=device)

    def pack(self, src, offset, stride_k, stride_n, batch_stride):
        """src: float32 CUDA tensor (storage base); offset in floats."""
        if src.dtype != torch.float32 or not src.is_cuda:
            raise ValueError("BatchedOperand.pack expects a float32 CUDA tensor")
        call("dispu_dense_pack_strided", self.K, self.N, src.data_ptr() + 4 * offset, stride_k, stride_n, self.batch,
             batch_stride, ptr(self.packed), self.stride, stream_ptr())
        return self

    def matmul(self, x, out=None, relu=False, n_blocks=1, stats_out=None, softmax_stats=None):
        """x (B, rows, K) [row-strided ok] -> out (B, rows, n_blocks*N); rows % 128 == 0.  With n_blocks > 1 this object
        holds B*n_blocks operands, operand (i*n_blocks + j) producing columns [j*N, (j+1)*N) of cloud i.
        stats_out: (B, rows, 2*n_blocks) float32 tensor that receives per-row partial maxima of the written values;
        softmax_stats: such a tensor for x — the product is then softmax(x, -1) . W without a softmax pass."""
        a = _lib.DenseArgs()
        xp, xm, ldx = rows(x, "x")
        if x.dim() != 3 or x.shape[0] * n_blocks != self.batch or x.shape[2] != self.K:
            raise ValueError("BatchedOperand.matmul: x must be (batch / n_blocks, rows, K)")
        if out is None:
            out = torch.empty((x.shape[0], x.shape[1], n_blocks * self.N), dtype=torch.float32, device=x.device)
        yp, ym, ldy = rows(out, "out")
        if ym != xm or out.shape[-1] != n_blocks * self.N:
            raise ValueError("BatchedOperand.matmul: out shape mismatch")
        a.M, a.K, a.N, a.X, a.ldx = xm, self.K, self.N, xp, ldx
        a.packed, a.relu, a.Y, a.ldy = ptr(self.packed), 1 if relu else 0, yp, ldy
        a.rows_per_batch, a.packed_batch_stride = x.shape[1], n_blocks * self.stride
        a.n_blocks, a.packed_block_stride = n_blocks, self.stride
        if stats_out is not None:
            if not stats_out.is_contiguous() or stats_out.numel() != xm * 2 * n_blocks or stats_out.dtype != torch.float32:
                raise ValueError("stats_out must be a contiguous float32 (B, rows, 2*n_blocks) tensor")
            a.stat_out = ptr(stats_out)
        if softmax_stats is not None:
            if not softmax_stats.is_contiguous() or softmax_stats.dtype != torch.float32 or softmax_stats.numel() % xm:
                raise ValueError("softmax_stats must be a contiguous float32 (B, rows, n_stat) tensor")
            a.stat_in, a.n_stat = ptr(softmax_stats), softmax_stats.numel() // xm
        call("dispu_dense_tf32x3_ex", C.byref(a), stream_ptr())
        return out


class BlockedLinear:
    """Several layers that share their input, run as ONE launch: weights[j] (K, N) with a common N <= 256 (pad narrower
    layers with zero columns), optional biases[j] (N,).  __call__(x (..., K)) -> (..., len(weights)*N), block j in
    columns [j*N, (j+1)*N) — e.g. PointShuffle2's key|value, query and hoisted conv0 products of the same features."""

    def __init__(self, weights, biases=None):
        ws = [req(w, torch.float32, 2, "weight") for w in weights]
        self.nb = len(ws)
        self.K, self.N = ws[0].shape
        if any(tuple(w.shape) != (self.K, self.N) for w in ws) or self.N > 256:
            raise ValueError("BlockedLinear: weights must share one (K, N <= 256) shape")
        stacked = torch.stack(ws).contiguous()
        nb_bytes = int(_lib.lib().dispu_dense_packed_bytes(self.K, self.N))
        self.stride = (nb_bytes + 15) // 16 * 16
        self.packed = torch.empty((self.nb * self.stride,), dtype=torch.uint8, device=stacked.device)
        call("dispu_dense_pack_strided", self.K, self.N, ptr(stacked), self.N, 1, self.nb, self.K * self.N, ptr(self.packed),
             self.stride, stream_ptr())
        self.bias = None
        if biases is not None and any(b is not None for b in biases):
            self.bias = torch.cat([torch.zeros((self.N,), device=stacked.device) if b is None else b.float() for b in biases]).contiguous()

    def __call__(self, x, relu=False, out=None, split=None, out2=None):
        """split = j: blocks [0, j) go to `out` (..., j*N) and blocks [j, nb) to a second tensor `out2` (..., (nb-j)*N);
        returns (out, out2) then."""
        if x.shape[-1] != self.K:
            raise ValueError(f"BlockedLinear expects last dim {self.K}")
        a = _lib.DenseArgs()
        xp, xm, ldx = rows(x, "x")
        n1 = self.nb if split is None else split
        if not (0 < n1 <= self.nb) or (split is not None and split == self.nb):
            raise ValueError("BlockedLinear: split must be in 1..nb-1")
        if out is None:
            out = torch.empty((*x.shape[:-1], n1 * self.N), dtype=torch.float32, device=x.device)
        yp, ym, ldy = rows(out, "out")
        if ym != xm or out.shape[-1] != n1 * self.N:
            raise ValueError("BlockedLinear: out shape mismatch")
        a.M, a.K, a.N, a.X, a.ldx = xm, self.K, self.N, xp, ldx
        a.packed, a.bias, a.relu, a.Y, a.ldy = ptr(self.packed), ptr(self.bias), 1 if relu else 0, yp, ldy
        a.n_blocks, a.packed_block_stride = self.nb, self.stride
        if split is not None:
            if out2 is None:
                out2 = torch.empty((*x.shape[:-1], (self.nb - n1) * self.N), dtype=torch.float32, device=x.device)
            y2p, y2m, ldy2 = rows(out2, "out2")
            if y2m != xm or out2.shape[-1] != (self.nb - n1) * self.N:
                raise ValueError("BlockedLinear: out2 shape mismatch")
            a.Y2, a.ldy2, a.y2_block = y2p, ldy2, n1
        call("dispu_dense_tf32x3_ex", C.byref(a), stream_ptr())
        return out if split is None else (out, out2)


def attention_fused(kvq, out=None):
    """PointNonLocalCell's softmax(q.k^T).v (Common/ops.py:326-339) in ONE tcgen05 kernel (csrc/attention_tc.cu,
    dispu_attention_tf32x3): kvq (b,m,256) = [key 64 | value 64 | query 64 pre-scaled by 1/sqrt(64) | pad] -> (b,m,64);
    the (b,m,m) score tensor is never written.  m % 128 == 0."""
    if kvq.dim() != 3 or kvq.shape[2] != 256 or not kvq.is_contiguous() or kvq.shape[1] % 128:
        raise ValueError("attention_fused expects a contiguous (b, m, 256) tensor with m % 128 == 0")
    b, m, _ = kvq.shape
    nb = m // 128
    key = (b, m, str(kvq.device))
    bufs = _attn_operands.get(key)
    if bufs is None:  # packed K / V operands: 64 KB per (cloud, 128-key block) each
        bufs = _attn_operands[key] = (torch.empty((b * nb * 65536,), dtype=torch.uint8, device=kvq.device),
                                      torch.empty((b * nb * 65536,), dtype=torch.uint8, device=kvq.device))
    kp, vp = bufs
    if os.environ.get("DISPU_ATTENTION_PACK", "fused") == "fused":
        call("dispu_attention_pack", b, m, ptr(kvq), 256, ptr(kp), ptr(vp), stream_ptr())
    else:  # the generic strided pack kernel, twice (kept for A/B)
        call("dispu_dense_pack_strided", 64, 128, ptr(kvq), 1, 256, b * nb, 128 * 256, ptr(kp), 65536, stream_ptr())
        call("dispu_dense_pack_strided", 128, 64, kvq.data_ptr() + 4 * 64, 256, 1, b * nb, 128 * 256, ptr(vp), 65536, stream_ptr())
    if out is None:
        out = torch.empty((b, m, 64), dtype=torch.float32, device=kvq.device)
    call("dispu_attention_tf32x3", b, m, kvq.data_ptr() + 4 * 128, 256, ptr(kp), ptr(vp), ptr(out), 64, stream_ptr())
    return out


_attn_operands = {}


def linear_small(x, w, bias=None, relu=False, out=None):
    """y = act(x . W + b) for K <= 8 input channels on CUDA cores (dispu_linear_small): the generator's 3 -> 24 layer0."""
    w = req(w, torch.float32, 2, "weight")
    K, N = w.shape
    if x.shape[-1] != K:
        raise ValueError(f"linear_small expects last dim {K}")
    xp, xm, ldx = rows(x, "x")
    if out is None:
        out = torch.empty((*x.shape[:-1], N), dtype=torch.float32, device=x.device)
    yp, ym, ldy = rows(out, "out")
    if ym != xm or out.shape[-1] != N:
        raise ValueError("linear_small: out shape mismatch")
    call("dispu_linear_small", xm, K, N, xp, ldx, ptr(w), ptr(bias), 1 if relu else 0, yp, ldy, stream_ptr())
    return out


class ChainedMLP:
    """Two or three dense layers as ONE launch (csrc/mlp_chain_tc.cu, dispu_mlp_chain_tf32x3): the intermediate
    activations stay in tensor memory.  layers: [(W (K,N), bias | None, relu: bool), ...] with N1 <= 256, N2 <= 64,
    N3 <= 256.  `head` (1..4): the LAST layer's first `head` output columns are the "head" (e.g. the 3 coordinates of the
    regressor, Common/ops.py:1089-1110); the remaining columns of the last layer, if any, are the "tail".
    __call__(x, head_out=[tensors (..., head)], head_base=None | (..., head) tensor, tail_out=None) -> (head tensor, tail
    tensor | None); with head_base the head is head_base + sigmoid(x) - 0.5 (the offset regressor + generator.py:78-81)."""

    def __init__(self, layers, head=0):
        if len(layers) not in (2, 3):
            raise ValueError("ChainedMLP takes two or three layers")
        self.L = len(layers)
        self.head = head
        dev = layers[0][0].device
        self.packed, self.bias, self.relu, self.N = [], [], [], []
        k_prev = None
        for i, (w, bb, relu) in enumerate(layers):
            w = req(w, torch.float32, 2, f"weight[{i}]")
            K, N = w.shape
            if k_prev is not None and K != k_prev:
                raise ValueError(f"layer {i}: K={K} does not match the previous layer's width {k_prev}")
            bb = torch.zeros((N,), dtype=torch.float32, device=dev) if bb is None else req(bb, torch.float32, 1, f"bias[{i}]")
            if i == self.L - 1 and head > 0:
                # last layer laid out as [16-column head block | tail]: head columns first, zero padding up to 16
                if not (0 < head <= 4) or head > N:
                    raise ValueError("head must be in 1..4 and <= the last layer's width")
                pad = torch.zeros((K, 16 - head), dtype=torch.float32, device=dev)
                w = torch.cat([w[:, :head], pad, w[:, head:]], dim=1).contiguous()
                bb = torch.cat([bb[:head], torch.zeros((16 - head,), device=dev), bb[head:]]).contiguous()
                self.tail = N - head
                N = w.shape[1]
            elif i == self.L - 1:
                self.tail = N
            nbytes = int(_lib.lib().dispu_dense_packed_bytes(K, N))
            if nbytes == 0:
                raise ValueError("ChainedMLP: layer wider than 256 columns")
            pk = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            call("dispu_dense_pack", K, N, ptr(w), ptr(pk), stream_ptr())
            self.packed.append(pk)
            self.bias.append(bb)
            self.relu.append(1 if relu else 0)
            self.N.append(N)
            k_prev = layers[i][0].shape[1]
        self.K0 = layers[0][0].shape[0]

    def __call__(self, x, head_out=(), head_base=None, tail_out=None):
        if x.shape[-1] != self.K0:
            raise ValueError(f"ChainedMLP expects last dim {self.K0}, got {tuple(x.shape)}")
        a = _lib.MlpChainArgs()
        xp, xm, ldx = rows(x, "x")
        a.M, a.K0, a.X, a.ldx, a.n_layers, a.head = xm, self.K0, xp, ldx, self.L, self.head
        for l in range(self.L):
            a.N[l], a.W[l], a.bias[l], a.relu[l] = self.N[l], ptr(self.packed[l]), ptr(self.bias[l]), self.relu[l]
        head_t = None
        if self.head > 0:
            outs = list(head_out)
            if not outs:
                outs = [torch.empty((*x.shape[:-1], self.head), dtype=torch.float32, device=x.device)]
            if len(outs) > 3:
                raise ValueError("at most three head destinations")
            for d, o in enumerate(outs):
                op, om, ldo = rows(o, "head_out")
                if om != xm or o.shape[-1] != self.head:
                    raise ValueError("head_out: expected (..., head) tensors with one row per input row")
                a.Yh[d], a.ldh[d] = op, ldo
            head_t = outs[0]
            if head_base is not None:
                bp, bm, ldb = rows(head_base, "head_base")
                if bm != xm or head_base.shape[-1] != self.head:
                    raise ValueError("head_base shape mismatch")
                a.head_base, a.ldb = bp, ldb
        tail_t = None
        if self.tail > 0:
            tail_t = tail_out if tail_out is not None else torch.empty((*x.shape[:-1], self.tail), dtype=torch.float32, device=x.device)
            tp, tm, ldt = rows(tail_t, "tail_out")
            if tm != xm or tail_t.shape[-1] != self.tail:
                raise ValueError("tail_out shape mismatch")
            a.Yt, a.ldt = tp, ldt
        call("dispu_mlp_chain_tf32x3", C.byref(a), stream_ptr())
        return head_t, tail_t

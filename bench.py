#!/usr/bin/env python
"""bench.py — headline benchmark of the Dis-PU hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config selects the workload (default c2 = the configuration BASELINE.json's metric is quoted on; the others are the
remaining GPU configs of BASELINE.json, same JSON schema): c3 = B=128 sample_and_group + grouped MLP on tcgen05 (two
set-abstraction levels), c4 = two-stage 16x generator (256 -> 1024 -> 4096) + Chamfer, batch-sharded, c5 = FPS +
ball-query sweep N in {1k,4k,16k,64k}, B=64 over all ranks.

Workload (BASELINE.json configs[1], "c2"): B=32 patches of 256 points per GPU -> 4x upsampling generator
forward (1024 points) + Chamfer/Hausdorff of coarse and fine against a 1024-point ground truth
(DisPU/model.py:74-79).  Metric: input points / second, whole job.  One step = one such batch.

 * `value`   : device-timed (CUDA events on the launching stream), inputs resident in HBM, the step
               replayed from a CUDA graph; L2 flushed between timed steps.
 * `e2e`     : same step through the public API with HOST inputs: pinned host -> device copy of the
               patch batch + ground truth and device -> host read of the loss scalars inside the timed
               region, every step.
 * `roofline`: the dominant kernel of this library inside the step, timed live with CUDA events.
 * `cpu_baseline` / `--impl reference`: the same workload on the host cores: the torch restatement of
               the generator on CPU threads with the REFERENCE's own kNN (nanoflann,
               libs/nearest_neighbors/knn_.cxx, OpenMP over the batch) and NnDistance host loop
               (tf_nndistance.cpp:21-43) from oracle/_ref, on a bounded sample of the batch.
Multi-GPU: the batch is sharded (B per rank, weak scaling), no data-path collective except ONE
all-reduce of the per-rank Chamfer partial sums.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 32
N_IN = 256
UP = 4
RESULT_OUT = sys.stdout  # main() replaces it with a private copy of the real stdout
METRIC = "input_points_per_sec_4x_upsample_fwd_chamfer"
UNIT = "points/s"


# ------------------------------------------------------------------------------------------------
def synth_batch(rank, b=B_PER_GPU, n_in=N_IN, up=UP):
    """gt = 1024 unit-ball points per patch, input = a 256-subset; seed 2000+rank."""
    rng = np.random.default_rng(2000 + rank)
    n_gt = n_in * up
    gt = np.empty((b, n_gt, 3), np.float32)
    for i in range(b):
        got = 0
        while got < n_gt:
            c = rng.uniform(-1, 1, (3 * n_gt, 3)).astype(np.float32)
            c = c[(c.astype(np.float64) ** 2).sum(1) < 1.0]
            k = min(len(c), n_gt - got)
            gt[i, got:got + k] = c[:k]
            got += k
    sel = np.stack([rng.permutation(n_gt)[:n_in] for _ in range(b)])
    inp = np.take_along_axis(gt, sel[:, :, None], axis=1)
    return np.ascontiguousarray(inp), gt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (reference host code + torch-on-CPU restatement of the stock dense layers)
class CpuRefOps:
    """Backend of dispu_b200.generator.Generator for the CPU arm: the reference's own nanoflann kNN
    (Common/ops.py:110-118 -> knn_.cxx:104) and stock gathers."""

    def __init__(self):
        from oracle import ref as R
        self.R = R
        self.use_ref = R.available("knn")
        if not self.use_ref:
            from oracle import oracle as O
            self.O = O

    def knn3d(self, k, support, query):
        import torch
        s, q = support.numpy(), query.numpy()
        idx = self.R.cpu_knn_batch(s, q, k, omp=True) if self.use_ref else self.O.knn(k, s, q)[1]
        return torch.from_numpy(idx.astype(np.int64))

    def group(self, feature, idx):
        import torch
        b = feature.shape[0]
        return feature[torch.arange(b).view(b, 1, 1), idx.long()]

    def feature_knn(self, k, feat):
        import torch
        r = (feat * feat).sum(dim=2, keepdim=True)
        d = r - 2.0 * torch.bmm(feat, feat.transpose(1, 2)) + r.transpose(1, 2)
        return torch.topk(-d, k, dim=2, sorted=True)[1]


def cpu_chamfer(pred, gt):
    """chamfer()+hausdorff_loss() on the reference NnDistance host loop (single thread, as the TF CPU op)."""
    from oracle import ref as R
    if R.available("cpu"):
        df, _ = R.cpu_nnsearch(gt, pred)
        db, _ = R.cpu_nnsearch(pred, gt)
    else:
        from oracle import oracle as O
        df, _, db, _ = O.nn_distance(gt, pred, mode=O.CPU)
    cd = float((df.mean(1) + db.mean(1)).mean())
    hd = float((df.max(1) + db.max(1)).max())
    return cd, hd


def _time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def cpu_arm_generator(steps, warmup, sample_b, stages=1):
    """c2 / c4 on the host: the torch restatement of the generator on CPU threads with the REFERENCE's own kNN
    (nanoflann, OpenMP over the batch) and NnDistance host loop, on `sample_b` clouds per step."""
    import torch
    from oracle import ref as R
    from dispu_b200.generator import Generator
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp, gt = synth_batch(0)
    inp = inp[:sample_b]
    gt = gt[:sample_b] if stages == 1 else synth_gt(4000, sample_b, N_IN * UP * UP)
    gen = Generator(device="cpu", ops=CpuRefOps())
    x = torch.from_numpy(inp)

    def step():
        with torch.no_grad():
            coarse, fine = gen(x)
            if stages == 2:
                _, fine = gen(fine.contiguous())
        if stages == 1:
            cpu_chamfer(coarse.numpy(), gt)
        cpu_chamfer(fine.numpy(), gt)

    dt = _time_cpu(step, steps, warmup)
    kind = "reference" if (R.available("knn") and R.available("cpu")) else "port"
    sample = (f"{sample_b} of {B_PER_GPU} patches per step x {steps} steps; reference nanoflann kNN (OpenMP, {cores} threads) "
              f"+ reference nnsearch NnDistance (1 thread, as the TF CPU op) from oracle/_ref; stock dense layers = torch CPU "
              f"({cores} threads) restatement of generator.py (TF is not installable here)")
    return sample_b * N_IN / dt, dt * 1e3, cores, kind, sample


def synth_gt(seed, b, n):
    rng = np.random.default_rng(seed)
    gt = np.empty((b, n, 3), np.float32)
    for i in range(b):
        got = 0
        while got < n:
            c = rng.uniform(-1, 1, (3 * n, 3)).astype(np.float32)
            c = c[(c.astype(np.float64) ** 2).sum(1) < 1.0]
            k = min(len(c), n - got)
            gt[i, got:got + k] = c[:k]
            got += k
    return gt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, 1965.0, "fallback (B200_PROFILING.md)"


def fp32_peak_tflops(sm_mhz):
    """148 SMs x 128 FP32 FMA lanes x 2 flop at the maximum SM clock (the bound of the distance-test kernels)."""
    return 148 * 128 * 2 * sm_mhz * 1e6 / 1e12


# ------------------------------------------------------------------------------------------------
# Workloads: one class per BASELINE.json config.  A workload owns its pinned host inputs, the matching device buffers,
# a graph-capturable step_body() and what it costs algorithmically.
class Workload:
    metric = METRIC
    unit = UNIT
    scaling = "weak"

    def __init__(self, args, rank, world, dev):
        self.args, self.rank, self.world, self.dev = args, rank, world, dev
        self.host = []      # pinned host inputs, copied every end-to-end step
        self.device = []    # their device copies
        self.h_result = None  # pinned host result buffer, read every end-to-end step
        self.d_result = None
        self.units_per_step = 0   # metric units THIS rank processes per step
        self.collective = None    # device tensor the ONE data-path collective sums (None: no collective)
        self.gen = None

    def pin(self, arr):
        import torch
        h = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        d = torch.empty_like(h, device=self.dev)
        d.copy_(h)
        self.host.append(h)
        self.device.append(d)
        return d

    def step_body(self):
        raise NotImplementedError

    def profile_body(self):
        """The step on ONE stream (the per-call event pairs of the roofline leg must bracket one kernel each)."""
        if self.gen is not None:
            saved, self.gen.use_streams = self.gen.use_streams, False
            try:
                self.step_body_no_collective()
            finally:
                self.gen.use_streams = saved
        else:
            self.step_body_no_collective()

    def step_body_no_collective(self):
        self.step_body()

    def alg(self, key):
        """('flops' | 'bytes', amount per launch) of a profiled C-ABI call, or None."""
        return None

    def config(self):
        return {}

    def extra(self):
        return {}

    def cpu_arm(self, steps, warmup):
        raise NotImplementedError


def dense_flops(key):
    f = key.split(" ")
    M_, K_, N_ = int(f[1][1:]), int(f[2][1:]), int(f[3][1:])
    fl = 2.0 * M_ * K_ * N_
    for tok in f[4:]:
        if tok.startswith("x") and tok.endswith("blocks"):
            fl *= int(tok[1:-6])  # column blocks: N per block
        if tok == "+weightnet":
            fl += 2.0 * M_ * N_ * 16  # transposed product with the weight net in the epilogue
    return fl


class GeneratorWorkload(Workload):
    """c2 (stages=1) and c4 (stages=2): generator forward + fused Chamfer/Hausdorff, batch-sharded, one all-reduce."""

    def __init__(self, args, rank, world, dev, stages):
        super().__init__(args, rank, world, dev)
        import torch
        from dispu_b200.generator import Generator
        self.stages = stages
        self.B = B_PER_GPU
        inp_np, gt_np = synth_batch(rank)
        if stages == 2:  # gt (B,4096,3), seed 4000+rank
            gt_np = synth_gt(4000 + rank, self.B, N_IN * UP * UP)
        self.d_inp = self.pin(inp_np)
        self.d_gt = self.pin(gt_np)
        self.gen = Generator(device=dev)
        self.out_c = torch.zeros((4,), dtype=torch.float32, device=dev)
        self.out_f = torch.zeros((4,), dtype=torch.float32, device=dev)
        # [sum cd_coarse, sum cd_fine | hd_coarse, hd_fine]: the fused Chamfer kernels deposit their partials here themselves
        # (dispu_chamfer_ex); the first two floats are what the ONE collective of the path sums over the ranks
        self.partial = torch.zeros((4,), dtype=torch.float32, device=dev)
        self.d_result = self.partial
        self.h_result = torch.zeros((4,), dtype=torch.float32).pin_memory()
        self.collective = self.partial[:2]
        self.units_per_step = self.B * N_IN
        self.name = "c2" if stages == 1 else "c4"

    def step_body_no_collective(self):
        import torch
        from dispu_b200 import loss_utils
        p = self.partial
        with torch.no_grad():
            if self.stages == 1:
                # the coarse cloud's loss is handed to the generator as a hook: with DISPU_STREAMS=1 it runs on a side stream
                coarse, fine = self.gen(self.d_inp, on_coarse=lambda c: loss_utils.chamfer_hausdorff(
                    c, self.d_gt, out=self.out_c, sum_out=p[0:1], max_out=p[2:3]))
            else:  # DisPU/model.py:116-118: G(G(x).fine), shared weights; the loss is on the final 16x cloud
                _, fine1 = self.gen(self.d_inp)
                _, fine = self.gen(fine1)
            loss_utils.chamfer_hausdorff(fine, self.d_gt, out=self.out_f, sum_out=p[1:2], max_out=p[3:4])

    def step_body(self):
        self.step_body_no_collective()

    def alg(self, key):
        nm = key.split(" ")[0]
        if nm == "dispu_dense_tf32x3_ex":
            return ("flops", dense_flops(key))
        ints = [int(x) for x in key.split(" ")[1].split(",")] if " " in key and key.split(" ")[1][0].isdigit() else []
        if nm == "dispu_attention_tf32x3":
            b, m = ints[0], ints[1]
            return ("flops", 4.0 * b * m * m * 64)
        if nm == "dispu_knn":
            b, n, m, k = ints[:4]
            return ("bytes", b * (12 * n + 12 * m + 4 * m * k))
        if nm in ("dispu_chamfer", "dispu_chamfer_ex"):
            b, n, m = ints[0], ints[1], ints[2]
            return ("bytes", b * (12 * n + 12 * m) + 16)
        if nm == "dispu_group_max_ld":
            b, n, c, m, ns = ints[:5]
            return ("bytes", b * (4 * m * ns + 4 * n * c + 4 * m * c))
        if nm in ("dispu_dense_conv", "dispu_dense_conv_knn"):
            b, n, c, k = ints[0], ints[1], ints[2], ints[3]
            return ("bytes", b * n * (4 * c + 4 * k + 4 * (72 + c)))
        return None

    def config(self):
        w = ("c2: B=32/GPU, N=256->1024 (4x) generator forward + Chamfer/Hausdorff(coarse,fine)" if self.stages == 1 else
             "c4: B=32/GPU (256 over 8 GPUs), two-stage 16x generator 256->1024->4096 (shared weights) + Chamfer/Hausdorff vs 4096-point gt")
        return {"workload": w, "batch_per_gpu": self.B, "n_in": N_IN, "up_ratio": UP if self.stages == 1 else UP * UP,
                "weights": "Xavier-uniform seed 7, zero bias (random init; no checkpoint offline)",
                "attention": self.gen.attention,
                "streams": "kNN of the coarse cloud and the coarse Chamfer on side streams (graph fork/join)" if self.gen.use_streams else "single stream"}

    def extra(self):
        loss = self.h_result.tolist()
        return {"loss": {"cd_coarse_sum": loss[0], "cd_fine_sum": loss[1], "hd_coarse": loss[2], "hd_fine": loss[3]}}

    def cpu_arm(self, steps, warmup):
        return cpu_arm_generator(steps, warmup, sample_b=self.args.ref_sample if self.stages == 1 else max(1, self.args.ref_sample // 4),
                                 stages=self.stages)


class SetAbstractionWorkload(Workload):
    """c3: B=128 clouds of 4096 points per GPU -> SA level 1 (npoint 1024, r 0.2, ns 32, mlp [32,32,64]) -> SA level 2
    (npoint 512, r 0.4, ns 32, mlp [64,64,128] on 3+64 channels): FPS, gather, ball query and the fused group + MLP + max
    on tcgen05 (3xTF32), the shapes of Common/ops.py:564-570.  No collective (independent clouds)."""
    name = "c3"
    metric = "input_points_per_sec_sample_and_group_grouped_mlp"

    def __init__(self, args, rank, world, dev):
        super().__init__(args, rank, world, dev)
        import torch
        self.B = 128
        rng = np.random.default_rng(3000 + rank)
        self.d_xyz = self.pin(synth_gt(3000 + rank, self.B, 4096))
        g = torch.Generator().manual_seed(11)

        def layers(widths):
            return [((torch.randn((widths[i], widths[i + 1]), generator=g) / widths[i] ** 0.5).to(dev), torch.zeros(widths[i + 1], device=dev))
                    for i in range(len(widths) - 1)]
        self.l1, self.l2 = layers([3, 32, 32, 64]), layers([67, 64, 64, 128])
        self.units_per_step = self.B * 4096
        self.d_result = torch.zeros((self.B, 512, 128), dtype=torch.float32, device=dev)
        self.h_result = torch.zeros((self.B, 512, 128), dtype=torch.float32).pin_memory()
        del rng

    def step_body(self):
        import torch
        import dispu_b200.pointnet_util as PU
        with torch.no_grad():
            xyz1, f1, _ = PU.pointnet_sa_module(self.d_xyz, None, 1024, 0.2, 32, self.l1, precision=PU.MLP_TF32X3)
            _, f2, _ = PU.pointnet_sa_module(xyz1, f1, 512, 0.4, 32, self.l2, precision=PU.MLP_TF32X3)
            self.d_result.copy_(f2)

    def alg(self, key):
        nm = key.split(" ")[0]
        ints = [int(x) for x in key.split(" ")[1].split(",")] if " " in key else []
        if nm == "dispu_sa_mlp":
            b, n, m, ns, c_pts = ints[:5]
            w = [3, 32, 32, 64] if c_pts == 0 else [67, 64, 64, 128]
            return ("flops", 2.0 * b * m * ns * sum(w[i] * w[i + 1] for i in range(3)))
        if nm == "dispu_fps":
            b, n, m = ints[:3]
            return ("bytes", b * (m - 1) * n * 16)
        if nm in ("dispu_query_ball_point", "dispu_query_ball_point_ws"):
            b, n, m = ints[:3]
            return ("bytes", b * (12 * n + 12 * m + 4 * m * 32 + 4 * m))
        if nm == "dispu_gather_point":
            b, n, m = ints[:3]
            return ("bytes", b * (4 * m + 24 * m))
        return None

    def config(self):
        return {"workload": "c3: B=128/GPU, 4096 -> npoint 1024 (r 0.2, ns 32, mlp [32,32,64]) -> npoint 512 (r 0.4, ns 32, mlp [64,64,128]): "
                            "FPS + gather + ball query + fused group/MLP/max on tcgen05 (3xTF32)",
                "batch_per_gpu": self.B, "n_in": 4096, "weights": "N(0, 1/sqrt(C_in)) seed 11, zero bias"}

    def cpu_arm(self, steps, warmup):
        return self.cpu_arm_static(steps, warmup)

    @staticmethod
    def cpu_arm_static(steps, warmup):
        """Reference host loops (query_ball_point_cpu + group_point_cpu from oracle/_ref) around the literal FPS restatement
        (the reference has NO CPU FPS kernel) and torch-CPU matmuls for the 1x1 convs, on 2 clouds."""
        import torch
        from oracle import oracle as O
        from oracle import ref as R
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sb, B = 2, 128
        xyz = synth_gt(3000, sb, 4096)
        g = torch.Generator().manual_seed(11)

        def mk(widths):
            return [((torch.randn((widths[i], widths[i + 1]), generator=g) / widths[i] ** 0.5), torch.zeros(widths[i + 1])) for i in range(len(widths) - 1)]
        l1, l2 = mk([3, 32, 32, 64]), mk([67, 64, 64, 128])
        use_ref = R.available("cpu")

        def sa(x, pts, npoint, r, ns, layers):
            new_xyz = O.gather_point(x, O.farthest_point_sample(npoint, x))
            idx = R.cpu_query_ball_point(r, ns, x, new_xyz) if use_ref else O.query_ball_point(r, ns, x, new_xyz, mode=O.CPU)[0]
            idx = np.maximum(idx, 0)
            g_ = (R.cpu_group_point(x, idx) if use_ref else O.group_point(x, idx)) - new_xyz[:, :, None, :]
            if pts is not None:
                g_ = np.concatenate([g_, R.cpu_group_point(pts, idx) if use_ref else O.group_point(pts, idx)], -1)
            h = torch.from_numpy(g_)
            for w, b in layers:
                h = torch.relu(h @ w + b)
            return new_xyz, h.max(dim=2)[0].numpy()

        def step():
            x1, f1 = sa(xyz, None, 1024, 0.2, 32, l1)
            sa(x1, f1, 512, 0.4, 32, l2)

        dt = _time_cpu(step, steps, warmup)
        kind = "reference" if use_ref else "port"
        sample = (f"{sb} of {B} clouds per step x {steps} steps; reference query_ball_point_cpu / group_point_cpu host loops (oracle/_ref), "
                  f"literal single-thread FPS restatement (the reference has no CPU FPS), torch CPU ({cores} threads) 1x1 convs")
        return sb * 4096 / dt, dt * 1e3, cores, kind, sample


class SweepWorkload(Workload):
    """c5: FPS + ball query, N in {1k,4k,16k,64k}, B=64 clouds over ALL ranks (strong scaling: 64 / world per rank),
    FPS m=N/4, ball query of those centres into the N points, ns=32, r=(32/N)^(1/3).  No collective."""
    name = "c5"
    metric = "input_points_per_sec_fps_ball_query_sweep"
    scaling = "strong"
    SIZES = (1024, 4096, 16384, 65536)

    def __init__(self, args, rank, world, dev):
        super().__init__(args, rank, world, dev)
        import torch
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import unit_ball
        if 64 % world:
            raise SystemExit("c5 shards 64 clouds: --gpus must divide 64")
        self.B = 64 // world
        self.x = []
        for n in self.SIZES:
            rng = np.random.default_rng(5000 + int(np.log2(n)) + 100 * rank)
            self.x.append(self.pin(unit_ball(rng, self.B, n)))
        self.units_per_step = self.B * sum(self.SIZES)
        tot = sum(n // 4 for n in self.SIZES)
        self.d_result = torch.zeros((self.B, 2 * tot), dtype=torch.int32, device=dev)  # [FPS idx | pts_cnt] of every size
        self.h_result = torch.zeros((self.B, 2 * tot), dtype=torch.int32).pin_memory()

    def step_body(self):
        import torch
        import dispu_b200.tf_grouping as G
        import dispu_b200.tf_sampling as S
        o = 0
        with torch.no_grad():
            for n, x in zip(self.SIZES, self.x):
                m, r = n // 4, (32.0 / n) ** (1.0 / 3.0)
                fidx = S.farthest_point_sample(m, x)
                centres = S.gather_point(x, fidx)
                _, cnt = G.query_ball_point(r, 32, x, centres)
                self.d_result[:, o:o + m].copy_(fidx)
                self.d_result[:, o + m:o + 2 * m].copy_(cnt)
                o += 2 * m

    def alg(self, key):
        nm = key.split(" ")[0]
        ints = [int(x) for x in key.split(" ")[1].split(",")] if " " in key else []
        if nm == "dispu_fps":
            b, n, m = ints[:3]
            return ("bytes", b * (m - 1) * n * 16)
        if nm in ("dispu_query_ball_point", "dispu_query_ball_point_ws"):
            b, n, m = ints[:3]
            return ("bytes", b * (12 * n + 12 * m + 4 * m * 32 + 4 * m))
        if nm == "dispu_gather_point":
            b, n, m = ints[:3]
            return ("bytes", b * 28 * m)
        return None

    def config(self):
        return {"workload": "c5: FPS (m=N/4) + ball query (ns=32, r=(32/N)^(1/3)) sweep, N in {1024,4096,16384,65536}, B=64 clouds over all ranks",
                "batch_per_gpu": self.B, "sizes": list(self.SIZES)}

    def cpu_arm(self, steps, warmup):
        return self.cpu_arm_static(steps, warmup)

    @staticmethod
    def cpu_arm_static(steps, warmup):
        from oracle import oracle as O
        from oracle import ref as R
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import unit_ball
        use_ref = R.available("cpu")
        sb, sizes = 1, (1024, 4096)
        xs = [unit_ball(np.random.default_rng(5000 + int(np.log2(n))), sb, n) for n in sizes]

        def step():
            for n, x in zip(sizes, xs):
                m, r = n // 4, (32.0 / n) ** (1.0 / 3.0)
                c = O.gather_point(x, O.farthest_point_sample(m, x))
                R.cpu_query_ball_point(r, 32, x, c) if use_ref else O.query_ball_point(r, 32, x, c, mode=O.CPU)

        dt = _time_cpu(step, steps, warmup)
        sample = (f"{sb} cloud per size, sizes {list(sizes)} only (bounded sample) x {steps} steps; literal single-thread FPS restatement "
                  f"(the reference has no CPU FPS kernel) + reference query_ball_point_cpu host loop (oracle/_ref), 1 thread")
        return sb * sum(sizes) / dt, dt * 1e3, 1, ("reference" if use_ref else "port"), sample


def make_workload(args, rank, world, dev):
    if args.config == "c2":
        return GeneratorWorkload(args, rank, world, dev, stages=1)
    if args.config == "c4":
        return GeneratorWorkload(args, rank, world, dev, stages=2)
    if args.config == "c3":
        return SetAbstractionWorkload(args, rank, world, dev)
    return SweepWorkload(args, rank, world, dev)


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the selected workload on the host cores (rank 0 only)."""
    if rank != 0:
        return
    if args.config in ("c2", "c4"):
        stages = 1 if args.config == "c2" else 2
        val, ms, cores, kind, sample = cpu_arm_generator(args.steps, args.warmup, stages=stages,
                                                         sample_b=args.ref_sample if stages == 1 else max(1, args.ref_sample // 4))
        metric = METRIC
        workload = ("c2: B=32/GPU, N=256->1024 (4x) generator forward + Chamfer/Hausdorff(coarse,fine)" if stages == 1 else
                    "c4: two-stage 16x generator 256->1024->4096 + Chamfer/Hausdorff vs 4096-point gt")
    else:
        cls = SetAbstractionWorkload if args.config == "c3" else SweepWorkload
        val, ms, cores, kind, sample = cls.cpu_arm_static(args.steps, args.warmup)
        metric, workload = cls.metric, (cls.__doc__ or args.config).split(":")[0]
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dispu_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — dispu_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.lib()  # fail loudly if the CUDA library is missing
    torch.backends.cuda.matmul.allow_tf32 = False  # stock dense layers stay FP32 like the reference's
    torch.backends.cudnn.allow_tf32 = False

    wl = make_workload(args, rank, world, dev)
    flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)  # > 126 MB L2
    has_coll = world > 1 and wl.collective is not None
    in_graph_collective = has_coll and not args.eager_collective

    peer = None
    if has_coll and args.collective == "peer":
        # this library's own NVLink peer-memory all-reduce kernel (csrc/peer_reduce.cu); every rank takes the same
        # decision inside PeerAllReduce, so a fallback to NCCL is collective too
        from dispu_b200.dist import PeerAllReduce
        try:
            peer = PeerAllReduce(dev)
        except RuntimeError as e:
            if rank == 0:
                print(f"[bench] {e}; using ncclAllReduce", file=sys.stderr)

    def allreduce():
        # ONE collective on the path: sum of the per-rank Chamfer partials, enqueued on the step's
        # stream — inside the captured CUDA graph, so a replay contains the all-reduce kernel and no host work;
        # (Hausdorff is only logged by the trainer — max over ranks is taken on the host from the same 4 floats)
        if peer is not None:
            peer(wl.collective)
        else:
            dist.all_reduce(wl.collective, op=dist.ReduceOp.SUM)

    def step_body():
        wl.step_body()
        if in_graph_collective:
            allreduce()

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(3):
            step_body()
        stream.synchronize()
        l0 = _lib.launch_count()
        step_body()
        launches_per_step = _lib.launch_count() - l0
        graph = None
        if not args.no_graph:
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    step_body()
            except Exception as e:  # pragma: no cover
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                graph = None
        stream.synchronize()

        def step():
            if graph is not None:
                graph.replay()
            else:
                step_body()
            if has_coll and not in_graph_collective:
                allreduce()

        def barrier():
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for _ in range(max(args.warmup, 3)):
            step()
        # ---------------- device-timed value
        with ClockSampler(local_rank) as clocks:
            barrier()
            evs = []
            for _ in range(args.steps):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step()
                e1.record(stream)
                evs.append((e0, e1))
            barrier()
            dev_ms = sum(a.elapsed_time(b) for a, b in evs)
            # ---------------- end-to-end (host buffers in, host result out, every step)
            barrier()
            evs = []
            for _ in range(args.steps):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for h, d in zip(wl.host, wl.device):
                    d.copy_(h, non_blocking=True)
                step()
                wl.h_result.copy_(wl.d_result, non_blocking=True)
                e1.record(stream)
                e1.synchronize()  # the caller reads the result every step
                evs.append((e0, e1))
            barrier()
            e2e_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])

        # ---------------- roofline of this library's dominant kernel in the step (rank 0)
        roof = None
        cpu_base = None
        if rank == 0:
            roof = dominant_kernel_roofline(wl, stream, flush)

    if rank == 0:
        if args.cpu_baseline:
            v, ms, cores, kind, sample = wl.cpu_arm(steps=1, warmup=1)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        units = wl.units_per_step * world * args.steps
        cfg = wl.config()
        cfg.update({"cuda_graph": graph is not None,
                    "l2": "flushed between timed steps (192 MiB memset outside the event pairs)",
                    "sharding": ("single GPU" if world == 1 else
                                 (f"batch-sharded x{world}, one all-reduce of 2 Chamfer partial sums, "
                                  f"{'dispu_peer_allreduce_sum (own kernel, NVLink peer stores)' if peer is not None else 'ncclAllReduce'} "
                                  f"({'captured in the CUDA graph' if (in_graph_collective and graph is not None) else 'eager, after the replay'})")
                                 if wl.collective is not None else f"clouds sharded x{world}, no data-path collective")})
        line = {
            "metric": wl.metric, "value": units / (dev_ms * 1e-3), "unit": wl.unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": wl.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": wl.unit, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in wl.host)),
                    "d2h_bytes_per_step": int(wl.h_result.numel() * wl.h_result.element_size())},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks.summary(),
            "roofline": roof,
            "cpu_baseline": cpu_base,
        }
        line.update(wl.extra())
        print(json.dumps(line), file=RESULT_OUT, flush=True)


def kernel_traffic(family):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture, or None.  The file is
    stamped with the SHA-256 of the kernel's source: a capture taken from different code is refused (None), so the number
    cannot silently go stale."""
    import hashlib
    tp = os.path.join(ROOT, "profiles", "dense_traffic.json")
    if not os.path.exists(tp):
        return None, "no capture committed"
    d = json.load(open(tp))
    ent = d.get(family)
    if not ent:
        return None, "no capture for this kernel"
    src = os.path.join(ROOT, "dispu_b200", "csrc", ent["source"])
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
    if sha != ent.get("source_sha256"):
        return None, f"profiles/dense_traffic.json was captured from a different {ent['source']} (stale): re-capture with tools/gpu_evidence.sh and stamp it with tools/ncu_traffic.py"
    return ent["dram_bytes_per_launch"], f"profiles/dense_traffic.json ({ent.get('how', 'ncu dram__bytes_read+write')})"


def dominant_kernel_roofline(wl, stream, flush, reps=10):
    """Per-kernel device time INSIDE the step: the step runs eagerly `reps` times with a CUDA-event pair around
    every C-ABI call of this library (dispu_b200._t._profile; a spin kernel at the head of each step keeps the
    launch queue full so the pairs bracket execution, not host launch latency; L2 flushed before each step).
    Reports the roofline of the kernel family with the largest share of the step:
    achieved = ALGORITHMIC flops or bytes (Workload.alg) / measured duration."""
    import torch
    from dispu_b200 import _t
    hbm, bf16_peak, sm_mhz, which = measured_peaks()
    recs = []
    for rep in range(reps + 2):
        flush.zero_()
        torch.cuda._sleep(24_000_000)  # ~12 ms of device spin: the whole step is enqueued behind it
        _t._profile = [] if rep >= 2 else None
        try:
            wl.profile_body()
            prof = _t._profile
        finally:
            _t._profile = None
        stream.synchronize()
        if prof:
            recs.append([(k, a.elapsed_time(b)) for k, a, b in prof])
    per_key = {}
    for step in recs:
        for k, ms in step:
            e = per_key.setdefault(k, [0.0, 0])
            e[0] += ms
            e[1] += 1
    n = len(recs)
    fp32 = fp32_peak_tflops(sm_mhz)
    fam = {}
    detail = {}
    for k, (ms_tot, cnt) in per_key.items():
        ms_step, per_step = ms_tot / n, cnt / n
        ms_l = ms_tot / cnt
        ent = {"ms_per_launch": ms_l, "launches_per_step": per_step, "ms_per_step": ms_step}
        nm = k.split(" ")[0]
        a = wl.alg(k)
        tensor_fam = nm in ("dispu_dense_tf32x3_ex", "dispu_attention_tf32x3", "dispu_sa_mlp")
        fname = {"dispu_dense_tf32x3_ex": "dense_tf32x3_kernel", "dispu_attention_tf32x3": "attention_tf32x3_kernel",
                 "dispu_sa_mlp": "sa_mlp_tc_kernel"}.get(nm, nm)
        fm = fam.setdefault(fname, {"ms": 0.0, "launches": 0.0, "flops": 0.0, "bytes": 0.0, "tensor": tensor_fam})
        if a is not None and a[0] == "flops":
            ent["TFLOPs_algorithmic"] = a[1] / ms_l / 1e9
            fm["flops"] += a[1] * per_step
        elif a is not None:
            ent["alg_bytes"] = a[1]
            ent["GBps"] = a[1] / (ms_l * 1e-3) / 1e9
            ent["frac_of_hbm"] = ent["GBps"] / hbm
        if a is not None and a[0] == "bytes":
            fm["bytes"] += a[1] * per_step
        ints = [int(x) for x in k.split(" ")[1].split(",")] if (" " in k and k.split(" ")[1][:1].isdigit()) else []
        if nm == "dispu_fps" and len(ints) >= 3:
            b, nn, m = ints[:3]
            # the cloud is register / shared-memory resident: the sweep-model bytes never reach HBM.  What bounds the kernel is
            # the dependent step (m-1 of them) and, inside a step, the FP32 distance updates:
            ent["point_updates_per_ns"] = b * (m - 1) * nn / ms_l / 1e6
            ent["frac_of_fp32_peak"] = b * (m - 1) * nn * 10.0 / (ms_l * 1e-3) / 1e12 / fp32  # ~10 flop per update
            ent["us_per_dependent_step"] = ms_l * 1e3 / max(m - 1, 1)
            ent["compulsory_hbm_bytes"] = b * (12 * nn + 4 * m)
        if nm in ("dispu_query_ball_point", "dispu_query_ball_point_ws", "dispu_knn", "dispu_chamfer", "dispu_chamfer_ex") and len(ints) >= 3:
            b, nn, m = ints[:3]
            pairs = (2.0 if nm.startswith("dispu_chamfer") else 1.0) * b * nn * m
            # brute-force-equivalent distance tests (the grid ball query visits far fewer candidates: an ALGORITHMIC rate)
            ent["equiv_distance_tests_per_ns"] = pairs / ms_l / 1e6
            ent["equiv_frac_of_fp32_peak"] = pairs * 8.0 / (ms_l * 1e-3) / 1e12 / fp32
        fm["ms"] += ms_step
        fm["launches"] += per_step
        detail[k] = ent
    lib_ms = sum(f["ms"] for f in fam.values())
    name = max(fam, key=lambda q: fam[q]["ms"])
    top = fam[name]
    shares = {q: round(f["ms"] / lib_ms, 4) for q, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    common = {"ms_per_launch": top["ms"] / top["launches"], "launches_per_step": top["launches"],
              "ms_per_step": top["ms"], "library_ms_per_step": lib_ms, "share_of_library_time": shares,
              "how": f"CUDA-event pair around every launch of the step, eager behind a 12 ms device spin, {n} steps, L2 flushed per step",
              "fp32_peak_tflops": fp32, "all": detail}
    if top["tensor"]:
        achieved = top["flops"] / top["ms"] / 1e9
        traffic, tsrc = kernel_traffic(name)
        return {"bound": "tensor", "kernel": name + " (tcgen05.mma kind::tf32, 3xTF32, A operand + FP32 accumulators in TMEM)",
                "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak, "traffic": traffic,
                "traffic_source": tsrc, "peak_source": which + " dense bf16 burst",
                "algorithmic_flops_per_launch": top["flops"] / top["launches"],
                "note": "algorithmic FLOPs = 2*M*K*N of the fp32 contraction; the kernel issues 3x that on the tensor pipe at the TF32 "
                        "rate (half the bf16 rate), so the ceiling for this figure is peak/6; X is read from HBM as fp32 and split on the fly",
                **common}
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    note = None
    if name == "dispu_fps":
        note = ("sweep-model bytes B(m-1)n*16 (SURVEY.md §8d-i: what the reference streams per step); the cloud is on chip here, so the real "
                "bounds are the dependent-step latency and FP32 updates: see all[...].us_per_dependent_step / frac_of_fp32_peak / compulsory_hbm_bytes")
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": None, "peak_source": which, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"], "note": note, **common}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dispu", choices=["dispu", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json config: c2 (default, the headline), c3 set abstraction, c4 two-stage 16x, c5 FPS + ball-query sweep")
    ap.add_argument("--ref-sample", type=int, default=8, help="patches per CPU step (bounded sample of the batch)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--collective", choices=("peer", "nccl"), default="peer",
                    help="the path's one all-reduce: this library's NVLink peer-memory kernel (default; single node) or ncclAllReduce")
    ap.add_argument("--eager-collective", action="store_true", help="issue the all-reduce from the host after the graph replay (round-1 form)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line.  Libraries that print to fd 1 (NCCL's version banner under NCCL_DEBUG=VERSION
    # does) are sent to stderr: keep a private copy of the real stdout for the result line, point fd 1 at stderr.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION|INFO) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_product(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the Dis-PU hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

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


def cpu_arm(steps, warmup, sample_b):
    """Times the CPU path on `sample_b` clouds per step; returns (points/s, ms/step, cores, kind, sample)."""
    import torch
    from oracle import ref as R
    from dispu_b200.generator import Generator
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp, gt = synth_batch(0)
    inp, gt = inp[:sample_b], gt[:sample_b]
    gen = Generator(device="cpu", ops=CpuRefOps())
    x = torch.from_numpy(inp)

    def step():
        with torch.no_grad():
            coarse, fine = gen(x)
        c1 = cpu_chamfer(coarse.numpy(), gt)
        c2 = cpu_chamfer(fine.numpy(), gt)
        return c1, c2

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    kind = "reference" if (R.available("knn") and R.available("cpu")) else "port"
    sample = (f"{sample_b} of {B_PER_GPU} patches per step x {steps} steps; reference nanoflann kNN (OpenMP, {cores} threads) "
              f"+ reference nnsearch NnDistance (1 thread, as the TF CPU op) from oracle/_ref; stock dense layers = torch CPU "
              f"({cores} threads) restatement of generator.py (TF is not installable here)")
    return sample_b * N_IN * steps / dt, dt / steps * 1e3, cores, kind, sample


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    val, ms, cores, kind, sample = cpu_arm(args.steps, args.warmup, sample_b=args.ref_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: B=32/GPU, N=256->1024 (4x) generator forward + Chamfer/Hausdorff(coarse,fine)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dispu_b200 import _lib, loss_utils
    from dispu_b200.generator import Generator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — dispu_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.lib()  # fail loudly if the CUDA library is missing
    torch.backends.cuda.matmul.allow_tf32 = False  # stock dense layers stay FP32 like the reference's
    torch.backends.cudnn.allow_tf32 = False

    B = B_PER_GPU
    inp_np, gt_np = synth_batch(rank)
    h_inp = torch.from_numpy(inp_np).pin_memory()
    h_gt = torch.from_numpy(gt_np).pin_memory()
    d_inp = torch.empty_like(h_inp, device=dev)
    d_gt = torch.empty_like(h_gt, device=dev)
    d_inp.copy_(h_inp)
    d_gt.copy_(h_gt)
    gen = Generator(device=dev)
    out_c = torch.zeros((4,), dtype=torch.float32, device=dev)
    out_f = torch.zeros((4,), dtype=torch.float32, device=dev)
    # [sum cd_coarse, sum cd_fine | hd_coarse, hd_fine]: the fused Chamfer kernels deposit their partials here themselves
    # (dispu_chamfer_ex), the first two floats are what the ONE collective of the path sums over the ranks
    partial = torch.zeros((4,), dtype=torch.float32, device=dev)
    h_loss = torch.zeros((4,), dtype=torch.float32).pin_memory()
    flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)  # > 126 MB L2
    in_graph_collective = world > 1 and not args.eager_collective

    def allreduce():
        if world > 1:
            # ONE collective on the path: sum of the per-rank Chamfer partials, enqueued on the step's
            # stream — inside the captured CUDA graph, so a replay contains the ncclAllReduce kernel and no host work;
            # (Hausdorff is only logged by the trainer — max over ranks is taken on the host from the same 4 floats)
            dist.all_reduce(partial[:2], op=dist.ReduceOp.SUM)

    def step_body():
        with torch.no_grad():
            # the coarse cloud's loss runs on a side stream as soon as the coarse cloud exists (overlaps PointShuffle2)
            coarse, fine = gen(d_inp, on_coarse=lambda c: loss_utils.chamfer_hausdorff(
                c, d_gt, out=out_c, sum_out=partial[0:1], max_out=partial[2:3]))
            loss_utils.chamfer_hausdorff(fine, d_gt, out=out_f, sum_out=partial[1:2], max_out=partial[3:4])
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
            if not in_graph_collective:
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
            # ---------------- end-to-end (host buffers in, host scalars out, every step)
            barrier()
            evs = []
            for _ in range(args.steps):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                d_inp.copy_(h_inp, non_blocking=True)
                d_gt.copy_(h_gt, non_blocking=True)
                step()
                h_loss.copy_(partial, non_blocking=True)
                e1.record(stream)
                e1.synchronize()  # the caller reads the loss every step
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
            roof = dominant_kernel_roofline(gen, d_inp, d_gt, stream, flush)

    if rank == 0:
        if args.cpu_baseline:
            v, ms, cores, kind, sample = cpu_arm(steps=1, warmup=1, sample_b=args.ref_sample)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        pts = B * N_IN * world * args.steps
        loss = h_loss.tolist()
        line = {
            "metric": METRIC, "value": pts / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: B=32/GPU, N=256->1024 (4x) generator forward + Chamfer/Hausdorff(coarse,fine)",
                       "batch_per_gpu": B, "n_in": N_IN, "up_ratio": UP, "cuda_graph": graph is not None,
                       "l2": "flushed between timed steps (192 MiB memset outside the event pairs)",
                       "weights": "Xavier-uniform seed 7, zero bias (random init; no checkpoint offline)",
                       "sharding": (f"batch-sharded x{world}, one all-reduce of 2 Chamfer partial sums "
                                    f"({'captured in the CUDA graph' if (in_graph_collective and graph is not None) else 'eager, after the replay'})")
                       if world > 1 else "single GPU",
                       "streams": "kNN of the coarse cloud and the coarse Chamfer on side streams (graph fork/join)" if gen.use_streams else "single stream"},
            "e2e": {"value": pts / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(h_inp.numel() * 4 + h_gt.numel() * 4), "d2h_bytes_per_step": 16},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks.summary(),
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "loss": {"cd_coarse_sum": loss[0], "cd_fine_sum": loss[1], "hd_coarse": loss[2], "hd_fine": loss[3]},
        }
        print(json.dumps(line), file=RESULT_OUT, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def dominant_kernel_roofline(gen, d_inp, d_gt, stream, flush, reps=10):
    """Per-kernel device time INSIDE the step: the step runs eagerly `reps` times with a CUDA-event pair around
    every C-ABI call of this library (dispu_b200._t._profile; a spin kernel at the head of each step keeps the
    launch queue full so the pairs bracket execution, not host launch latency; L2 flushed before each step).
    Reports the roofline of the kernel family with the largest share of the step:
    achieved = ALGORITHMIC flops (2*M*K*N of each dense launch) or bytes / measured duration."""
    import torch
    from dispu_b200 import _t, loss_utils
    hbm, bf16_peak, which = measured_peaks()
    out_c = torch.zeros((4,), dtype=torch.float32, device=d_inp.device)
    recs = []
    use_streams, gen.use_streams = gen.use_streams, False  # one stream: the event pairs must bracket one kernel each
    with torch.no_grad():
        for rep in range(reps + 2):
            flush.zero_()
            torch.cuda._sleep(24_000_000)  # ~12 ms of device spin: the whole step is enqueued behind it
            _t._profile = [] if rep >= 2 else None
            try:
                coarse, fine = gen(d_inp)
                loss_utils.chamfer_hausdorff(coarse, d_gt, out=out_c)
                loss_utils.chamfer_hausdorff(fine, d_gt, out=out_c)
                prof = _t._profile
            finally:
                _t._profile = None
            stream.synchronize()
            if prof:
                recs.append([(k, a.elapsed_time(b)) for k, a, b in prof])
    gen.use_streams = use_streams
    per_key = {}
    for step in recs:
        for k, ms in step:
            e = per_key.setdefault(k, [0.0, 0])
            e[0] += ms
            e[1] += 1
    n = len(recs)
    b, m = d_inp.shape[0], d_inp.shape[1] * UP
    n_in, kk = d_inp.shape[1], 16

    def alg_bytes(key):
        """SURVEY.md §8d / DESIGN.md §4 algorithmic bytes of one launch of the non-GEMM kernels at the step's shapes."""
        nm = key.split(" ")[0]
        if nm == "dispu_knn":
            return b * (12 * m + 12 * m + 4 * m * kk)
        if nm == "dispu_chamfer":
            return b * (12 * m + 12 * m) + 16
        if nm == "dispu_group_max_ld":
            return b * (4 * m * kk + 4 * m * 132 + 4 * m * 132)
        if nm == "dispu_weightnet_aggregate":
            return b * m * (4 * kk * 128 + 4 * 2048 + 4 * kk + 12)
        if nm == "dispu_feature_knn":
            c = int(key.split(" ")[1].split(",")[3])
            return b * n_in * (4 * c + 4 * kk)
        if nm in ("dispu_dense_conv", "dispu_dense_conv_knn"):
            c = int(key.split(" ")[1].split(",")[2])
            return b * n_in * (4 * c + 4 * kk + 4 * (72 + c))
        return None

    fam = {}
    detail = {}
    for k, (ms_tot, cnt) in per_key.items():
        ms_step, per_step = ms_tot / n, cnt / n
        ent = {"ms_per_launch": ms_tot / cnt, "launches_per_step": per_step, "ms_per_step": ms_step}
        if k.startswith("dispu_dense_tf32x3_ex"):
            f = k.split(" ")
            M_, K_, N_ = int(f[1][1:]), int(f[2][1:]), int(f[3][1:])
            fl = 2.0 * M_ * K_ * N_
            for tok in f[4:]:
                if tok.startswith("x") and tok.endswith("blocks"):
                    fl *= int(tok[1:-6])  # column blocks: N per block
                if tok == "+weightnet":
                    fl += 2.0 * M_ * N_ * 16  # transposed product with the weight net in the epilogue
            ent["TFLOPs_algorithmic"] = fl / (ms_tot / cnt) / 1e9
            fm = fam.setdefault("dense_tf32x3_kernel", {"ms": 0.0, "launches": 0.0, "flops": 0.0})
            fm["flops"] += fl * per_step
        else:
            ab = alg_bytes(k)
            if ab is not None:
                ent["alg_bytes"] = ab
                ent["GBps"] = ab / (ms_tot / cnt * 1e-3) / 1e9
                ent["frac_of_hbm"] = ent["GBps"] / hbm
            fm = fam.setdefault(k.split(" ")[0], {"ms": 0.0, "launches": 0.0, "bytes": 0.0})
            fm["bytes"] = fm.get("bytes", 0.0) + (ab or 0) * per_step
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
              "all": detail}
    if name == "dense_tf32x3_kernel":
        achieved = top["flops"] / top["ms"] / 1e9
        traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of the same step
        tp = os.path.join(ROOT, "profiles", "r01_dense_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        return {"bound": "tensor", "kernel": "dense_tf32x3_kernel (tcgen05.mma kind::tf32 TS form, 3xTF32, A operand + FP32 accumulators in TMEM)",
                "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak, "traffic": traffic,
                "traffic_source": "profiles/r01_dense_traffic.json (ncu dram__bytes_read+write, average over the step's dense launches)",
                "peak_source": which + " dense bf16 burst",
                "algorithmic_flops_per_launch": top["flops"] / top["launches"],
                "note": "algorithmic FLOPs = 2*M*K*N of the fp32 contraction; the kernel issues 3x that on the tensor pipe at the TF32 "
                        "rate (half the bf16 rate), so the ceiling for this figure is peak/6; X is read from HBM as fp32 and split on the fly",
                **common}
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": None, "peak_source": which, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"], **common}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dispu", choices=["dispu", "reference"])
    ap.add_argument("--ref-sample", type=int, default=8, help="patches per CPU step (bounded sample of the batch)")
    ap.add_argument("--no-graph", action="store_true")
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

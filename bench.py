#!/usr/bin/env python
"""Benchmark of the hierarchical-loss hot path (BASELINE.json metric: hier-loss fwd+bwd Gpix/s and
fraction of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (the reference's own modules when
                                                             # /root/reference or baseline/_ref exists, else the oracle
                                                             # port; rank 0 only)

A "step" is one forward+backward of the loss module over one synthetic batch (per rank; weak scaling).
Default workload = BASELINE config 3 (the configuration the target is quoted on):
RMIHieraTripletLoss, 19->7->2 classes, logits fp32 [8, 28, 1024, 2048] per GPU, labels L-blob (32x32
constant tiles, 10 % ignore), embedding [8, 256, 32, 64].  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HI_19_7 = [[0, 2], [2, 5], [5, 8], [8, 10], [10, 11], [11, 13], [13, 19]]
HM_19_7 = [0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 6, 6, 6, 6, 6, 6]
F2H_19 = [0] * 11 + [1] * 8

WORKLOADS = {
    # name: kind, levels, B, H, W, dtype, algorithmic bytes/pixel formula pieces
    "cfg3": dict(kind="3level", nf=19, nm=7, nh=2, B=8, H=1024, W=2048, dtype="fp32",
                 desc="config 3: RMIHieraTripletLoss 19/7/2, 1024x2048, batch 8/GPU, fp32 logits"),
    "cfg3-bf16": dict(kind="3level", nf=19, nm=7, nh=2, B=8, H=1024, W=2048, dtype="bf16",
                      desc="config 3 with bf16 logits"),
    "cfg2": dict(kind="2level", nf=19, nc=7, B=16, H=512, W=1024, dtype="bf16",
                 desc="config 2: HieraTripletLoss 19/7, 512x1024, batch 16, bf16 logits"),
    "cfg4": dict(kind="3level", nf=150, nm=30, nh=6, B=32, H=512, W=512, dtype="fp32", triplet=False,
                 desc="config 4: 150/30/6 classes, 512x512, batch 32/GPU (triplet undefined in the reference, off)"),
    "train-cfg3": dict(kind="train", levels=3, nf=19, nm=7, nh=2, B=8, H=1024, W=2048, dtype="fp32", depth=101,
                       desc="config 3 end to end: ResNet-101 + SepASPP head (context, stock cuDNN) + RMIHieraTripletLoss, "
                            "1024x2048, batch 8/GPU, bf16 autocast network, fp32 full-resolution logits"),
    "train-cfg2": dict(kind="train", levels=2, nf=19, nc=7, B=16, H=512, W=1024, dtype="bf16", depth=50,
                       desc="config 2 end to end: ResNet-50 + SepASPP head (context) + HieraTripletLoss, 512x1024, "
                            "batch 16, bf16 network and logits"),
    "cfg5": dict(kind="decode", nf=19, nm=7, nh=2, B=64, H=2048, W=2048, dtype="bf16",
                 desc="config 5: hierarchical argmax decode, 2048x2048, batch 64, bf16 logits"),
}


def algorithmic_bytes_per_px(w):
    e = 4 if w["dtype"] == "fp32" else 2
    if w["kind"] == "3level":
        c = w["nf"] + w["nm"] + w["nh"]
        return dict(total=3 * c * e + 16, pass1=c * e + 8, pass2=2 * c * e + 8)
    if w["kind"] == "2level":
        c = w["nf"] + w["nc"]
        return dict(total=2 * c * e + 8, fused=2 * c * e + 8)
    c = w["nf"] + w["nm"] + w["nh"]
    return dict(total=c * e + 3 * 8, decode=c * e + 3 * 8)   # int64 outputs like torch.argmax


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None
        self.windows = []          # [t0, t1] wall-clock windows (timed regions); only samples inside them count

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            # a sample describes the ~20 ms before it was printed
            if self.windows and not any(a <= ts <= b + 0.03 for a, b in self.windows):
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


class StageTimer:
    """Brackets every kernel-level stage of the multi-kernel C entry points with CUDA events."""

    def __init__(self):
        import torch
        self.torch = torch
        self.pending, self.open = [], {}

    def start(self, name, bit):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        self.open[(name, bit)] = ev

    def stop(self, name, bit):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        self.pending.append(((name, bit), self.open.pop((name, bit)), ev))

    def totals(self):
        out = {}
        for key, a, b in self.pending:
            t, n = out.get(key, (0.0, 0))
            out[key] = (t + a.elapsed_time(b), n + 1)
        return out


def make_labels(torch, gen, b, h, w, n_fine, kind, dev):
    if kind == "iid":
        lab = torch.randint(0, n_fine, (b, h, w), generator=gen, device=dev)
        lab[torch.rand(b, h, w, generator=gen, device=dev) < 0.1] = 255
        return lab
    tile = 32
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    c = torch.randint(0, n_fine, (b, th, tw), generator=gen, device=dev)
    c[torch.rand(b, th, tw, generator=gen, device=dev) < 0.1] = 255
    return c.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w].contiguous()


def hierarchy_maps(w):
    if w["nf"] == 19:
        return HM_19_7, F2H_19
    return [f // 5 for f in range(w["nf"])], [f // 25 for f in range(w["nf"])]


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_shape(w):
    if w["kind"] == "decode":
        return 4, 512, 512
    if w["kind"] == "2level":
        return 2, 256, 512
    return (1, 512, 1024) if w["nf"] <= 32 else (1, 256, 256)


def _reference_root():
    """Where the UNMODIFIED reference can be imported from on this box (None on the GPU box: it does not travel)."""
    for cand in (os.environ.get("SEGHIERO_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "loss", "rmi_hiera_triplet_loss.py")):
            return cand
    return None


def run_cpu_oracle(w, steps, warmup, label_kind):
    """CPU arm on a bounded crop of the workload: the reference's own loss modules (kind "reference", CPU tensors,
    `.cuda()` shim for its hard-coded device moves) when the reference can be imported here, else the oracle port."""
    import torch
    import torch.nn.functional as F
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, h, wd = cpu_sample_shape(w)
    g = torch.Generator().manual_seed(1234)
    lab = make_labels(torch, g, b, h, wd, w["nf"], label_kind, "cpu")
    dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16
    f2m, f2h = hierarchy_maps(w)
    ref_root = _reference_root() if w["kind"] in ("2level", "3level") else None
    kind = "port"
    if ref_root is not None:
        try:
            sys.dont_write_bytecode = True
            sys.path.insert(0, ref_root)
            torch.Tensor.cuda = lambda self, *a, **k: self      # the reference hard-codes .cuda() in its triplet losses
            from models.loss.hiera_triplet_loss import HieraTripletLoss as RefH2
            from models.loss.rmi_hiera_triplet_loss import RMIHieraTripletLoss as RefH3
            kind = "reference"
        except Exception:                                        # noqa: BLE001 -- any import problem: fall back to the port
            kind = "port"
    from oracle import hiera_oracle as O
    if w["kind"] == "decode":
        x = torch.randn(b, w["nf"] + w["nm"] + w["nh"], h, wd, generator=g).to(dt)

        def step():
            O.argmax_decode(x, [w["nf"], w["nm"], w["nh"]])
    elif w["kind"] == "2level":
        x = (torch.randn(b, w["nf"] + w["nc"], h, wd, generator=g) * 2).to(dt).requires_grad_(True)
        emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g), dim=1).requires_grad_(True)
        ref_mod = RefH2(w["nf"], HM_19_7, HI_19_7) if kind == "reference" else None

        def step():
            x.grad = None
            emb.grad = None
            if ref_mod is not None:
                loss = ref_mod(torch.tensor([100000]), emb, None, x, lab)
            else:
                loss, _ = O.hiera_triplet_loss(100000, emb, x, lab, w["nf"], HM_19_7, HI_19_7)
            loss.backward()
    else:
        x = (torch.randn(b, w["nf"] + w["nm"] + w["nh"], h, wd, generator=g) * 2).to(dt).requires_grad_(True)
        emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g), dim=1).requires_grad_(True)
        trip = w.get("triplet", True)
        ref_mod = None
        if kind == "reference" and trip:       # config 4's labels are outside the reference's id lists: port
            ref_mod = RefH3(w["nf"], w["nm"], w["nh"], torch.tensor(f2m), torch.tensor(f2h))
        elif kind == "reference":
            kind = "port"

        def step():
            x.grad = None
            emb.grad = None
            if ref_mod is not None:
                loss = ref_mod(torch.tensor([100000]), emb, None, x, lab)
            else:
                loss, _ = O.rmi_hiera_triplet_loss(100000, emb, x, lab, w["nf"], w["nm"], w["nh"], f2m, f2h,
                                                   with_triplet=trip)
            loss.backward()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = (time.perf_counter() - t0) / max(steps, 1)
    px = b * h * wd
    what = ("the reference's own loss module (CPU tensors, .cuda() shim)" if kind == "reference"
            else "oracle port (torch-CPU restatement of the reference loss)")
    return dict(value=px / dt_s / 1e9, unit="Gpix/s", cores=cores, kind=kind,
                sample=f"{what}, {b}x{h}x{wd} crop of the workload, fwd+bwd, {steps} steps, all {cores} host threads",
                ms_per_step=dt_s * 1e3)


def bench_config(args, w, b, px, world, x_gb=None):
    cfg = {"workload": args.workload, "description": w["desc"], "labels": args.labels,
           "batch_per_gpu": b, "pixels_per_step_per_gpu": px,
           "parallelism": f"dp{world} by sample, no data-path collective"}
    if x_gb is not None:
        cfg["l2"] = "inputs larger than L2 (logits %.2f GB per GPU)" % x_gb
    if NUMA_NOTE["node"] is not None:
        cfg["host_binding"] = "each rank's host threads and pinned buffers bound to its GPU's NUMA node (rank 0: node %d)" % NUMA_NOTE["node"]
    return cfg


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    res = run_cpu_oracle(w, steps, 1, args.labels)
    b = args.batch or w["B"]
    e = 4 if w["dtype"] == "fp32" else 2
    c = w["nf"] + (w["nc"] if w["kind"] == "2level" else w["nm"] + w["nh"])
    line = {
        "impl": "reference", "metric": "hier_loss_fwd_bwd_throughput", "value": res["value"], "unit": "Gpix/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if w["dtype"] == "fp32" else "bf16 in / f32 accumulate", "data": "synthetic",
        "config": bench_config(args, w, b, b * w["H"] * w["W"], max(args.gpus, 1),
                               b * c * w["H"] * w["W"] * e / 1e9),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
STAGE_NAMES = {("sh_rmi3_forward", 1): "k3_prep", ("sh_rmi3_forward", 2): "k3_pass1",
               ("sh_rmi3_forward", 4): "k3_frame1", ("sh_rmi3_forward", 8): "k3_finalize",
               ("sh_rmi3_backward", 1): "k3_pass2", ("sh_rmi3_backward", 2): "k3_frame2",
               ("sh_bce2_fwdbwd", 1): "k_prep2", ("sh_bce2_fwdbwd", 2): "k_bce2_fused",
               ("sh_bce2_fwdbwd", 4): "k_reduce_partials", ("sh_upsample_bilinear", 0): "k_upsample",
               ("sh_upsample_bilinear_adjoint", 0): "k_upsample_adjoint", ("sh_aux_ce_fwdbwd", 0): "k_aux_ce"}
FAST_NAMES = {("sh_rmi3_forward", 1): "k3f_prep", ("sh_rmi3_forward", 2): "k3f_pass1",
              ("sh_rmi3_forward", 8): "k3f_finalize", ("sh_rmi3_backward", 1): "k3f_pass2",
              ("sh_bce2_fwdbwd", 2): "k_bce2_fast"}


NUMA_NOTE = {"node": None}


def _bind_to_gpu_numa_node(local):
    """One process per GPU: run this rank's host threads (and first-touch its pinned staging buffers) on the NUMA node the
    GPU hangs off.  torchrun does not bind ranks; unbound, the pinned buffers of all 8 ranks can land on one node and the
    host<->device copies of the e2e path share that node's memory controllers (round 1: e2e flat from 1 to 8 GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        node = int(open(path).read().strip())
        if node < 0:
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            NUMA_NOTE["node"] = node
    except Exception:      # noqa: BLE001 -- binding is an optimisation of the host side, never a requirement
        pass


def _dist_env():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("SEGHIERO_B200_NO_NUMA_BIND") is None:
        _bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh).get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except (OSError, ValueError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _timed_steps(torch, dist, world, dev, step, steps, timer=None):
    """K device-timed steps between barriers; returns (ms per step as the max over ranks, wall window, last output)."""
    from seghiero_b200 import ops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ctx = ops.stage_timing(timer) if timer is not None else None
    if ctx is not None:
        ctx.__enter__()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    if ctx is not None:
        ctx.__exit__(None, None, None)
    barrier()
    t1 = time.time()
    el = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    return float(el.item()), (t0, t1), out


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import seghiero_b200 as sb
    from seghiero_b200 import ops

    rank, world, local, dev = _dist_env()
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    b, h, wd = args.batch or w["B"], w["H"], w["W"]
    dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16
    px = b * h * wd
    f2m, f2h = hierarchy_maps(w)
    lab = make_labels(torch, g, b, h, wd, w["nf"], args.labels, dev)
    step_t = torch.tensor([100000], device=dev)
    mod = None

    if w["kind"] == "decode":
        c = w["nf"] + w["nm"] + w["nh"]
        x = torch.randn(b, c, h, wd, generator=g, device=dev, dtype=torch.float32).to(dt)
        emb = None

        def step():
            preds, counts = sb.hierarchical_argmax(x, [w["nf"], w["nm"], w["nh"]], lab)
            return counts
    else:
        if w["kind"] == "2level":
            c = w["nf"] + w["nc"]
            mod = sb.HieraTripletLoss(w["nf"], HM_19_7, HI_19_7)
        else:
            c = w["nf"] + w["nm"] + w["nh"]
            mod = sb.RMIHieraTripletLoss(w["nf"], w["nm"], w["nh"], torch.tensor(f2m), torch.tensor(f2h))
            if not w.get("triplet", True):
                mod.triplet_loss = None
        x = (torch.randn(b, c, h, wd, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g, device=dev), dim=1).requires_grad_(True)
        use_emb = emb if (w["kind"] == "2level" or w.get("triplet", True)) else None
        cur = {"lab": lab}

        def step():
            x.grad = None
            emb.grad = None
            loss = mod(step_t, use_emb, None, x, cur["lab"])
            loss.backward()
            if world > 1:   # scalar loss sum across ranks (the path's only data collective besides `ready`)
                dist.all_reduce(loss.detach(), op=dist.ReduceOp.SUM)
            return loss

    clk = ClockSampler(local)
    clk.__enter__()                       # nvidia-smi needs ~0.1 s before its first sample: start it before the warm-up
    for _ in range(warmup):
        step()

    # ---- timed region: K steps, device-timed, stage events on --------------------------------------
    timer = StageTimer()
    ops.LAUNCHES["n"] = 0
    ms_step, win, _ = _timed_steps(torch, dist, world, dev, step, steps, timer)
    clk.mark(*win)
    launches = ops.LAUNCHES["n"]
    value = world * px / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (from the stage events of the same timed region) ------------
    ab = algorithmic_bytes_per_px(w)
    peak, peak_src = _peak()
    names = dict(STAGE_NAMES)
    fast3 = w["kind"] == "3level" and mod.uses_fast_path(x, lab)
    fast2 = w["kind"] == "2level"
    if fast3 or fast2:
        names.update(FAST_NAMES)
    if fast3:
        names[("sh_rmi3_backward", 1)] = mod.pass2_kernel(x, lab)       # k3t_pass2 (TMA) or k3f_pass2 (cp.async)
    stage_bytes = {"k3_pass1": ab.get("pass1"), "k3_pass2": ab.get("pass2"), "k_bce2_fused": ab.get("fused"),
                   "k_bce2_fast": ab.get("fused"), "k3f_pass1": ab.get("pass1"), "k3f_pass2": ab.get("pass2"),
                   "k3t_pass2": ab.get("pass2")}

    def stage_ms(tm):
        return {names[k]: t / n for k, (t, n) in tm.totals().items() if k in names}
    stages = stage_ms(timer)
    roofline = None
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_per_px.json")) as fh:
            traffic_tab = json.load(fh)
    except OSError:
        pass
    cand = {k: v for k, v in stages.items() if stage_bytes.get(k)}
    if cand:
        top = max(cand, key=cand.get)
        achieved = stage_bytes[top] * px / (cand[top] * 1e-3) / 1e9
        tr = traffic_tab.get(f"{args.workload}:{top}")
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": (tr["bytes_per_px"] * px if tr else None),
                    "traffic_source": (tr.get("source") if tr else None),
                    "algorithmic_bytes_per_launch": stage_bytes[top] * px, "ms_per_launch": cand[top],
                    "peak_source": peak_src,
                    "whole_step_frac": ab["total"] * px / (ms_step * 1e-3) / 1e9 / peak,
                    "per_kernel_frac": {k: stage_bytes[k] * px / (v * 1e-3) / 1e9 / peak for k, v in cand.items()}}
    elif w["kind"] == "decode":
        achieved = ab["total"] * px / (ms_step * 1e-3) / 1e9
        tr = traffic_tab.get(f"{args.workload}:k_decode")
        roofline = {"bound": "hbm", "kernel": "k_decode_vec", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": (tr["bytes_per_px"] * px if tr else None),
                    "traffic_source": (tr.get("source") if tr else None),
                    "algorithmic_bytes_per_launch": ab["total"] * px,
                    "ms_per_launch": ms_step, "peak_source": peak_src}

    # ---- worst-case label distribution (per-pixel labels), same shapes, same timed-region rules ---------------
    iid = None
    if mod is not None and args.labels == "blob" and not args.no_iid:
        cur["lab"] = make_labels(torch, g, b, h, wd, w["nf"], "iid", dev)
        for _ in range(3):
            step()
        t2 = StageTimer()
        ms_iid, win, _ = _timed_steps(torch, dist, world, dev, step, steps, t2)
        clk.mark(*win)
        iid = {"labels": "iid", "ms_per_step": ms_iid, "value": world * px / (ms_iid * 1e-3) / 1e9, "unit": "Gpix/s",
               "whole_step_frac": ab["total"] * px / (ms_iid * 1e-3) / 1e9 / peak, "ratio_to_blob": ms_iid / ms_step,
               "kernel_ms": stage_ms(t2)}
        cur["lab"] = lab

    # ---- the step a training loop runs (rows N1, N2, N4): logits at the head's resolution + uint8 labels resident in HBM,
    #      upsample and its adjoint inside the op, aux-head CE from H/16 logits; device-timed like the main region ------
    h4 = None
    if mod is not None and w["kind"] == "3level" and h % 16 == 0 and wd % 16 == 0 and not args.no_iid:
        xl = (torch.randn(b, c, h // 4, wd // 4, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        xa = (torch.randn(b, w["nf"], h // 16, wd // 16, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        lab8 = lab.to(torch.uint8)

        def step_h4():
            xl.grad = None
            xa.grad = None
            emb.grad = None
            loss = mod(step_t, use_emb, None, xl, lab8) + 0.4 * sb.aux_cross_entropy(xa, lab8)
            loss.backward()
            return loss
        for _ in range(3):
            step_h4()
        t3 = StageTimer()
        ms_h4, win, _ = _timed_steps(torch, dist, world, dev, step_h4, steps, t3)
        clk.mark(*win)
        st3 = stage_ms(t3)
        es = 2 if dt != torch.float32 else 4
        up_bytes = c * es * (1 + 1 / 16) * px           # one full-resolution tensor written (read) + its H/4 source (result)
        h4 = {"what": "fwd+bwd of the 3-level loss fed with the head's H/4 logits + uint8 labels (upsample and adjoint "
                      "inside the op, train.py:282-284) plus the aux-head CE from H/16 logits (train.py:309-313); inputs "
                      "resident in HBM, device-timed",
              "ms_per_step": ms_h4, "value": world * px / (ms_h4 * 1e-3) / 1e9, "unit": "Gpix/s", "kernel_ms": st3,
              "hbm_frac": {k: up_bytes / (st3[k] * 1e-3) / 1e9 / peak for k in ("k_upsample", "k_upsample_adjoint")
                           if st3.get(k)}}
        del xl, xa, lab8

    # ---- decode from the head's H/4 logits (N3), same number of output pixels ---------------------------------
    decode_up = None
    if w["kind"] == "decode" and h % 4 == 0 and wd % 4 == 0:
        xl = torch.randn(b, c, h // 4, wd // 4, generator=g, device=dev, dtype=torch.float32).to(dt)
        lab8 = lab.to(torch.uint8)

        def step_up():
            preds, counts = sb.hierarchical_argmax(xl, [w["nf"], w["nm"], w["nh"]], lab8, out_dtype=torch.uint8)
            return counts
        for _ in range(3):
            step_up()
        ms_up, win, _ = _timed_steps(torch, dist, world, dev, step_up, steps)
        clk.mark(*win)
        alg = (c * (2 if dt != torch.float32 else 4) / 16 + 3 + 1) * px
        decode_up = {"what": "argmax of F.interpolate(H/4 logits) + accuracy counts, fused (no full-resolution logits), "
                             "uint8 labels and predictions", "ms_per_step": ms_up,
                     "value": world * px / (ms_up * 1e-3) / 1e9, "unit": "Gpix/s",
                     "algorithmic_bytes": alg, "hbm_frac": alg / (ms_up * 1e-3) / 1e9 / peak,
                     "bound": "issue (interpolation arithmetic), not HBM"}
        del xl, lab8

    # ---- end-to-end through the public API with HOST buffers -----------------------------------------
    e2e = None
    e2e_full = None
    if not args.no_e2e:
        if w["kind"] == "decode":
            # headline: what infer.py / validation hands over once it stops upsampling first (H/4 logits, row N3); the
            # reference's own tensors (full-resolution logits over PCIe) beside it
            if h % 4 == 0 and wd % 4 == 0:
                e2e = _e2e_decode(torch, dist, sb, w, world, dev, x, lab, clk, steps, low_res=True)
                e2e_full = _e2e_decode(torch, dist, sb, w, world, dev, x, lab, clk, min(steps, 3))
            else:
                e2e = _e2e_decode(torch, dist, sb, w, world, dev, x, lab, clk, steps)
        else:
            # the reference-facing call a training loop makes with the head's output: logits at H/4 (upsampled inside
            # the op, train.py:282-284), uint8 labels; gradients come back at H/4
            e2e = _e2e_loss(torch, dist, mod, w, world, dev, x, lab, emb, use_emb, step_t, clk, steps, low_res=True)
            e2e_full = _e2e_loss(torch, dist, mod, w, world, dev, x, lab, emb, use_emb, step_t, clk, min(steps, 3),
                                 low_res=False)

    x_gb = x.numel() * x.element_size() / 1e9
    # ---- context: the oracle's torch restatement run as eager ATen ON THIS GPU (what a user of the reference gets on
    #      this box; one sample per step because the fp64 RMI unfolds need ~21 GB per 1024x2048 image) -------------
    eager = None
    if not args.no_eager and rank == 0 and world == 1 and w["kind"] in ("3level", "2level"):
        from oracle import hiera_oracle as O
        del x
        torch.cuda.empty_cache()
        eb = 1 if w["kind"] == "3level" else b
        xe = (torch.randn(eb, c, h, wd, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        le, ee = lab[:eb].contiguous(), emb[:eb].detach().clone().requires_grad_(True)

        def eager_step():
            xe.grad = None
            ee.grad = None
            if w["kind"] == "3level":
                l_, _ = O.rmi_hiera_triplet_loss(100000, ee if use_emb is not None else None, xe, le, w["nf"], w["nm"],
                                                 w["nh"], f2m, f2h, with_triplet=w.get("triplet", True))
            else:
                l_, _ = O.hiera_triplet_loss(100000, ee, xe, le, w["nf"], HM_19_7, HI_19_7)
            l_.backward()
        try:
            eager_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                eager_step()
            torch.cuda.synchronize()
            dt_s = (time.perf_counter() - t0) / 3
            eager = {"value": eb * h * wd / dt_s / 1e9, "unit": "Gpix/s", "ms_per_step": dt_s * 1e3, "batch": eb,
                     "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
                     "what": "oracle restatement of the reference loss as eager ATen on this GPU (fwd+bwd, wall clock)"}
        except torch.cuda.OutOfMemoryError:
            eager = {"unavailable": "out of memory"}
        del xe, le, ee
        torch.cuda.empty_cache()
    clk.__exit__()
    clocks = clk.summary()

    # ---- the loss inside a training step (north-star metric, second half): few steps, every N --------------------
    train = None
    if not args.no_train and args.workload in ("cfg3", "cfg2"):
        try:
            del x
        except NameError:
            pass
        torch.cuda.empty_cache()
        train = train_record(args, WORKLOADS["train-" + args.workload], min(steps, 6), brief=True)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        res = run_cpu_oracle(w, 2, 1, args.labels)
        cpu_base = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "hier_loss_fwd_bwd_throughput", "value": value, "unit": "Gpix/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if w["dtype"] == "fp32" else "bf16 in / f32 accumulate",
            "data": "synthetic", "config": bench_config(args, w, b, px, world, x_gb),
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "kernel_ms": stages,
        }
        if e2e_full is not None:
            line["e2e_full_resolution_inputs"] = e2e_full
        if iid is not None:
            line["labels_iid"] = iid
        if h4 is not None:
            line["h4_step"] = h4
        if decode_up is not None:
            line["decode_from_h4_logits"] = decode_up
        if eager is not None:
            line["eager_gpu_baseline"] = eager
        if train is not None:
            line["train"] = train
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _e2e_decode(torch, dist, sb, w, world, dev, x, lab, clk, steps, low_res=False):
    """low_res=True: the logits cross PCIe at the head's resolution (H/4) and the fused decode interpolates on the fly
    (row N3): 1/16 of the logit bytes, no full-resolution tensor anywhere."""
    b, h, wd = lab.shape
    px = b * h * wd
    if low_res:
        x = torch.nn.functional.avg_pool2d(x.float(), 4).to(x.dtype)
    hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
    hlab = torch.empty(lab.shape, dtype=torch.uint8, pin_memory=True).copy_(lab.to(torch.uint8))
    hout = [torch.empty((b, h, wd), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    hcnt = torch.empty(2, dtype=torch.int64, pin_memory=True)
    # three streams, double-buffered device inputs: the copies of neighbouring steps overlap the decode
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmp = torch.cuda.current_stream(dev)
    dbuf = [(torch.empty_like(x), torch.empty(lab.shape, dtype=torch.uint8, device=dev)) for _ in range(2)]
    free_ev = [None, None]
    state = {"i": 0}

    def e2e_step():
        i = state["i"]
        state["i"] += 1
        xd, ld = dbuf[i % 2]
        with torch.cuda.stream(s_in):
            if free_ev[i % 2] is not None:
                s_in.wait_event(free_ev[i % 2])
            xd.copy_(hx, non_blocking=True)
            ld.copy_(hlab, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        s_cmp.wait_event(ev_in)
        preds, counts = sb.hierarchical_argmax(xd, [w["nf"], w["nm"], w["nh"]], ld, out_dtype=torch.uint8,
                                               **({"size": (h, wd)} if low_res else {}))
        ev_cmp = torch.cuda.Event()
        ev_cmp.record(s_cmp)
        free_ev[i % 2] = ev_cmp
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp)
            for o, p_ in zip(hout, preds):
                p_.record_stream(s_out)
                o.copy_(p_, non_blocking=True)
            counts.record_stream(s_out)
            hcnt.copy_(counts, non_blocking=True)
    for _ in range(3):
        e2e_step()
    s_out.synchronize()
    torch.cuda.synchronize()
    k2 = max(2, min(steps, 10))
    if world > 1:
        dist.barrier()
    t0 = time.time()
    for _ in range(k2):
        e2e_step()
    s_out.synchronize()             # the last step's predictions and counts are on the host
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    clk.mark(t0, t1)
    el = torch.tensor([(t1 - t0) * 1e3 / k2], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    return {"value": world * px / (float(el.item()) * 1e-3) / 1e9, "unit": "Gpix/s",
            "h2d_bytes_per_step": int(hx.numel() * hx.element_size() + hlab.numel()),
            "d2h_bytes_per_step": int(3 * px + 16), "ms_per_step": float(el.item()), "steps": k2,
            "inputs": "logits at the head's resolution (H/4), decode fused with the interpolation" if low_res else
                      "full-resolution logits (the reference's own tensors)",
            "note": "pinned host logits + uint8 labels -> H2D -> decode -> D2H of uint8 predictions and counts, every step; "
                    "three streams, copies of neighbouring steps overlap the decode; wall clock incl. the final drain"}


def _e2e_loss(torch, dist, mod, w, world, dev, x, lab, emb, use_emb, step_t, clk, steps, low_res):
    """Same metric through the module call with HOST buffers; every step copies its inputs host->device from pinned
    memory and its results (loss, logit gradient, embedding gradient) device->host.  Three streams so that the copies
    of neighbouring steps overlap the kernels; device input buffers are double buffered.
    low_res=True: what a training loop hands over -- the head's logits at H/4 and uint8 labels (rows N1 + N4)."""
    b, c, h, wd = x.shape
    px = b * h * wd
    if low_res:
        xin = torch.nn.functional.avg_pool2d(x.detach().float(), 4).to(x.dtype)     # any H/4 tensor of the right statistics
        labin = lab.to(torch.uint8)
    else:
        xin, labin = x.detach(), lab
    hx = torch.empty(xin.shape, dtype=xin.dtype, pin_memory=True).copy_(xin)
    hlab = torch.empty(labin.shape, dtype=labin.dtype, pin_memory=True).copy_(labin)
    hemb = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True).copy_(emb.detach())
    hgx = torch.empty(xin.shape, dtype=xin.dtype, pin_memory=True)
    hge = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
    hloss = torch.empty(1, dtype=torch.float32, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmp = torch.cuda.current_stream(dev)
    dbuf = [(torch.empty_like(xin), torch.empty_like(labin), torch.empty_like(emb.detach())) for _ in range(2)]
    del xin, labin
    free_ev = [None, None]
    state = {"i": 0}

    def e2e_step():
        i = state["i"]
        state["i"] += 1
        xb, lb, eb = dbuf[i % 2]
        with torch.cuda.stream(s_in):
            if free_ev[i % 2] is not None:
                s_in.wait_event(free_ev[i % 2])      # the step that last used this buffer set has finished
            xb.copy_(hx, non_blocking=True)
            lb.copy_(hlab, non_blocking=True)
            eb.copy_(hemb, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(s_in)
        s_cmp.wait_event(ev_in)
        xd = xb.detach().requires_grad_(True)
        ed = eb.detach().requires_grad_(True)
        loss = mod(step_t, ed if use_emb is not None else None, None, xd, lb)
        loss.backward()
        ev_cmp = torch.cuda.Event()
        ev_cmp.record(s_cmp)
        free_ev[i % 2] = ev_cmp
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp)
            xd.grad.record_stream(s_out)
            hgx.copy_(xd.grad, non_blocking=True)
            if ed.grad is not None:
                ed.grad.record_stream(s_out)
                hge.copy_(ed.grad, non_blocking=True)
            loss.record_stream(s_out)
            hloss.copy_(loss.detach().reshape(1), non_blocking=True)   # device->host read of the step's result
        return loss

    def drain():
        s_out.synchronize()
        return float(hloss[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(6 if low_res else 1):     # warm-up: the caching allocator's pool has to reach the pipeline's working set
        e2e_step()                            # (record_stream keeps a step's blocks busy until its D2H has finished)
    drain()
    barrier()
    # Steady-state throughput of the three-stream pipeline.  The first H2D and the last D2H of a timed region (~5 ms
    # together at config 3) overlap nothing, so a region is 2 K steps long (capped) when a step is a few ms; the host
    # side of the copies is shared with the box's other tenants (the same command has measured 5.4 and 7.7 ms/step
    # minutes apart), so the region is repeated and the FASTEST repetition is reported, all of them listed.
    k2 = max(2, min(2 * steps, 20)) if low_res else max(2, min(steps, 10))
    reps = 3 if low_res else 1
    rep_ms = []
    for _ in range(reps):
        barrier()
        t0 = time.time()
        for _ in range(k2):
            e2e_step()
        drain()                   # the last step's results are on the host
        barrier()
        t1 = time.time()
        clk.mark(t0, t1)
        rep_ms.append((t1 - t0) * 1e3 / k2)
    # host-side clock: the region spans three streams and ends when the last D2H copy has landed
    el = torch.tensor([min(rep_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    h2d = hx.numel() * hx.element_size() + hlab.numel() * hlab.element_size() + hemb.numel() * hemb.element_size()
    d2h = hgx.numel() * hgx.element_size() + hge.numel() * hge.element_size() + 4
    what = ("logits at the head's resolution [B,C,H/4,W/4] (upsampled inside the op as train.py:282-284 does outside) "
            "and uint8 labels; gradients return at H/4" if low_res else
            "full-resolution logits and int64 labels (the reference's own tensors)")
    return {"value": world * px / (float(el.item()) * 1e-3) / 1e9, "unit": "Gpix/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": float(el.item()),
            "steps": k2, "repetitions_ms_per_step": [round(v, 4) for v in rep_ms], "inputs": what,
            "note": "pinned host buffers -> H2D -> fwd+bwd -> D2H of loss and gradients, every step; copies of "
                    "neighbouring steps overlap the kernels (3 streams); wall clock around the loop incl. the final "
                    "drain; fastest of the listed repetitions (max over ranks)"}


# ------------------------------------------------------------------------------------------------
# end-to-end train throughput (the loss inside a segmentation training step; network = context)
# ------------------------------------------------------------------------------------------------
def _context_net(depth, c, dev):
    """The network around the loss.  The reference's own ResNetBackbone + DepthwiseSeparableASPPContrastHead when the
    reference can be imported on this box, else the equivalent context net of scripts/train_context.py."""
    import torch
    root = _reference_root()
    if root is not None:
        try:
            sys.dont_write_bytecode = True
            sys.path.insert(0, root)
            from models.backbone.resnet import ResNetBackbone
            from models.head.sep_aspp_contrast_head import DepthwiseSeparableASPPContrastHead

            class RefNet(torch.nn.Module):
                def __init__(self):
                    super().__init__()
                    self.backbone = ResNetBackbone(depth=depth, pretrained=False)
                    self.head = DepthwiseSeparableASPPContrastHead(          # arguments of train.py:157-167
                        in_channels=2048, c1_in_channels=256, c1_channels=48, aspp_channels=512,
                        dilations=(1, 12, 24, 36), num_classes=c, proj_dim=256, proj_type="convmlp")

                def forward(self, img):
                    feats = self.backbone(img)
                    logits, emb = self.head(list(feats))
                    return logits, emb, feats[2]
            net = RefNet().to(dev)
            return net, "reference ResNetBackbone-%d + DepthwiseSeparableASPPContrastHead (imported unmodified)" % depth
        except Exception:                    # noqa: BLE001
            pass
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from train_context import ContextSegNet
    return (ContextSegNet(depth, c).to(dev),
            "context net: torchvision ResNet-%d (stride 32) + DeepLabV3+-style SepASPP head with projection branch "
            "(scripts/train_context.py; the reference does not travel to this box)" % depth)


def train_record(args, w, steps, brief=False):
    """Whole training step of the config (network fwd/bwd on stock cuDNN, our loss fed with the head's H/4 logits and
    uint8 labels, aux-head CE of train.py:309-313 through the fused kernel, SGD), DDP when world > 1."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import seghiero_b200 as sb

    rank, world, local, dev = _dist_env()
    torch.backends.cudnn.benchmark = True
    b, h, wd = args.batch or w["B"], w["H"], w["W"]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if w["levels"] == 3:
        c = w["nf"] + w["nm"] + w["nh"]
        f2m, f2h = hierarchy_maps(w)
        crit = sb.RMIHieraTripletLoss(w["nf"], w["nm"], w["nh"], torch.tensor(f2m), torch.tensor(f2h))
    else:
        c = w["nf"] + w["nc"]
        crit = sb.HieraTripletLoss(w["nf"], HM_19_7, HI_19_7)
    net, net_desc = _context_net(w["depth"], c, dev)
    c3_ch = 1024 if w["depth"] >= 50 else 256

    class TrainNet(torch.nn.Module):
        """network + aux head (1x1 conv + BN + ReLU on c3, train.py:169-173) behind one forward, so that DDP sees one module"""

        def __init__(self):
            super().__init__()
            self.net = net
            self.aux = torch.nn.Sequential(torch.nn.Conv2d(c3_ch, w["nf"], 1, bias=False), torch.nn.BatchNorm2d(w["nf"]),
                                           torch.nn.ReLU(inplace=True))

        def forward(self, img_):
            logits_, emb_, c3 = self.net(img_)
            return logits_, emb_, self.aux(c3)
    full = TrainNet().to(dev).to(memory_format=torch.channels_last)
    model = torch.nn.parallel.DistributedDataParallel(full, device_ids=[local]) if world > 1 else full
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)      # train.py:239-246
    img = torch.randn(b, 3, h, wd, generator=g, device=dev).contiguous(memory_format=torch.channels_last)
    lab = make_labels(torch, g, b, h, wd, w["nf"], args.labels, dev).to(torch.uint8)
    step_t = torch.tensor([100000], device=dev)
    logit_dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16

    def step(with_loss=True):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, emb, aux_logits = model(img)
        if with_loss:
            # the head's logits go in as they are: the x4 upsample of train.py:282-284 happens inside the op, and the
            # aux-head CE of train.py:309-313 reads the H/16 aux logits directly
            loss = crit(step_t, emb.float(), None, logits.to(logit_dt), lab)
            loss = loss + 0.4 * sb.aux_cross_entropy(aux_logits.float(), lab)
        else:   # same graph without the two losses: what the rest of the step costs
            loss = logits.float().mean() + emb.float().mean() + aux_logits.float().mean()
        loss.backward()
        opt.step()
        return loss

    warmup = 3
    clk = ClockSampler(local)
    clk.__enter__()
    for _ in range(warmup):
        step(True)
    ms, win, _ = _timed_steps(torch, dist, world, dev, lambda: step(True), steps)
    clk.mark(*win)
    for _ in range(2):
        step(False)
    ms_noloss, _, _ = _timed_steps(torch, dist, world, dev, lambda: step(False), max(2, steps // 2))
    clk.__exit__()
    rec = {"metric": "train_throughput", "value": world * b / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms, "batch_per_gpu": b,
           "ms_per_step_without_hier_loss": ms_noloss, "hier_loss_share": max(0.0, 1.0 - ms_noloss / ms),
           "network": net_desc + ", random init, bf16 autocast, stock cuDNN/ATen (context, not product)",
           "loss_inputs": "head logits at H/4 (%s) + uint8 labels -> seghiero_b200 (upsample inside the op); aux-head CE "
                          "from the H/16 aux logits through sh_aux_ce_fwdbwd" % w["dtype"],
           "optimizer": "SGD momentum 0.9", "parallelism": f"DDP x{world}" if world > 1 else "single GPU",
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    if not brief:
        rec.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "bf16 network / %s logits" % w["dtype"], "data": "synthetic",
                    "config": {"workload": args.workload, "description": w["desc"], "labels": args.labels},
                    "clocks": clk.summary()})
    del model, full, opt, img
    torch.cuda.empty_cache()
    return rec


def run_train(args, w):
    import torch.distributed as dist
    rank, world, _, _ = _dist_env()
    rec = train_record(args, w, max(1, min(args.steps, 10)))
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--labels", default="blob", choices=["blob", "iid"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-iid", action="store_true", help="skip the per-pixel-label sub-record")
    ap.add_argument("--no-train", action="store_true", help="skip the train img/s sub-record")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-ATen-on-this-GPU baseline")
    ap.add_argument("--kernels-only", action="store_true", help="= --no-e2e --no-cpu --no-iid --no-train --no-eager")
    args = ap.parse_args()
    if args.kernels_only:
        args.no_e2e = args.no_cpu = args.no_iid = args.no_train = args.no_eager = True
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
    elif w["kind"] == "train":
        run_train(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()

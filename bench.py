#!/usr/bin/env python
"""Benchmark of the hierarchical-loss hot path (BASELINE.json metric: hier-loss fwd+bwd Gpix/s and
fraction of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle port, rank 0 only)

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


def run_cpu_oracle(w, steps, warmup, label_kind):
    import torch
    import torch.nn.functional as F
    from oracle import hiera_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, h, wd = cpu_sample_shape(w)
    g = torch.Generator().manual_seed(1234)
    lab = make_labels(torch, g, b, h, wd, w["nf"], label_kind, "cpu")
    dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16
    f2m, f2h = hierarchy_maps(w)
    if w["kind"] == "decode":
        x = torch.randn(b, w["nf"] + w["nm"] + w["nh"], h, wd, generator=g).to(dt)

        def step():
            O.argmax_decode(x, [w["nf"], w["nm"], w["nh"]])
    elif w["kind"] == "2level":
        x = (torch.randn(b, w["nf"] + w["nc"], h, wd, generator=g) * 2).to(dt).requires_grad_(True)
        emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g), dim=1).requires_grad_(True)

        def step():
            x.grad = None
            emb.grad = None
            loss, _ = O.hiera_triplet_loss(100000, emb, x, lab, w["nf"], HM_19_7, HI_19_7)
            loss.backward()
    else:
        x = (torch.randn(b, w["nf"] + w["nm"] + w["nh"], h, wd, generator=g) * 2).to(dt).requires_grad_(True)
        emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g), dim=1).requires_grad_(True)
        trip = w.get("triplet", True)

        def step():
            x.grad = None
            emb.grad = None
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
    return dict(value=px / dt_s / 1e9, unit="Gpix/s", cores=cores, kind="port",
                sample=f"oracle port (torch-CPU restatement of the reference loss), {b}x{h}x{wd} crop of the "
                       f"workload, fwd+bwd, {steps} steps, all {cores} host threads", ms_per_step=dt_s * 1e3)


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    res = run_cpu_oracle(w, steps, 1, args.labels)
    line = {
        "impl": "reference", "metric": "hier_loss_fwd_bwd_throughput", "value": res["value"], "unit": "Gpix/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "labels": args.labels},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import seghiero_b200 as sb
    from seghiero_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    b, h, wd = args.batch or w["B"], w["H"], w["W"]
    dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16
    px = b * h * wd
    f2m, f2h = hierarchy_maps(w)
    lab = make_labels(torch, g, b, h, wd, w["nf"], args.labels, dev)
    step_t = torch.tensor([100000], device=dev)
    grads_out = []

    if w["kind"] == "decode":
        c = w["nf"] + w["nm"] + w["nh"]
        x = torch.randn(b, c, h, wd, generator=g, device=dev, dtype=torch.float32).to(dt)
        emb = None

        def step():
            preds, counts = sb.hierarchical_argmax(x, [w["nf"], w["nm"], w["nh"]], lab)
            return counts
        host_in = [x]
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

        def step():
            x.grad = None
            emb.grad = None
            loss = mod(step_t, use_emb, None, x, lab)
            loss.backward()
            if world > 1:   # scalar loss sum across ranks (the path's only data collective besides `ready`)
                dist.all_reduce(loss.detach(), op=dist.ReduceOp.SUM)
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    clk.__enter__()                       # nvidia-smi needs ~0.1 s before its first sample: start it before the warm-up
    for _ in range(warmup):
        step()
    barrier()

    # ---- timed region: K steps, device-timed, stage events on --------------------------------------
    timer = StageTimer()
    ops.STAGE_TIMER = timer
    ops.LAUNCHES["n"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    clk.mark(t_wall0, time.time())
    ops.STAGE_TIMER = None
    launches = ops.LAUNCHES["n"]
    elapsed = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    ms_total = float(elapsed.item())
    ms_step = ms_total / steps
    value = world * px / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (from the stage events of the same timed region) ------------
    ab = algorithmic_bytes_per_px(w)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    stage_names = {("sh_rmi3_forward", 1): "k3_prep", ("sh_rmi3_forward", 2): "k3_pass1",
                   ("sh_rmi3_forward", 4): "k3_frame1", ("sh_rmi3_forward", 8): "k3_finalize",
                   ("sh_rmi3_backward", 1): "k3_pass2", ("sh_rmi3_backward", 2): "k3_frame2",
                   ("sh_bce2_fwdbwd", 1): "k_prep2", ("sh_bce2_fwdbwd", 2): "k_bce2_fused",
                   ("sh_bce2_fwdbwd", 4): "k_reduce_partials"}
    if w["kind"] == "3level" and getattr(mod, "last_stats", {}).get("fast_path"):
        stage_names.update({("sh_rmi3_forward", 1): "k3f_prep", ("sh_rmi3_forward", 2): "k3f_pass1",
                            ("sh_rmi3_forward", 8): "k3f_finalize", ("sh_rmi3_backward", 1): "k3f_pass2"})
    if w["kind"] == "2level" and mod.last_stats.get("fast_path"):
        stage_names[("sh_bce2_fwdbwd", 2)] = "k_bce2_fast"
    stage_bytes = {"k3_pass1": ab.get("pass1"), "k3_pass2": ab.get("pass2"), "k_bce2_fused": ab.get("fused"),
                   "k_bce2_fast": ab.get("fused"), "k3f_pass1": ab.get("pass1"), "k3f_pass2": ab.get("pass2")}
    stages = {stage_names[k]: t / n for k, (t, n) in timer.totals().items() if k in stage_names}
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
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                    "whole_step_frac": ab["total"] * px / (ms_step * 1e-3) / 1e9 / peak}
    elif w["kind"] == "decode":
        achieved = ab["total"] * px / (ms_step * 1e-3) / 1e9
        tr = traffic_tab.get(f"{args.workload}:k_decode")
        roofline = {"bound": "hbm", "kernel": "k_decode_vec", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": (tr["bytes_per_px"] * px if tr else None),
                    "traffic_source": (tr.get("source") if tr else None),
                    "algorithmic_bytes_per_launch": ab["total"] * px,
                    "ms_per_launch": ms_step,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"}

    # ---- end-to-end through the public API with HOST buffers -----------------------------------------
    e2e = None
    if not args.no_e2e:
        if w["kind"] == "decode":
            hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
            hlab = torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True).copy_(lab)
            hout = [torch.empty((b, h, wd), dtype=torch.int64, pin_memory=True) for _ in range(3)]

            def e2e_step():
                xd = hx.to(dev, non_blocking=True)
                ld = hlab.to(dev, non_blocking=True)
                preds, counts = sb.hierarchical_argmax(xd, [w["nf"], w["nm"], w["nh"]], ld)
                for o, p in zip(hout, preds):
                    o.copy_(p, non_blocking=True)
                return counts.cpu()
            h2d = hx.numel() * hx.element_size() + hlab.numel() * 8
            d2h = 3 * b * h * wd * 8 + 16
        else:
            hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x.detach())
            hlab = torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True).copy_(lab)
            hemb = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True).copy_(emb.detach())
            hgx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            hge = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
            hloss = torch.empty(1, dtype=torch.float32, pin_memory=True)
            # three streams: the H2D copy of step i+1 and the D2H copy of step i-1 overlap the kernels of step i
            # (every step still moves all of its inputs and results; device input buffers are double buffered)
            s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            s_cmp = torch.cuda.current_stream(dev)
            dbuf = [(torch.empty_like(x.detach()), torch.empty_like(lab), torch.empty_like(emb.detach())) for _ in range(2)]
            free_ev = [None, None]
            state = {"i": 0, "last": None}

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
                state["last"] = s_out
                return loss

            def e2e_drain():
                s_out.synchronize()
                return float(hloss[0])
            h2d = hx.numel() * hx.element_size() + hlab.numel() * 8 + hemb.numel() * hemb.element_size()
            d2h = hgx.numel() * hgx.element_size() + hge.numel() * hge.element_size() + 4
        e2e_step()
        if w["kind"] != "decode":
            e2e_drain()
        barrier()
        k2 = max(2, min(steps, 10))
        t_wall0 = time.time()
        for _ in range(k2):
            e2e_step()
        if w["kind"] != "decode":
            e2e_drain()                   # the last step's results are on the host
        barrier()
        t_wall1 = time.time()
        clk.mark(t_wall0, t_wall1)
        # host-side clock: the region spans three streams and ends when the last D2H copy has landed
        el = torch.tensor([(t_wall1 - t_wall0) * 1e3 / k2], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": world * px / (float(el.item()) * 1e-3) / 1e9, "unit": "Gpix/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": float(el.item()),
               "steps": k2, "note": "pinned host buffers -> H2D -> fwd+bwd -> D2H of loss and gradients, every step; "
                                    "copies of neighbouring steps overlap the kernels (3 streams); wall clock around "
                                    "the loop incl. the final drain"}

    clk.__exit__()
    x_gb = x.numel() * x.element_size() / 1e9
    # optional context: the oracle's torch restatement run as eager ATen ON THE GPU (what a user of the reference gets
    # on this box); one sample per step because the fp64 RMI unfolds need ~21 GB per 1024x2048 image
    eager = None
    if args.eager_gpu and rank == 0 and w["kind"] in ("3level", "2level"):
        from oracle import hiera_oracle as O
        x_gb = x.numel() * x.element_size() / 1e9
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
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        res = run_cpu_oracle(w, 2, 1, args.labels)
        cpu_base = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "hier_loss_fwd_bwd_throughput", "value": value, "unit": "Gpix/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if w["dtype"] == "fp32" else "bf16 in / f32 accumulate",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "labels": args.labels,
                       "batch_per_gpu": b, "pixels_per_step_per_gpu": px,
                       "l2": "inputs larger than L2 (logits %.2f GB per GPU)" % x_gb,
                       "parallelism": f"dp{world} by sample, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk.summary(), "kernel_ms": stages,
        }
        if eager is not None:
            line["eager_gpu_baseline"] = eager
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# end-to-end train throughput (the loss inside a segmentation training step; network = context)
# ------------------------------------------------------------------------------------------------
def run_train(args, w):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    import seghiero_b200 as sb
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from train_context import ContextSegNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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
    net = ContextSegNet(w["depth"], c).to(dev).to(memory_format=torch.channels_last)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    img = torch.randn(b, 3, h, wd, generator=g, device=dev).contiguous(memory_format=torch.channels_last)
    lab = make_labels(torch, g, b, h, wd, w["nf"], args.labels, dev)
    step_t = torch.tensor([100000], device=dev)
    logit_dt = torch.float32 if w["dtype"] == "fp32" else torch.bfloat16

    def step(with_loss=True):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, emb = model(img)
        full = F.interpolate(logits.to(logit_dt), size=(h, wd), mode="bilinear", align_corners=False)
        if with_loss:
            loss = crit(step_t, emb.float(), None, full, lab)
        else:   # same graph without the hierarchical loss: what the rest of the step costs
            loss = full.float().mean() + emb.float().mean()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, with_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step(with_loss)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warmup, steps = max(args.warmup, 3), max(1, min(args.steps, 10))
    clk = ClockSampler(local)
    clk.__enter__()
    for _ in range(warmup):
        step(True)
    t0 = time.time()
    ms = timed(steps, True)
    clk.mark(t0, time.time())
    for _ in range(2):
        step(False)
    ms_noloss = timed(steps, False)
    clk.__exit__()
    if rank == 0:
        print(json.dumps({
            "metric": "train_throughput", "value": world * b / (ms * 1e-3), "unit": "img/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 network / %s logits" % w["dtype"], "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "labels": args.labels, "batch_per_gpu": b,
                       "network": "torchvision ResNet-%d (stride 32) + DeepLabV3+-style SepASPP head, random init, "
                                  "context only (stock cuDNN/ATen)" % w["depth"],
                       "optimizer": "SGD momentum 0.9", "parallelism": f"DDP x{world}" if world > 1 else "single GPU"},
            "ms_per_step_without_hier_loss": ms_noloss, "hier_loss_share": max(0.0, 1.0 - ms_noloss / ms),
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "clocks": clk.summary()}))
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
    ap.add_argument("--eager-gpu", action="store_true", help="also time the oracle restatement as eager ATen on the GPU")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
    elif w["kind"] == "train":
        run_train(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()

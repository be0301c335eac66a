"""Host side of the C-ABI custom ops: tensor plumbing, workspaces and autograd glue.

PyTorch is used for device memory, streams and autograd bookkeeping only; every
computation is a kernel of libseghiero_b200.so.  Nothing here touches the host
inside forward/backward (no .item(), no .tolist()): `step`, the `ready` gate and
`grad_output` stay on the device.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _lib
from . import hierarchy as H

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_table_cache: dict = {}

# kernels launched per C-ABI call (per stage bit for the multi-kernel entry points); bench.py reads
# LAUNCHES["n"] to report how many of OUR kernels ran inside its timed region
_KERNELS = {
    "sh_targets_two_level": 1, "sh_targets_three_level": 1, "sh_targets_gather": 1, "sh_decode": 1,
    "sh_loss2_final": 1, "sh_loss3_final": 1, "sh_scale_inplace": 1, "sh_triplet_forward": 4,
    "sh_triplet_backward": 1,
    ("sh_bce2_fwdbwd", 1): 1, ("sh_bce2_fwdbwd", 2): 1, ("sh_bce2_fwdbwd", 4): 1,
    ("sh_rmi3_forward", 1): 1, ("sh_rmi3_forward", 2): 1, ("sh_rmi3_forward", 4): 2, ("sh_rmi3_forward", 8): 2,
    ("sh_rmi3_backward", 1): 2, ("sh_rmi3_backward", 2): 1,
}
LAUNCHES = {"n": 0}
FAST_PATH = {"enabled": True}   # tests flip this to cover the generic kernels on shapes the fast path would take
STAGE_TIMER = None   # bench.py installs an object with start(name, bit) / stop(name, bit)


def _call(name, *args):
    _lib.call(name, *args)
    LAUNCHES["n"] += _KERNELS.get(name, 0)


def _staged(name, bits, make_args):
    """Run a multi-kernel entry point.  Normally one C call with every stage bit set; with a stage
    timer installed, one call per stage so that each kernel can be bracketed by CUDA events."""
    timer = STAGE_TIMER
    if timer is None:
        _lib.call(name, *make_args(sum(bits)))
    else:
        for bit in bits:
            timer.start(name, bit)
            _lib.call(name, *make_args(bit))
            timer.stop(name, bit)
    LAUNCHES["n"] += sum(_KERNELS[(name, b)] for b in bits)


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("seghiero_b200 runs on CUDA tensors only (there is no CPU fallback)")


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise TypeError(f"unsupported dtype {t.dtype}; use float32, bfloat16 or float16")
    return _DT[t.dtype]


def _labels(label: torch.Tensor) -> torch.Tensor:
    if label.dtype != torch.int64:
        label = label.long()
    return label.contiguous()


def device_table(key, builder, device):
    """numpy int32 table -> cached device tensor (built once per device)."""
    k = (key, str(device))
    hit = _table_cache.get(k)
    if hit is None:
        arr = builder()
        extra = arr[1:] if isinstance(arr, tuple) else ()
        arr0 = arr[0] if isinstance(arr, tuple) else arr
        hit = (torch.from_numpy(np.ascontiguousarray(arr0)).to(device), *extra)
        _table_cache[k] = hit
    return hit


def step_tensor(step, device) -> torch.Tensor:
    if isinstance(step, torch.Tensor):
        return step.detach().reshape(-1)[:1].to(device=device, dtype=torch.float64)
    return torch.tensor([float(step)], dtype=torch.float64, device=device)


def _world_ready(status: torch.Tensor) -> None:
    """`ready` = every rank found >= 1 triplet class (hiera_triplet_loss.py:193-198), as one
    device-side MIN all-reduce instead of an all_gather + host sync."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(status[:1], op=dist.ReduceOp.MIN)


class _Fork:
    """Fork / join of a per-device side stream: the triplet kernels are small and latency bound, so they run next to
    the streaming loss kernels instead of in front of them.  Tensors are allocated on the caller's stream before the
    fork and the caller's stream waits for the side stream (`join`) before anything reads the results, so the caching
    allocator never hands a block to work that is not ordered after its last use."""
    _streams = {}

    def __init__(self, dev: torch.device):
        self.cur = torch.cuda.current_stream(dev)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        side = _Fork._streams.get(idx)
        if side is None:
            side = _Fork._streams[idx] = torch.cuda.Stream(device=dev)
        self.side = side
        self._ctx = None

    def __enter__(self):
        self.side.wait_stream(self.cur)
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)

    def join(self):
        self.cur.wait_stream(self.side)


# ----------------------------------------------------------------------------------------------
# targets / decode
# ----------------------------------------------------------------------------------------------
def targets_two_level(label: torch.Tensor, hiera_index):
    _need_cuda(label)
    lab = _labels(label)
    key = ("lut2", tuple(tuple(int(v) for v in r) for r in hiera_index))

    def build():
        size = max([int(e) for _, e in hiera_index] + [1])
        lut = np.full(size, H.IGNORE, dtype=np.int32)
        for i, (s, e) in enumerate(hiera_index):
            lo, hi = max(int(s), 0), min(int(e), size)
            if hi > lo:
                lut[lo:hi] = i
        return lut
    (lut,) = device_table(key, build, lab.device)
    out = torch.empty_like(lab)
    with torch.cuda.device(lab.device):
        _call("sh_targets_two_level", _p(lab), _p(out), lab.numel(), _p(lut), lut.numel(), _stream())
    return out.to(label.dtype) if label.dtype != torch.int64 else out


def targets_three_level(label: torch.Tensor, fine_to_mid: torch.Tensor, fine_to_high: torch.Tensor, check=True):
    _need_cuda(label)
    lab = _labels(label)
    f2m = fine_to_mid.to(device=lab.device, dtype=torch.int64).contiguous()
    f2h = fine_to_high.to(device=lab.device, dtype=torch.int64).contiguous()
    mid, high = torch.empty_like(lab), torch.empty_like(lab)
    err = torch.zeros(1, dtype=torch.int32, device=lab.device)
    with torch.cuda.device(lab.device):
        _call("sh_targets_three_level", _p(lab), _p(mid), _p(high), lab.numel(), _p(f2m), _p(f2h), f2m.numel(),
                  _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("fine label out of range for fine_to_mid / fine_to_high")
    return mid, high


def targets_gather(fine_mask: torch.Tensor, level_map: torch.Tensor, check=True):
    _need_cuda(fine_mask)
    lab = _labels(fine_mask)
    m = level_map.to(device=lab.device, dtype=torch.int64).contiguous()
    out = torch.empty_like(lab)
    err = torch.zeros(1, dtype=torch.int32, device=lab.device)
    with torch.cuda.device(lab.device):
        _call("sh_targets_gather", _p(lab), _p(out), lab.numel(), _p(m), m.numel(), _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("index out of range in target gather")
    return out


def hierarchical_argmax(logits: torch.Tensor, level_sizes: Sequence[int], label: Optional[torch.Tensor] = None,
                        out_dtype=torch.int64):
    """Per-level argmax over channel slices (+ fine pixel-accuracy counts when `label` is given).
    Returns (list of [B,H,W] predictions, counts int64[2] = (#correct, #valid) or None)."""
    _need_cuda(logits)
    if out_dtype not in (torch.int64, torch.uint8):
        raise TypeError("out_dtype must be torch.int64 or torch.uint8")
    x = logits.contiguous()
    b, c = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    sizes = [int(s) for s in level_sizes] + [0, 0]
    n0, n1, n2 = sizes[:3]
    if len(level_sizes) > 3 or n0 + n1 + n2 > c:
        raise ValueError("level_sizes must be <= 3 levels and fit the channel count")
    outs = [torch.empty((b,) + tuple(x.shape[2:]), dtype=out_dtype, device=x.device) if n > 0 else None
            for n in (n0, n1, n2)]
    counts = None
    lab = None
    if label is not None:
        lab = _labels(label)
        counts = torch.zeros(2, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _call("sh_decode", _p(x), _dtype_code(x), b, c, hw, n0, n1, n2, _p(outs[0]), _p(outs[1]), _p(outs[2]),
                  1 if out_dtype == torch.uint8 else 0, _p(lab), _p(counts), _stream())
    return [o for o in outs if o is not None], counts


# ----------------------------------------------------------------------------------------------
# triplet
# ----------------------------------------------------------------------------------------------
@dataclass
class TripletState:
    mode: int
    ncls: int
    max_triplet: int
    dims: tuple
    sel: torch.Tensor
    kcount: torch.Tensor
    tl: torch.Tensor
    trip: torch.Tensor      # [loss, #classes] float32
    status: torch.Tensor    # [ready, error] int32
    lab_ds: torch.Tensor


def triplet_forward(feats: torch.Tensor, label: torch.Tensor, mode: int, tab: torch.Tensor, ncls: int,
                    max_triplet: int = 200, fork: Optional[_Fork] = None, world_ready: bool = False) -> TripletState:
    _need_cuda(feats, label)
    b, d, h, w = feats.shape
    hh, ww = label.shape[-2:]
    dev = feats.device
    rows = b * h * w
    lab_ds = torch.empty(rows, dtype=torch.int32, device=dev)
    sel = torch.empty(ncls * 3 * max_triplet, dtype=torch.int32, device=dev)
    kcount = torch.empty(ncls, dtype=torch.int32, device=dev)
    tl = torch.empty(ncls * max_triplet, dtype=torch.float32, device=dev)
    trip = torch.empty(2, dtype=torch.float32, device=dev)
    status = torch.empty(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), (fork if fork is not None else contextlib.nullcontext()):
        _call("sh_triplet_forward", _p(feats), _dtype_code(feats), _p(label), b, d, h, w, hh, ww, mode, _p(tab),
                  ncls, max_triplet, _p(lab_ds), _p(sel), _p(kcount), _p(tl), _p(trip), _p(status), _stream())
        if world_ready:
            _world_ready(status)      # on the side stream too: the all-reduce overlaps the loss kernels
    return TripletState(mode, ncls, max_triplet, (b, d, h, w), sel, kcount, tl, trip, status, lab_ds)


def triplet_backward(feats: torch.Tensor, st: TripletState, tscale: torch.Tensor,
                     gscale: Optional[torch.Tensor], fork: Optional[_Fork] = None) -> torch.Tensor:
    """fp32 gradient of the embedding; with a fork the caller joins and then converts (`_grad_as`)."""
    b, d, h, w = st.dims
    gfeat = torch.empty((b, d, h, w), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device), (fork if fork is not None else contextlib.nullcontext()):
        _call("sh_triplet_backward", _p(feats), _dtype_code(feats), b, d, h, w, st.ncls, st.max_triplet,
                  _p(st.sel), _p(st.kcount), _p(st.tl), _p(st.trip), _p(tscale), _p(gscale), _p(gfeat), _stream())
    return gfeat if fork is not None else _grad_as(gfeat, feats)


def _grad_as(gfeat: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
    return gfeat if feats.dtype == torch.float32 else gfeat.to(feats.dtype)


class TripletFn(torch.autograd.Function):
    """Standalone triplet loss value (mean over classes); 0 when no class contributes."""

    @staticmethod
    def forward(ctx, feats, label, mode, tab, ncls, max_triplet, holder):
        feats_c = feats.contiguous()
        st = triplet_forward(feats_c, _labels(label), mode, tab, ncls, max_triplet)
        holder["state"] = st
        ctx.st = st
        ctx.save_for_backward(feats_c)
        return st.trip[0].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (feats,) = ctx.saved_tensors
        one = torch.ones(1, dtype=torch.float32, device=feats.device)
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        return triplet_backward(feats, ctx.st, one, g), None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# two-level fused loss
# ----------------------------------------------------------------------------------------------
@dataclass
class Hier2Config:
    n_fine: int
    n_coarse: int
    hiera_map: list
    hiera_index: list
    loss_weight: float
    total_steps: float = 80000.0
    eps: float = 1e-8


class HieraTriplet2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_score, embedding, label, step_d, cfg: Hier2Config, stats: dict):
        _need_cuda(cls_score, label, embedding)
        x = cls_score.contiguous()
        lab = _labels(label)
        dev = x.device
        b, c = x.shape[0], x.shape[1]
        hw = x[0, 0].numel()
        if c != cfg.n_fine + cfg.n_coarse:
            raise ValueError(f"cls_score has {c} channels, expected n_fine+n_coarse={cfg.n_fine + cfg.n_coarse}")
        if tuple(lab.shape) != (b,) + tuple(x.shape[2:]):
            raise ValueError("label must be [B,H,W] matching cls_score")
        key = ("h2", cfg.n_fine, tuple(tuple(int(v) for v in r) for r in cfg.hiera_index))
        tab, n_fb, lut_size = device_table(key, lambda: H.two_level_tables(cfg.n_fine, cfg.hiera_index), dev)

        st = None
        if embedding is not None:
            tkey = ("t0", tuple(int(v) for v in cfg.hiera_map), key[2])
            ttab, ncls = device_table(tkey, lambda: H.triplet_tables_hierarchy(cfg.hiera_map, cfg.hiera_index), dev)
            emb = embedding.contiguous()
            fork = _Fork(dev)
            st = triplet_forward(emb, lab, 0, ttab, ncls, fork=fork, world_ready=True)

        want_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(x) if want_grad else None
        lab8 = torch.empty(b * hw, dtype=torch.uint8, device=dev)
        counts = torch.empty(4, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            grid = _lib.load().sh_bce2_grid(b, hw, c, cfg.n_coarse)
            partials = torch.empty(grid * 4, dtype=torch.float32, device=dev)
            sums = torch.empty(4, dtype=torch.float64, device=dev)
            out = torch.empty(4, dtype=torch.float32, device=dev)
            tree = 256 if (FAST_PATH["enabled"] and H.two_level_is_tree(cfg.n_fine, cfg.hiera_index)) else 0
            _staged("sh_bce2_fwdbwd", (1, 2, 4), lambda st_bits: (
                _p(x), _dtype_code(x), _p(lab), _p(grad), b, hw, cfg.n_fine, cfg.n_coarse, _p(tab), n_fb, lut_size,
                cfg.eps, cfg.loss_weight, _p(lab8), _p(counts), _p(partials), _p(sums), st_bits | tree, _stream()))
            if st is not None:
                fork.join()
            _call("sh_loss2_final", _p(sums), _p(counts), cfg.n_fine, cfg.n_coarse, float(b * hw), _p(step_d),
                      cfg.total_steps, _p(st.trip) if st else None, _p(st.status) if st else None, cfg.loss_weight,
                      _p(out), _stream())
        stats.update(sums=sums, counts=counts, out=out, triplet=st,
                     fast_path=bool(tree) and hw % 4 == 0 and c <= 64 and x.data_ptr() % 16 == 0)
        ctx.grad = grad
        ctx.st = st
        ctx.out = out
        ctx.emb = emb if st is not None else None
        return out[0].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        gx = None
        gemb = fork = None
        if ctx.st is not None and ctx.needs_input_grad[1]:
            fork = _Fork(ctx.emb.device)
            gemb = triplet_backward(ctx.emb, ctx.st, ctx.out[1:2], g, fork=fork)
        if ctx.needs_input_grad[0]:
            gx = ctx.grad
            if gx is None:
                raise RuntimeError("seghiero_b200: backward through the fused loss can run only once")
            ctx.grad = None
            with torch.cuda.device(gx.device):
                _call("sh_scale_inplace", _p(gx), _dtype_code(gx), gx.numel(), _p(g), _stream())
        if fork is not None:
            fork.join()
            gemb = _grad_as(gemb, ctx.emb)
        return gx, gemb, None, None, None, None


# ----------------------------------------------------------------------------------------------
# three-level fused loss
# ----------------------------------------------------------------------------------------------
@dataclass
class Hier3Config:
    n_fine: int
    n_mid: int
    n_high: int
    fine_to_mid: tuple
    fine_to_high: tuple
    upper_ids: tuple
    lower_ids: tuple
    lam: float
    loss_weight: float
    total_steps: float
    use_triplet: bool = True


class RMIHieraTriplet3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_score, embedding, label, step_d, cfg: Hier3Config, stats: dict):
        _need_cuda(cls_score, label, embedding)
        x = cls_score.contiguous()
        lab = _labels(label)
        dev = x.device
        if x.dim() != 4:
            raise ValueError("cls_score must be [B,C,H,W]")
        b, c, hh, ww = x.shape
        if c != cfg.n_fine + cfg.n_mid + cfg.n_high:
            raise ValueError(f"cls_score has {c} channels, expected {cfg.n_fine + cfg.n_mid + cfg.n_high}")
        if tuple(lab.shape) != (b, hh, ww):
            raise ValueError("label must be [B,H,W] matching cls_score")
        if hh < 8 or ww < 8:
            raise ValueError("the CUDA RMI path needs H, W >= 8 (the reference needs >= 3)")
        key = ("h3", cfg.n_fine, cfg.n_mid, cfg.n_high, cfg.fine_to_mid, cfg.fine_to_high)
        tab, n_mh, fast_ok = device_table(key, lambda: H.three_level_tables(cfg.n_fine, cfg.n_mid, cfg.n_high,
                                                                            cfg.fine_to_mid, cfg.fine_to_high), dev)
        if not FAST_PATH["enabled"]:
            fast_ok = 0
        st = None
        if embedding is not None and cfg.use_triplet:
            tkey = ("t1", cfg.upper_ids, cfg.lower_ids)
            ttab, ncls = device_table(tkey, lambda: H.triplet_tables_id_lists(cfg.upper_ids, cfg.lower_ids), dev)
            emb = embedding.contiguous()
            fork = _Fork(dev)
            st = triplet_forward(emb, lab, 1, ttab, ncls, fork=fork, world_ready=True)
        with torch.cuda.device(dev):
            lib = _lib.load()
            nbytes = lib.sh_rmi3_workspace_bytes(b, hh, ww, cfg.n_fine, cfg.n_mid, cfg.n_high)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            out = torch.empty(4, dtype=torch.float32, device=dev)
            _staged("sh_rmi3_forward", (1, 2, 4, 8), lambda st_bits: (
                _p(x), _dtype_code(x), _p(lab), b, hh, ww, cfg.n_fine, cfg.n_mid, cfg.n_high, _p(tab), n_mh, fast_ok,
                cfg.lam, cfg.loss_weight, _p(ws), st_bits, _stream()))
            if st is not None:
                fork.join()
            _call("sh_loss3_final", b, hh, ww, cfg.n_fine, cfg.n_mid, cfg.n_high, _p(ws), cfg.lam, _p(step_d),
                      cfg.total_steps, _p(st.trip) if st else None, _p(st.status) if st else None, cfg.loss_weight,
                      _p(out), _stream())
        stats.update(out=out, triplet=st, workspace=ws,
                     fast_path=bool(lib.sh_rmi3_fast_path(_p(x), None, _dtype_code(x), hh, ww, cfg.n_fine, cfg.n_mid,
                                                          cfg.n_high, fast_ok)))
        ctx.cfg = cfg
        ctx.tab, ctx.n_mh, ctx.fast_ok = tab, n_mh, fast_ok
        ctx.st = st
        ctx.out = out
        ctx.ws = ws if ctx.needs_input_grad[0] else None
        ctx.x = x if ctx.needs_input_grad[0] else None
        ctx.emb = emb if st is not None else None
        return out[0].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        cfg = ctx.cfg
        gx = None
        gemb = fork = None
        if ctx.st is not None and ctx.needs_input_grad[1]:
            fork = _Fork(ctx.emb.device)
            gemb = triplet_backward(ctx.emb, ctx.st, ctx.out[1:2], g, fork=fork)
        if ctx.needs_input_grad[0]:
            x = ctx.x
            b, c, hh, ww = x.shape
            gx = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _staged("sh_rmi3_backward", (1, 2), lambda st_bits: (
                    _p(x), _dtype_code(x), _p(gx), b, hh, ww, cfg.n_fine, cfg.n_mid, cfg.n_high, _p(ctx.tab), ctx.n_mh,
                    ctx.fast_ok, cfg.loss_weight, _p(ctx.ws), _p(g), st_bits, _stream()))
        if fork is not None:
            fork.join()
            gemb = _grad_as(gemb, ctx.emb)
        return gx, gemb, None, None, None, None

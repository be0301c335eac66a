"""Host side of the C-ABI custom ops: tensor plumbing, workspaces and autograd glue.

Every computation is a kernel of libseghiero_b200.so reached through ctypes (`_lib`); PyTorch is used for device
memory, streams and autograd bookkeeping only.  The fused losses are registered torch custom ops
(`torch.library.custom_op` + `register_autograd` + fake implementations, namespace `seghiero_b200`), so they are
opaque, traceable nodes for `torch.compile(fullgraph=True)` and save their tensors through `save_for_backward`
(in-place edits of the logits between forward and backward are detected by autograd's version counters).
Nothing here touches the host inside forward/backward (no .item(), no .tolist()): `step`, the `ready` gate and
`grad_output` stay on the device.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from . import hierarchy as H

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_LAB = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}
_table_cache: dict = {}

# kernels launched per C-ABI call (per stage bit for the multi-kernel entry points); bench.py reads
# LAUNCHES["n"] to report how many of OUR kernels ran inside its timed region
_KERNELS = {
    "sh_targets_two_level": 1, "sh_targets_three_level": 1, "sh_targets_gather": 1, "sh_decode": 1, "sh_colorize": 1,
    "sh_loss2_final": 1, "sh_loss3_final": 1, "sh_scale_inplace": 1, "sh_triplet_forward": 4,
    "sh_triplet_backward": 1, "sh_upsample_bilinear": 1, "sh_upsample_bilinear_adjoint": 1, "sh_aux_ce_fwdbwd": 2,
    "sh_decode_upsampled": 1,
    ("sh_bce2_fwdbwd", 1): 1, ("sh_bce2_fwdbwd", 2): 1, ("sh_bce2_fwdbwd", 4): 1,
    ("sh_rmi3_forward", 1): 1, ("sh_rmi3_forward", 2): 1, ("sh_rmi3_forward", 4): 2, ("sh_rmi3_forward", 8): 2,
    ("sh_rmi3_backward", 1): 2, ("sh_rmi3_backward", 2): 1,
}
LAUNCHES = {"n": 0}


class _Debug:
    """Measurement hook, off in the product path: bench.py brackets single kernels of the multi-kernel entry points
    with CUDA events through `stage_timing(timer)` (one C call per stage bit instead of one per entry point)."""
    timer = None


@contextlib.contextmanager
def stage_timing(timer):
    """`timer` has start(name, bit) / stop(name, bit); active inside the with-block only."""
    prev, _Debug.timer = _Debug.timer, timer
    try:
        yield timer
    finally:
        _Debug.timer = prev


def _call(name, *args):
    _lib.call(name, *args)
    LAUNCHES["n"] += _KERNELS.get(name, 0)


def _timed_call(name, *args):
    """Single-kernel entry point that bench.py may want to see as a stage of its own."""
    timer = _Debug.timer
    if timer is not None:
        timer.start(name, 0)
    _call(name, *args)
    if timer is not None:
        timer.stop(name, 0)


def _staged(name, bits, make_args):
    """Run a multi-kernel entry point.  Normally one C call with every stage bit set; with a stage
    timer installed, one call per stage so that each kernel can be bracketed by CUDA events."""
    timer = _Debug.timer
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


def _labels(label: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Labels as the kernels read them: int64 (the reference's dtype), int32 or uint8 pass through untouched
    (1 byte per pixel from the dataloader to the loss); any other integer type is widened."""
    if label.dtype not in _LAB:
        label = label.long()
    label = label.contiguous()
    return label, _LAB[label.dtype]


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
# targets / decode / colourise (integer outputs, no autograd)
# ----------------------------------------------------------------------------------------------
def build_fine_to_level_map(map_cfg, n_fine: int) -> torch.Tensor:
    """Drop-in for dataset/dataloader.py:12-34: a list of [lbl] / [start, end] (inclusive) entries, one per level
    class -> LongTensor [n_fine] mapping each fine id to its level id.  Same assertions and ValueErrors."""
    mapping = [-1] * n_fine
    for lvl, sub in enumerate(map_cfg):
        if len(sub) == 1:
            lbl = int(sub[0])
            assert 0 <= lbl < n_fine, f"Label {lbl} outside [0..{n_fine-1}]"
            mapping[lbl] = lvl
        elif len(sub) == 2:
            start, end = int(sub[0]), int(sub[1])
            assert 0 <= start <= end < n_fine, f"Range [{start},{end}] invalid"
            mapping[start:end + 1] = [lvl] * (end + 1 - start)
        else:
            raise ValueError(f"Each entry must be [lbl] or [start,end], got {sub}")
    missing = [i for i, m in enumerate(mapping) if m < 0]
    if missing:
        raise ValueError(f"Fine‐labels not mapped: {missing}")
    return torch.tensor(mapping, dtype=torch.long)


def targets_two_level(label: torch.Tensor, hiera_index):
    """hiera_triplet_loss.py:11-38.  The output has the dtype of `label` (int64 / int32 / uint8 without a copy)."""
    _need_cuda(label)
    lab, lcode = _labels(label)
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
        _call("sh_targets_two_level", _p(lab), lcode, _p(out), lab.numel(), _p(lut), lut.numel(), _stream())
    return out.to(label.dtype) if label.dtype != lab.dtype else out


def targets_three_level(label: torch.Tensor, fine_to_mid: torch.Tensor, fine_to_high: torch.Tensor, check=True):
    """rmi_hiera_triplet_loss.py:21-63 (255 passes through; out-of-range labels raise IndexError when `check`)."""
    _need_cuda(label)
    lab, lcode = _labels(label)
    f2m = fine_to_mid.to(device=lab.device, dtype=torch.int64).contiguous()
    f2h = fine_to_high.to(device=lab.device, dtype=torch.int64).contiguous()
    mid, high = torch.empty_like(lab), torch.empty_like(lab)
    err = torch.zeros(1, dtype=torch.int32, device=lab.device)
    with torch.cuda.device(lab.device):
        _call("sh_targets_three_level", _p(lab), lcode, _p(mid), _p(high), lab.numel(), _p(f2m), _p(f2h), f2m.numel(),
              _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("fine label out of range for fine_to_mid / fine_to_high")
    return mid, high


def targets_gather(fine_mask: torch.Tensor, level_map: torch.Tensor, check=True):
    """`level_map[fine_mask]` of dataset/dataloader.py:166-177 on the GPU (no ignore handling: 255 raises IndexError
    like the reference's gather unless the map has 256 entries)."""
    _need_cuda(fine_mask)
    lab, lcode = _labels(fine_mask)
    m = level_map.to(device=lab.device, dtype=torch.int64).contiguous()
    out = torch.empty_like(lab)
    err = torch.zeros(1, dtype=torch.int32, device=lab.device)
    with torch.cuda.device(lab.device):
        _call("sh_targets_gather", _p(lab), lcode, _p(out), lab.numel(), _p(m), m.numel(), _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("index out of range in target gather")
    return out


def colorize(mask: torch.Tensor, colormap, check=True) -> torch.Tensor:
    """GPU form of infer.py:117-131 (`mask_to_color_image`, a per-pixel Python loop in the reference): class-id mask
    [..., H, W] (any integer dtype) + colormap (list of (r, g, b)) -> uint8 [..., H, W, 3].  Negative ids are black,
    ids beyond the colormap raise IndexError."""
    _need_cuda(mask)
    m, lcode = _labels(mask)
    pal = torch.as_tensor(np.asarray(colormap, dtype=np.uint8).reshape(-1, 3)).to(m.device)
    out = torch.empty(tuple(m.shape) + (3,), dtype=torch.uint8, device=m.device)
    err = torch.zeros(1, dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        _call("sh_colorize", _p(m), lcode, m.numel(), _p(pal), pal.shape[0], _p(out), _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("list index out of range (class id beyond the colormap)")
    return out


def hierarchical_argmax(logits: torch.Tensor, level_sizes: Sequence[int], label: Optional[torch.Tensor] = None,
                        out_dtype=torch.int64, size: Optional[Sequence[int]] = None):
    """Per-level argmax over channel slices (+ fine pixel-accuracy counts when `label` is given).
    Returns (list of [B,H,W] predictions, counts int64[2] = (#correct, #valid) or None).

    `size=(H, W)` (or a `label` larger than the logits) decodes F.interpolate(logits, size, "bilinear",
    align_corners=False) as train.py:345-352 / infer.py:296-312 do -- fused for the head's 4x geometry (the
    full-resolution logits are never written), through our upsample kernel otherwise."""
    _need_cuda(logits)
    if out_dtype not in (torch.int64, torch.uint8):
        raise TypeError("out_dtype must be torch.int64 or torch.uint8")
    x = logits.contiguous()
    if x.dim() != 4:
        raise ValueError("logits must be [B,C,H,W]")
    b, c, h, w = x.shape
    if size is None and label is not None and tuple(label.shape[-2:]) != (h, w):
        size = tuple(label.shape[-2:])
    hh, ww = (int(size[0]), int(size[1])) if size is not None else (h, w)
    sizes = [int(s) for s in level_sizes] + [0, 0]
    n0, n1, n2 = sizes[:3]
    if len(level_sizes) > 3 or n0 + n1 + n2 > c:
        raise ValueError("level_sizes must be <= 3 levels and fit the channel count")
    outs = [torch.empty((b, hh, ww), dtype=out_dtype, device=x.device) if n > 0 else None for n in (n0, n1, n2)]
    counts = None
    lab, lcode = None, 0
    if label is not None:
        lab, lcode = _labels(label)
        counts = torch.zeros(2, dtype=torch.int64, device=x.device)
    u8 = 1 if out_dtype == torch.uint8 else 0
    with torch.cuda.device(x.device):
        if (hh, ww) != (h, w):
            lib = _lib.load()
            rc = lib.sh_decode_upsampled(_p(x), _dtype_code(x), b, c, h, w, hh, ww, n0, n1, n2, _p(outs[0]), _p(outs[1]),
                                         _p(outs[2]), u8, _p(lab), lcode, _p(counts), _stream())
            if rc == 0:
                LAUNCHES["n"] += 1
                return [o for o in outs if o is not None], counts
            if rc != -2:
                _lib.check(rc, "sh_decode_upsampled")
            full = torch.empty((b, c, hh, ww), dtype=x.dtype, device=x.device)
            _call("sh_upsample_bilinear", _p(x), _dtype_code(x), _p(full), b * c, h, w, hh, ww, _stream())
            x = full
        _call("sh_decode", _p(x), _dtype_code(x), b, c, hh * ww, n0, n1, n2, _p(outs[0]), _p(outs[1]), _p(outs[2]),
              u8, _p(lab), lcode, _p(counts), _stream())
    return [o for o in outs if o is not None], counts


# ----------------------------------------------------------------------------------------------
# shared pieces of the fused ops
# ----------------------------------------------------------------------------------------------
def _empty(dev, dtype=torch.float32):
    return torch.empty(0, dtype=dtype, device=dev)


def _opt(t: Tensor) -> Optional[Tensor]:
    return t if t.numel() else None


def _upsampled(x: Tensor, hh: int, ww: int) -> Tensor:
    """The head's logits at label resolution (train.py:282-284) when the caller passed them at their own."""
    b, c, h, w = x.shape
    if (h, w) == (hh, ww):
        return x
    full = torch.empty((b, c, hh, ww), dtype=x.dtype, device=x.device)
    _timed_call("sh_upsample_bilinear", _p(x), _dtype_code(x), _p(full), b * c, h, w, hh, ww, _stream())
    return full


def _triplet_alloc(dev, rows: int, ncls: int, max_triplet: int):
    return (torch.empty(rows, dtype=torch.int32, device=dev),                      # lab_ds
            torch.empty(ncls * 3 * max_triplet, dtype=torch.int32, device=dev),    # sel
            torch.empty(ncls, dtype=torch.int32, device=dev),                      # kcount
            torch.empty(ncls * max_triplet, dtype=torch.float32, device=dev),      # tl
            torch.empty(2, dtype=torch.float32, device=dev),                       # trip  [loss, #classes]
            torch.empty(2, dtype=torch.int32, device=dev))                         # status [ready, error]


def _triplet_forward(feats: Tensor, lab: Tensor, lcode: int, mode: int, tab: Tensor, ncls: int, max_triplet: int,
                     fork: Optional[_Fork], world_ready: bool):
    b, d, h, w = feats.shape
    hh, ww = lab.shape[-2:]
    lab_ds, sel, kcount, tl, trip, status = _triplet_alloc(feats.device, b * h * w, ncls, max_triplet)
    with (fork if fork is not None else contextlib.nullcontext()):
        _call("sh_triplet_forward", _p(feats), _dtype_code(feats), _p(lab), lcode, b, d, h, w, hh, ww, mode, _p(tab),
              ncls, max_triplet, _p(lab_ds), _p(sel), _p(kcount), _p(tl), _p(trip), _p(status), _stream())
        if world_ready:
            _world_ready(status)      # on the side stream too: the all-reduce overlaps the loss kernels
    return sel, kcount, tl, trip, status


def _triplet_backward(feats: Tensor, ncls: int, max_triplet: int, sel, kcount, tl, trip, tscale: Tensor,
                      gscale: Tensor, fork: Optional[_Fork]) -> Tensor:
    """fp32 gradient of the embedding (the caller joins the fork and converts to the embedding's dtype)."""
    b, d, h, w = feats.shape
    gfeat = torch.empty((b, d, h, w), dtype=torch.float32, device=feats.device)
    with (fork if fork is not None else contextlib.nullcontext()):
        scratch = torch.empty((b, d, h, w), dtype=torch.int64, device=feats.device)   # fixed-point accumulator
        _call("sh_triplet_backward", _p(feats), _dtype_code(feats), b, d, h, w, ncls, max_triplet, _p(sel), _p(kcount),
              _p(tl), _p(trip), _p(tscale), _p(gscale), _p(gfeat), _p(scratch), _stream())
    return gfeat


def _grad_as(gfeat: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
    return gfeat if feats.dtype == torch.float32 else gfeat.to(feats.dtype)


def _gscale(g: Tensor) -> Tensor:
    """grad_output of the loss vector -> the device scalar of its first entry (the loss)."""
    return g.detach().reshape(-1)[:1].to(torch.float32).contiguous()


# ----------------------------------------------------------------------------------------------
# standalone triplet loss
# ----------------------------------------------------------------------------------------------
def _triplet_tab(mode: int, a: Sequence[int], bflat: Sequence[int], dev):
    if mode == 0:
        hi = tuple((int(bflat[2 * i]), int(bflat[2 * i + 1])) for i in range(len(bflat) // 2))
        key = ("t0", tuple(int(v) for v in a), hi)
        return device_table(key, lambda: H.triplet_tables_hierarchy(list(a), [list(r) for r in hi]), dev)
    key = ("t1", tuple(int(v) for v in a), tuple(int(v) for v in bflat))
    return device_table(key, lambda: H.triplet_tables_id_lists(list(a), list(bflat)), dev)


@torch.library.custom_op("seghiero_b200::triplet_fwd", mutates_args=())
def triplet_fwd(feats: Tensor, label: Tensor, mode: int, list_a: List[int], list_b: List[int],
                max_triplet: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (trip [loss, #classes], status [ready, error], sel, kcount, tl).  mode 0: list_a = hiera_map, list_b =
    flattened hiera_index (tree_triplet_loss.py:15-65); mode 1: upper_ids / lower_ids (rmi_tree_triplet_loss.py:14-70)."""
    _need_cuda(feats, label)
    with torch.cuda.device(feats.device):
        tab, ncls = _triplet_tab(mode, list_a, list_b, feats.device)
        lab, lcode = _labels(label)
        sel, kcount, tl, trip, status = _triplet_forward(feats.contiguous(), lab, lcode, mode, tab, ncls, max_triplet,
                                                         None, False)
    return trip, status, sel, kcount, tl


@triplet_fwd.register_fake
def _(feats, label, mode, list_a, list_b, max_triplet):
    ncls = len(list_a) if mode == 0 else 256
    dev = feats.device
    return (torch.empty(2, dtype=torch.float32, device=dev), torch.empty(2, dtype=torch.int32, device=dev),
            torch.empty(ncls * 3 * max_triplet, dtype=torch.int32, device=dev),
            torch.empty(ncls, dtype=torch.int32, device=dev),
            torch.empty(ncls * max_triplet, dtype=torch.float32, device=dev))


@torch.library.custom_op("seghiero_b200::triplet_bwd", mutates_args=())
def triplet_bwd(gout: Tensor, feats: Tensor, sel: Tensor, kcount: Tensor, tl: Tensor, trip: Tensor, tscale: Tensor,
                max_triplet: int) -> Tensor:
    with torch.cuda.device(feats.device):
        f = feats.contiguous()
        g = _triplet_backward(f, kcount.numel(), max_triplet, sel, kcount, tl, trip, tscale, _gscale(gout), None)
    return _grad_as(g, feats)


@triplet_bwd.register_fake
def _(gout, feats, sel, kcount, tl, trip, tscale, max_triplet):
    return torch.empty_like(feats, memory_format=torch.contiguous_format)


def _triplet_setup(ctx, inputs, output):
    feats, _label, _mode, _a, _b, max_triplet = inputs
    trip, status, sel, kcount, tl = output
    ctx.save_for_backward(feats, sel, kcount, tl, trip)
    ctx.set_materialize_grads(False)      # no zero-filled "gradients" for the bookkeeping outputs
    ctx.max_triplet = max_triplet
    ctx.mark_non_differentiable(status, sel, kcount, tl)


def _triplet_backward_fn(ctx, g_trip, *_unused):
    feats, sel, kcount, tl, trip = ctx.saved_tensors
    if g_trip is None:
        return (None,) * 6
    one = torch.ones(1, dtype=torch.float32, device=feats.device)
    return triplet_bwd(g_trip, feats, sel, kcount, tl, trip, one, ctx.max_triplet), None, None, None, None, None


triplet_fwd.register_autograd(_triplet_backward_fn, setup_context=_triplet_setup)


# ----------------------------------------------------------------------------------------------
# two-level fused loss: HieraTripletLoss.forward, models/loss/hiera_triplet_loss.py:152-211
# ----------------------------------------------------------------------------------------------
def _hier2_tables(n_fine: int, index_flat: Sequence[int], dev):
    hi = tuple((int(index_flat[2 * i]), int(index_flat[2 * i + 1])) for i in range(len(index_flat) // 2))
    key = ("h2", n_fine, hi)
    tab, n_fb, lut_size = device_table(key, lambda: H.two_level_tables(n_fine, [list(r) for r in hi]), dev)
    return tab, n_fb, lut_size, H.two_level_is_tree(n_fine, [list(r) for r in hi]), hi


def two_level_supported(n_fine: int, n_coarse: int, fast: bool) -> bool:
    """Shared-memory limits of the 2-level kernels (csrc/bce2.cu): the tree-order kernel parks e^-x of <= 64 channels,
    the any-bucket kernel parks sigmoid and e^x of every channel for 256 pixels, or 128 for wider hierarchies
    (up to ~220 channels, e.g. 150 fine + 30 coarse)."""
    c = n_fine + n_coarse
    return (fast and c <= 64) or (c * 128 * 8 + n_coarse * 128 <= 227 * 1024)


@torch.library.custom_op("seghiero_b200::hier2_fwd", mutates_args=())
def hier2_fwd(cls_score: Tensor, label: Tensor, embedding: Optional[Tensor], step: Tensor, n_fine: int,
              hiera_map: List[int], hiera_index: List[int], loss_weight: float, total_steps: float, fast_path: bool,
              want_grad: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (out [loss, triplet scale, -, -], grad (for grad_output = 1; empty unless want_grad), counts [#valid fine,
    #valid coarse, label error, -], sel, kcount, tl, trip, status).  `cls_score` may be at the label's resolution (the
    reference's contract) or at the head's: it is then upsampled here and `grad` comes back at the head's resolution."""
    _need_cuda(cls_score, label, embedding)
    dev = cls_score.device
    nc = len(hiera_index) // 2
    x_in = cls_score.contiguous()
    lab, lcode = _labels(label)
    if x_in.dim() != 4 or lab.dim() != 3:
        raise ValueError("cls_score must be [B,C,H,W] and label [B,H,W]")
    b, c, h_in, w_in = x_in.shape
    hh, ww = int(lab.shape[1]), int(lab.shape[2])
    if c != n_fine + nc:
        raise ValueError(f"cls_score has {c} channels, expected n_fine+n_coarse={n_fine + nc}")
    if lab.shape[0] != b:
        raise ValueError("label must be [B,H,W] with the batch size of cls_score")
    with torch.cuda.device(dev):
        tab, n_fb, lut_size, is_tree, hi = _hier2_tables(n_fine, hiera_index, dev)
        tree = 256 if (fast_path and is_tree) else 0
        if not two_level_supported(n_fine, nc, bool(tree)):
            raise ValueError(f"HieraTripletLoss on sm_100a: {c} channels exceed the shared-memory tiling of the 2-level "
                             "kernels (<= ~220 channels)")
        fork = None
        if embedding is not None:
            ttab, ncls = _triplet_tab(0, hiera_map, hiera_index, dev)
            fork = _Fork(dev)
            sel, kcount, tl, trip, status = _triplet_forward(embedding.contiguous(), lab, lcode, 0, ttab, ncls, 200,
                                                             fork, True)
        else:
            sel, kcount, tl, trip, status = (_empty(dev, torch.int32), _empty(dev, torch.int32), _empty(dev),
                                             _empty(dev), _empty(dev, torch.int32))
        x = _upsampled(x_in, hh, ww)
        hw = hh * ww
        grad_full = torch.empty_like(x) if want_grad else None
        lab8 = torch.empty(b * hw, dtype=torch.uint8, device=dev)
        counts = torch.empty(4, dtype=torch.int64, device=dev)
        grid = _lib.load().sh_bce2_grid(b, hw, c, nc)
        partials = torch.empty(grid * 4, dtype=torch.float32, device=dev)
        sums = torch.empty(4, dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        _staged("sh_bce2_fwdbwd", (1, 2, 4), lambda st_bits: (
            _p(x), _dtype_code(x), _p(lab), lcode, _p(grad_full), b, hw, n_fine, nc, _p(tab), n_fb, lut_size,
            1e-8, loss_weight, _p(lab8), _p(counts), _p(partials), _p(sums), st_bits | tree, _stream()))
        if fork is not None:
            fork.join()
        _call("sh_loss2_final", _p(sums), _p(counts), n_fine, nc, float(b * hw), _p(step), total_steps,
              _p(_opt(trip)), _p(_opt(status)), loss_weight, _p(out), _stream())
        grad = _empty(dev, x.dtype)
        if want_grad:
            grad = grad_full
            if x is not x_in:        # back to the head's resolution through the adjoint of the interpolation
                grad = torch.empty_like(x_in)
                _timed_call("sh_upsample_bilinear_adjoint", _p(grad_full), _dtype_code(x), _p(grad), b * c, h_in, w_in,
                            hh, ww, _stream())
    return out, grad, counts, sel, kcount, tl, trip, status


@hier2_fwd.register_fake
def _(cls_score, label, embedding, step, n_fine, hiera_map, hiera_index, loss_weight, total_steps, fast_path, want_grad):
    dev = cls_score.device
    ncls, mt = len(hiera_map), 200
    has = embedding is not None
    i32, f32 = torch.int32, torch.float32
    return (torch.empty(4, dtype=f32, device=dev),
            torch.empty_like(cls_score, memory_format=torch.contiguous_format) if want_grad
            else torch.empty(0, dtype=cls_score.dtype, device=dev),
            torch.empty(4, dtype=torch.int64, device=dev),
            torch.empty(ncls * 3 * mt if has else 0, dtype=i32, device=dev),
            torch.empty(ncls if has else 0, dtype=i32, device=dev),
            torch.empty(ncls * mt if has else 0, dtype=f32, device=dev),
            torch.empty(2 if has else 0, dtype=f32, device=dev),
            torch.empty(2 if has else 0, dtype=i32, device=dev))


@torch.library.custom_op("seghiero_b200::hier2_bwd", mutates_args=("grad",))
def hier2_bwd(gout: Tensor, grad: Tensor, embedding: Optional[Tensor], sel: Tensor, kcount: Tensor, tl: Tensor,
              trip: Tensor, out: Tensor, need_x: bool, need_emb: bool) -> Tensor:
    """Scales the precomputed logits gradient in place by grad_output (a no-op kernel exit when it is 1) and returns
    the embedding gradient (empty when there is none)."""
    dev = gout.device
    g = _gscale(gout)
    gemb = _empty(dev)
    with torch.cuda.device(dev):
        fork = None
        if need_emb and embedding is not None:
            fork = _Fork(dev)
            emb = embedding.contiguous()
            gemb = _triplet_backward(emb, kcount.numel(), 200, sel, kcount, tl, trip, out[1:2], g, fork)
        if need_x and grad.numel():
            _call("sh_scale_inplace", _p(grad), _dtype_code(grad), grad.numel(), _p(g), _stream())
        if fork is not None:
            fork.join()
            gemb = _grad_as(gemb, embedding)
    return gemb


@hier2_bwd.register_fake
def _(gout, grad, embedding, sel, kcount, tl, trip, out, need_x, need_emb):
    if need_emb and embedding is not None:
        return torch.empty_like(embedding, memory_format=torch.contiguous_format)
    return torch.empty(0, dtype=torch.float32, device=gout.device)


def _hier2_setup(ctx, inputs, output):
    embedding = inputs[2]
    out, grad, counts, sel, kcount, tl, trip, status = output
    ctx.save_for_backward(grad, embedding, sel, kcount, tl, trip, out)
    ctx.set_materialize_grads(False)      # no zero-filled "gradients" for the bookkeeping outputs
    ctx.mark_non_differentiable(grad, counts, sel, kcount, tl, trip, status)
    ctx.consumed = False


def _hier2_backward_fn(ctx, g_out, *_unused):
    grad, embedding, sel, kcount, tl, trip, out = ctx.saved_tensors
    if g_out is None:
        return (None,) * 11
    need_x, need_emb = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
    if need_x and ctx.consumed:
        raise RuntimeError("seghiero_b200: backward through the fused 2-level loss can run only once "
                           "(its gradient is computed in the forward kernel and scaled in place)")
    ctx.consumed = True
    gemb = hier2_bwd(g_out, grad, embedding, sel, kcount, tl, trip, out, need_x, need_emb)
    gx = grad if need_x else None
    ge = gemb if (need_emb and embedding is not None) else None
    return (gx, None, ge) + (None,) * 8


hier2_fwd.register_autograd(_hier2_backward_fn, setup_context=_hier2_setup)


# ----------------------------------------------------------------------------------------------
# three-level fused loss: RMIHieraTripletLoss.forward, models/loss/rmi_hiera_triplet_loss.py:323-546
# ----------------------------------------------------------------------------------------------
def _hier3_tables(nf: int, nm: int, nh: int, f2m: Sequence[int], f2h: Sequence[int], dev):
    key = ("h3", nf, nm, nh, tuple(int(v) for v in f2m), tuple(int(v) for v in f2h))
    return device_table(key, lambda: H.three_level_tables(nf, nm, nh, list(f2m), list(f2h)), dev)


@torch.library.custom_op("seghiero_b200::hier3_fwd", mutates_args=())
def hier3_fwd(cls_score: Tensor, label: Tensor, embedding: Optional[Tensor], step: Tensor, n_fine: int, n_mid: int,
              n_high: int, fine_to_mid: List[int], fine_to_high: List[int], upper_ids: List[int], lower_ids: List[int],
              lam: float, loss_weight: float, total_steps: float, fast_path: bool, use_triplet: bool
              ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (out [loss, triplet scale, rmi term, -], workspace, x_full (the upsampled logits when `cls_score` came at the
    head's resolution, else empty), sel, kcount, tl, trip, status)."""
    _need_cuda(cls_score, label, embedding)
    dev = cls_score.device
    x_in = cls_score.contiguous()
    lab, lcode = _labels(label)
    if x_in.dim() != 4 or lab.dim() != 3:
        raise ValueError("cls_score must be [B,C,H,W] and label [B,H,W]")
    b, c, _, _ = x_in.shape
    hh, ww = int(lab.shape[1]), int(lab.shape[2])
    if c != n_fine + n_mid + n_high:
        raise ValueError(f"cls_score has {c} channels, expected {n_fine + n_mid + n_high}")
    if lab.shape[0] != b:
        raise ValueError("label must be [B,H,W] with the batch size of cls_score")
    if hh < 8 or ww < 8:
        raise ValueError("the CUDA RMI path needs H, W >= 8 (the reference needs >= 3)")
    with torch.cuda.device(dev):
        tab, n_mh, fast_ok = _hier3_tables(n_fine, n_mid, n_high, fine_to_mid, fine_to_high, dev)
        if not fast_path:
            fast_ok = 0
        fork = None
        if embedding is not None and use_triplet:
            ttab, ncls = _triplet_tab(1, upper_ids, lower_ids, dev)
            fork = _Fork(dev)
            sel, kcount, tl, trip, status = _triplet_forward(embedding.contiguous(), lab, lcode, 1, ttab, ncls, 200,
                                                             fork, True)
        else:
            sel, kcount, tl, trip, status = (_empty(dev, torch.int32), _empty(dev, torch.int32), _empty(dev),
                                             _empty(dev), _empty(dev, torch.int32))
        x = _upsampled(x_in, hh, ww)
        lib = _lib.load()
        ws = torch.empty(lib.sh_rmi3_workspace_bytes(b, hh, ww, n_fine, n_mid, n_high), dtype=torch.uint8, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        _staged("sh_rmi3_forward", (1, 2, 4, 8), lambda st_bits: (
            _p(x), _dtype_code(x), _p(lab), lcode, b, hh, ww, n_fine, n_mid, n_high, _p(tab), n_mh, fast_ok,
            lam, loss_weight, _p(ws), st_bits, _stream()))
        if fork is not None:
            fork.join()
        _call("sh_loss3_final", b, hh, ww, n_fine, n_mid, n_high, _p(ws), lam, _p(step), total_steps,
              _p(_opt(trip)), _p(_opt(status)), loss_weight, _p(out), _stream())
    x_full = x if x is not x_in else _empty(dev, x.dtype)
    return out, ws, x_full, sel, kcount, tl, trip, status


@hier3_fwd.register_fake
def _(cls_score, label, embedding, step, n_fine, n_mid, n_high, fine_to_mid, fine_to_high, upper_ids, lower_ids, lam,
      loss_weight, total_steps, fast_path, use_triplet):
    dev = cls_score.device
    b, c = cls_score.shape[0], cls_score.shape[1]
    hh, ww = label.shape[1], label.shape[2]
    has = embedding is not None and use_triplet
    ncls, mt = 256, 200
    i32, f32 = torch.int32, torch.float32
    nbytes = _lib.load().sh_rmi3_workspace_bytes(int(b), int(hh), int(ww), n_fine, n_mid, n_high)
    up = (cls_score.shape[2], cls_score.shape[3]) != (hh, ww)
    return (torch.empty(4, dtype=f32, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev),
            torch.empty((b, c, hh, ww) if up else (0,), dtype=cls_score.dtype, device=dev),
            torch.empty(ncls * 3 * mt if has else 0, dtype=i32, device=dev),
            torch.empty(ncls if has else 0, dtype=i32, device=dev),
            torch.empty(ncls * mt if has else 0, dtype=f32, device=dev),
            torch.empty(2 if has else 0, dtype=f32, device=dev),
            torch.empty(2 if has else 0, dtype=i32, device=dev))


@torch.library.custom_op("seghiero_b200::hier3_bwd", mutates_args=())
def hier3_bwd(gout: Tensor, cls_score: Tensor, x_full: Tensor, ws: Tensor, embedding: Optional[Tensor], sel: Tensor,
              kcount: Tensor, tl: Tensor, trip: Tensor, out: Tensor, n_fine: int, n_mid: int, n_high: int,
              fine_to_mid: List[int], fine_to_high: List[int], loss_weight: float, fast_path: bool, need_x: bool,
              need_emb: bool) -> Tuple[Tensor, Tensor]:
    """Pass 2 (+ the image frame) -> d loss / d cls_score at the resolution `cls_score` came in, and the embedding
    gradient."""
    dev = cls_score.device
    g = _gscale(gout)
    gx, gemb = _empty(dev, cls_score.dtype), _empty(dev)
    with torch.cuda.device(dev):
        fork = None
        if need_emb and embedding is not None and sel.numel():
            fork = _Fork(dev)
            emb = embedding.contiguous()
            gemb = _triplet_backward(emb, kcount.numel(), 200, sel, kcount, tl, trip, out[1:2], g, fork)
        if need_x:
            x_in = cls_score.contiguous()
            x = x_full if x_full.numel() else x_in
            b, c, hh, ww = x.shape
            tab, n_mh, fast_ok = _hier3_tables(n_fine, n_mid, n_high, fine_to_mid, fine_to_high, dev)
            if not fast_path:
                fast_ok = 0
            gfull = torch.empty_like(x)
            _staged("sh_rmi3_backward", (1, 2), lambda st_bits: (
                _p(x), _dtype_code(x), _p(gfull), b, hh, ww, n_fine, n_mid, n_high, _p(tab), n_mh, fast_ok,
                loss_weight, _p(ws), _p(g), st_bits, _stream()))
            gx = gfull
            if x is not x_in:
                gx = torch.empty_like(x_in)
                _timed_call("sh_upsample_bilinear_adjoint", _p(gfull), _dtype_code(x), _p(gx), b * c, x_in.shape[2],
                            x_in.shape[3], hh, ww, _stream())
        if fork is not None:
            fork.join()
            gemb = _grad_as(gemb, embedding)
    return gx, gemb


@hier3_bwd.register_fake
def _(gout, cls_score, x_full, ws, embedding, sel, kcount, tl, trip, out, n_fine, n_mid, n_high, fine_to_mid,
      fine_to_high, loss_weight, fast_path, need_x, need_emb):
    dev = cls_score.device
    gx = (torch.empty_like(cls_score, memory_format=torch.contiguous_format) if need_x
          else torch.empty(0, dtype=cls_score.dtype, device=dev))
    if need_emb and embedding is not None and sel.numel():
        ge = torch.empty_like(embedding, memory_format=torch.contiguous_format)
    else:
        ge = torch.empty(0, dtype=torch.float32, device=dev)
    return gx, ge


def _hier3_setup(ctx, inputs, output):
    (cls_score, _label, embedding, _step, n_fine, n_mid, n_high, f2m, f2h, _up, _lo, _lam, loss_weight, _ts, fast_path,
     _ut) = inputs
    out, ws, x_full, sel, kcount, tl, trip, status = output
    ctx.save_for_backward(cls_score, x_full, ws, embedding, sel, kcount, tl, trip, out)
    ctx.set_materialize_grads(False)      # else autograd zero-fills a "gradient" for the 0.5 GB workspace every step
    ctx.mark_non_differentiable(ws, x_full, sel, kcount, tl, trip, status)
    ctx.cfg = (n_fine, n_mid, n_high, list(f2m), list(f2h), float(loss_weight), bool(fast_path))


def _hier3_backward_fn(ctx, g_out, *_unused):
    cls_score, x_full, ws, embedding, sel, kcount, tl, trip, out = ctx.saved_tensors
    if g_out is None:
        return (None,) * 16
    need_x, need_emb = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
    nf, nm, nh, f2m, f2h, lw, fast = ctx.cfg
    gx, gemb = hier3_bwd(g_out, cls_score, x_full, ws, embedding, sel, kcount, tl, trip, out, nf, nm, nh, f2m, f2h, lw,
                         fast, need_x, need_emb)
    has_emb = need_emb and embedding is not None and sel.numel() > 0
    return (gx if need_x else None, None, gemb if has_emb else None) + (None,) * 13


hier3_fwd.register_autograd(_hier3_backward_fn, setup_context=_hier3_setup)


# ----------------------------------------------------------------------------------------------
# aux-head cross entropy fused with its upsample (train.py:309-313), SURVEY 8f N2
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("seghiero_b200::aux_ce_fwd", mutates_args=())
def aux_ce_fwd(logits: Tensor, label: Tensor, want_grad: bool) -> Tuple[Tensor, Tensor]:
    """-> (loss [1], gradient w.r.t. the low-resolution logits for grad_output = 1 (empty unless want_grad))."""
    _need_cuda(logits, label)
    x = logits.contiguous()
    lab, lcode = _labels(label)
    if x.dim() != 4 or lab.dim() != 3 or lab.shape[0] != x.shape[0]:
        raise ValueError("logits must be [B,C,h,w] and label [B,H,W]")
    b, c, h, w = x.shape
    hh, ww = int(lab.shape[1]), int(lab.shape[2])
    dev = x.device
    with torch.cuda.device(dev):
        lib = _lib.load()
        ws = torch.empty(lib.sh_aux_ce_workspace_bytes(b, c, h, w), dtype=torch.uint8, device=dev)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x) if want_grad else None
        _timed_call("sh_aux_ce_fwdbwd", _p(x), _dtype_code(x), _p(lab), lcode, b, c, h, w, hh, ww, _p(grad), None,
                    _p(out), _p(ws), _stream())
    return out, (grad if want_grad else _empty(dev, x.dtype))


@aux_ce_fwd.register_fake
def _(logits, label, want_grad):
    return (torch.empty(1, dtype=torch.float32, device=logits.device),
            torch.empty_like(logits, memory_format=torch.contiguous_format) if want_grad
            else torch.empty(0, dtype=logits.dtype, device=logits.device))


def _aux_setup(ctx, inputs, output):
    out, grad = output
    ctx.save_for_backward(grad)
    ctx.set_materialize_grads(False)
    ctx.mark_non_differentiable(grad)


def _aux_backward_fn(ctx, g_out, _g_grad):
    (grad,) = ctx.saved_tensors
    if g_out is None:
        return None, None, None
    # out-of-place: the loss is usually weighted (0.4 * aux) and small ([B, C, H/16, W/16])
    return grad * g_out.reshape(-1)[:1].to(grad.dtype), None, None


aux_ce_fwd.register_autograd(_aux_backward_fn, setup_context=_aux_setup)


def aux_cross_entropy(logits: Tensor, label: Tensor) -> Tensor:
    """nn.CrossEntropyLoss(ignore_index=255)(F.interpolate(logits, label.shape[-2:], mode="bilinear",
    align_corners=False), label) -- the aux-head loss of train.py:309-313 -- from the LOW-resolution logits."""
    want = logits.requires_grad and torch.is_grad_enabled()
    out, _ = aux_ce_fwd(logits, label, want)
    return out[0]

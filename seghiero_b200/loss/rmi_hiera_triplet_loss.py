"""RMIHieraTripletLoss -- drop-in for models/loss/rmi_hiera_triplet_loss.py:180-546 on sm_100a kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import hierarchy as H
from .. import ops
from .hiera_triplet_loss import FusedCrossEntropy
from .rmi_tree_triplet_loss import TreeTripletLoss


class RMIHieraTripletLoss(nn.Module):
    """3-level (fine -> mid -> high) tree BCE + RMI lower bound + softmax CE + scheduled triplet.

    Same constructor / forward as the reference; `rmi_pool_*`, `weight`, `cls_score_before` and
    `**kwargs` are accepted and ignored as there (rmi_radius must be 3, the only value the reference's
    arithmetic is defined for with its 9x9 matrices).  The maps are validated (ids inside the level
    sizes) instead of silently indexing garbage.  Observable differences: see HieraTripletLoss.
    """

    def __init__(self, n_fine: int, n_mid: int, n_high: int, fine_to_mid: torch.Tensor, fine_to_high: torch.Tensor,
                 rmi_radius: int = 3, rmi_pool_way: int = 0, rmi_pool_size: int = 3, rmi_pool_stride: int = 3,
                 loss_weight_lambda: float = 0.5, loss_weight: float = 1.0, ignore_index: int = 255,
                 strict: bool = False, fast_path: bool = True):
        super().__init__()
        assert fine_to_mid.dtype == torch.long
        assert fine_to_high.dtype == torch.long
        assert fine_to_mid.numel() == n_fine
        assert fine_to_high.numel() == n_fine
        if ignore_index != 255:
            raise ValueError("only ignore_index=255 is supported (the reference's builders hard-code 255)")
        if rmi_radius != 3:
            raise ValueError("rmi_radius must be 3")
        self.n_fine, self.n_mid, self.n_high = n_fine, n_mid, n_high
        self.fine_to_mid = fine_to_mid.clone()
        self.fine_to_high = fine_to_high.clone()
        self.ignore_index = ignore_index
        self.rmi_radius = rmi_radius
        self.rmi_pool_way = rmi_pool_way
        self.rmi_pool_size = rmi_pool_size
        self.rmi_pool_stride = rmi_pool_stride
        assert self.rmi_pool_size == self.rmi_pool_stride
        if n_fine > 15:
            self.upper_ids = [1, 2, 3, 4, 5, 6, 7, 10, 11, 13, 14, 15]
            self.lower_ids = [8, 9, 12, 16, 17, 18, 19]
        else:
            self.upper_ids = [1, 2, 3, 4]
            self.lower_ids = [5, 6]
        self.loss_weight_lambda = loss_weight_lambda
        self.loss_weight = loss_weight
        self.half_d = self.rmi_radius * self.rmi_radius
        self.d = 2 * self.half_d
        self.kernel_padding = self.rmi_pool_size // 2
        self.triplet_loss = TreeTripletLoss(num_classes=self.n_fine, upper_ids=self.upper_ids,
                                            lower_ids=self.lower_ids, ignore_index=self.ignore_index)
        self.ce = FusedCrossEntropy(ignore_index)
        self.strict = strict
        self.fast_path = fast_path      # False: run the generic kernels even for tree-shaped hierarchies (tests)
        self.last_stats: dict = {}
        # host copies of the maps (the kernels' tables are built from them once per device); validated here
        # (raises ValueError on out-of-range map entries)
        self._f2m = [int(v) for v in self.fine_to_mid.cpu().tolist()]
        self._f2h = [int(v) for v in self.fine_to_high.cpu().tolist()]
        H.three_level_tables(n_fine, n_mid, n_high, self._f2m, self._f2h)

    def forward(self, step, embedding, cls_score_before, cls_score, label, weight=None, **kwargs):
        """`cls_score` [B, n_fine+n_mid+n_high, H, W] at the label's resolution as in the reference, or at the head's
        own (e.g. H/4): it is then upsampled inside the op exactly like train.py:282-284 does outside and the gradient
        arrives at the head's resolution.  `label` may be int64 (reference), int32 or uint8."""
        step_d = ops.step_tensor(step, cls_score.device)
        use_triplet = self.triplet_loss is not None
        out, ws, _xf, _sel, _kc, _tl, trip, status = ops.hier3_fwd(
            cls_score, label, embedding, step_d, int(self.n_fine), int(self.n_mid), int(self.n_high), self._f2m,
            self._f2h, [int(v) for v in self.upper_ids], [int(v) for v in self.lower_ids],
            float(self.loss_weight_lambda), float(self.loss_weight), 160000.0 if self.n_fine > 15 else 60000.0,
            bool(self.fast_path), use_triplet)
        if not torch.compiler.is_compiling():
            # detached: a live grad_fn here would keep the whole autograd graph of the call alive until the next one
            self.last_stats = {"out": out.detach(), "trip": trip, "status": status}
        if self.strict:
            if int(ws[16:24].view(torch.int64).item()):
                raise RuntimeError("Class values must be smaller than num_classes.")
            if status.numel() and int(status[1].item()):
                raise ValueError("list.remove(x): x not in list (label in neither upper_ids nor lower_ids)")
        return out[0]

    def pass2_kernel(self, cls_score, label) -> str:
        """Name of the pass-2 kernel that serves this call (diagnostics / bench stage names)."""
        from .. import _lib
        tab, n_mh, fast_ok = H.three_level_tables(self.n_fine, self.n_mid, self.n_high, self._f2m, self._f2h)
        hh, ww = label.shape[-2:]
        same = tuple(cls_score.shape[-2:]) == (hh, ww)
        ptr = ops._p(cls_score) if same else None
        kind = _lib.load().sh_rmi3_pass2_kind(ptr, ptr, ops._dtype_code(cls_score), int(hh), int(ww), self.n_fine, self.n_mid,
                                              self.n_high, int(fast_ok) if self.fast_path else 0)
        return ("k3_pass2", "k3f_pass2", "k3t_pass2")[kind]

    def uses_fast_path(self, cls_score, label) -> bool:
        """Whether the warp-specialised / tiled kernels (not the generic ones) serve this call (diagnostics)."""
        from .. import _lib
        tab, n_mh, fast_ok = H.three_level_tables(self.n_fine, self.n_mid, self.n_high, self._f2m, self._f2h)
        hh, ww = label.shape[-2:]
        same = tuple(cls_score.shape[-2:]) == (hh, ww)
        return bool(self.fast_path and _lib.load().sh_rmi3_fast_path(
            ops._p(cls_score) if same else None, None, ops._dtype_code(cls_score), int(hh), int(ww), self.n_fine,
            self.n_mid, self.n_high, int(fast_ok)))

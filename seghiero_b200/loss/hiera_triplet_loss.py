"""HieraTripletLoss -- drop-in for models/loss/hiera_triplet_loss.py:110-211 on sm_100a kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .tree_triplet_loss import TreeTripletLoss


class FusedCrossEntropy(nn.Module):
    """Stands where the reference keeps `self.ce = CrossEntropyLoss()` (hiera_triplet_loss.py:141): same attribute
    names, no parameters.  The per-level softmax CE itself (cross_entropy_loss.py:7-30, mean over ALL pixels) is
    computed inside the fused loss kernel from the same read of the logits, so this module is never called by
    the parent; calling it directly is refused rather than answered by an eager fallback."""

    def __init__(self, ignore_index: int = 255):
        super().__init__()
        self.use_sigmoid = False
        self.use_mask = False
        self.reduction = "mean"
        self.loss_weight = 1.0
        self.class_weight = None
        self.ignore_index = ignore_index

    def forward(self, *args, **kwargs):
        raise RuntimeError("the per-level cross entropy is fused into the hierarchical loss kernel; "
                           "call the parent loss module (seghiero_b200 has no eager fallback)")


class HieraTripletLoss(nn.Module):
    """2-level (fine -> coarse) hierarchical BCE + softmax CE + scheduled triplet loss.

    Constructor and forward signatures are the reference's.  `use_sigmoid`, `weight`,
    `cls_score_before` and `**kwargs` are accepted and ignored exactly as there.  Differences
    that a caller can observe (all documented in DESIGN.md):
      * no host synchronisation: `step`, the `ready` gate and the schedule stay on the device;
        consequently `embedding.grad` is a zero tensor (not None) when no triplet class was found;
      * labels outside [0, num_classes) U {255} (where the reference's F.one_hot raises) make the
        loss NaN; pass `strict=True` to get the RuntimeError at the cost of one sync per call;
      * CUDA tensors only.
    """

    def __init__(self, num_classes: int, hiera_map: list, hiera_index: list, ignore_index: int = 255,
                 use_sigmoid: bool = False, loss_weight: float = 1.0, strict: bool = False, fast_path: bool = True):
        super().__init__()
        if ignore_index != 255:
            raise ValueError("only ignore_index=255 is supported (the reference's builders hard-code 255)")
        self.num_classes = num_classes
        self.hiera_map = hiera_map
        self.hiera_index = hiera_index
        self.ignore_index = ignore_index
        self.ce = FusedCrossEntropy(ignore_index)
        self.triplet_loss_fn = TreeTripletLoss(num_classes=len(hiera_map), hiera_map=hiera_map,
                                               hiera_index=hiera_index, ignore_index=ignore_index)
        self.loss_weight = loss_weight
        self.strict = strict
        self.fast_path = fast_path      # False: run the any-bucket kernels even for tree-shaped hierarchies (tests)
        self.last_stats: dict = {}
        if not ops.two_level_supported(int(num_classes), len(hiera_index), True):
            raise ValueError(f"HieraTripletLoss on sm_100a: {int(num_classes) + len(hiera_index)} channels exceed the "
                             "shared-memory tiling of the 2-level kernels (<= ~220 channels)")

    def forward(self, step, embedding, cls_score_before, cls_score, label, weight=None, **kwargs):
        """`cls_score` [B, n_fine+n_coarse, H, W] at the label's resolution as in the reference, or at the head's own
        (e.g. H/4): it is then upsampled inside the op exactly like train.py:282-284 does outside
        (F.interpolate(size=label.shape[-2:], mode="bilinear", align_corners=False)) and the gradient arrives at the
        head's resolution.  `label` may be int64 (reference), int32 or uint8."""
        step_d = ops.step_tensor(step, cls_score.device)
        want = cls_score.requires_grad and torch.is_grad_enabled()
        index_flat = [int(v) for r in self.hiera_index for v in r]
        out, _grad, counts, _sel, _kc, _tl, trip, status = ops.hier2_fwd(
            cls_score, label, embedding, step_d, int(self.num_classes), [int(v) for v in self.hiera_map], index_flat,
            float(self.loss_weight), 80000.0, bool(self.fast_path), want)
        if not torch.compiler.is_compiling():
            # detached: a live grad_fn here would keep the whole autograd graph of the call alive until the next one
            self.last_stats = {"out": out.detach(), "counts": counts, "trip": trip, "status": status}
        if self.strict:
            if int(counts[2].item()):
                raise RuntimeError("Class values must be smaller than num_classes.")
            if status.numel() and int(status[1].item()):
                raise IndexError("label outside hiera_map in the triplet term")
        return out[0]

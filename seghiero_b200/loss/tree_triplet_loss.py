"""TreeTripletLoss, hierarchy flavour -- drop-in for models/loss/tree_triplet_loss.py:6-65."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class TreeTripletLoss(nn.Module):
    """Same constructor / forward as the reference.  forward returns (loss | None, LongTensor[1]).

    Returning None needs the class count on the host, so this *standalone* module syncs once per
    call exactly where the reference does (`min(...)`, `.cuda()` tensors built from Python ints);
    the fused loss modules use the sync-free internal path instead.
    Labels outside [0, num_classes) U {255} raise IndexError like `self.hiera_map[ii]` would.
    """

    def __init__(self, num_classes, hiera_map, hiera_index, ignore_index=255):
        super().__init__()
        self.ignore_label = ignore_index
        self.num_classes = num_classes
        self.hiera_map = hiera_map
        self.hiera_index = hiera_index

    def forward(self, feats, labels=None, max_triplet=200):
        trip, status, _sel, _kc, _tl = ops.triplet_fwd(feats, labels, 0, [int(v) for v in self.hiera_map], [int(v) for r in self.hiera_index for v in r],
                                                       int(max_triplet))
        ready, err = (int(v) for v in status.tolist())
        if err:
            raise IndexError("TreeTripletLoss: label outside [0, num_classes) (hiera_map lookup)")
        count = trip[1:2].to(torch.int64)
        if not ready:
            return None, count
        return trip[0], count

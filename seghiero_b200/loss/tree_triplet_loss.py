"""TreeTripletLoss, hierarchy flavour -- drop-in for models/loss/tree_triplet_loss.py:6-65."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import hierarchy as H
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
        ops._need_cuda(feats, labels)
        key = ("t0", tuple(int(v) for v in self.hiera_map),
               tuple(tuple(int(v) for v in r) for r in self.hiera_index))
        tab, ncls = ops.device_table(key, lambda: H.triplet_tables_hierarchy(self.hiera_map, self.hiera_index),
                                     feats.device)
        holder = {}
        loss = ops.TripletFn.apply(feats, labels, 0, tab, ncls, int(max_triplet), holder)
        st = holder["state"]
        ready, err = (int(v) for v in st.status.tolist())
        if err:
            raise IndexError("TreeTripletLoss: label outside [0, num_classes) (hiera_map lookup)")
        count = st.trip[1:2].to(torch.int64)
        if not ready:
            return None, count
        return loss, count

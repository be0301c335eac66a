"""TreeTripletLoss, id-list flavour -- drop-in for models/loss/rmi_tree_triplet_loss.py:5-70."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class TreeTripletLoss(nn.Module):
    """Same constructor / forward as the reference: positives come from the id list that holds the
    anchor class, negatives from the other list; classes 0 and 255 never anchor.  A present class in
    neither list raises ValueError, as the reference's `list.remove` does."""

    def __init__(self, num_classes, upper_ids, lower_ids, ignore_index=255):
        super().__init__()
        self.ignore_label = ignore_index
        self.num_classes = num_classes
        self.upper_ids = upper_ids
        self.lower_ids = lower_ids

    def forward(self, feats, labels=None, max_triplet=200):
        trip, status, _sel, _kc, _tl = ops.triplet_fwd(feats, labels, 1, [int(v) for v in self.upper_ids], [int(v) for v in self.lower_ids],
                                                       int(max_triplet))
        ready, err = (int(v) for v in status.tolist())
        if err:
            raise ValueError("list.remove(x): x not in list (label in neither upper_ids nor lower_ids)")
        count = trip[1:2].to(torch.int64)
        if not ready:
            return None, count
        return trip[0], count

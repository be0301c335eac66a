"""seghiero_b200: B200-native (sm_100a) hierarchical-loss path of SegHiero.

Public surface mirrors the reference's `models/loss` package plus the functional
target builders and the hierarchical argmax decode:

    from seghiero_b200 import HieraTripletLoss, RMIHieraTripletLoss, TreeTripletLoss
"""
from .loss import HieraTripletLoss, RMIHieraTripletLoss, TreeTripletLoss
from .loss.rmi_tree_triplet_loss import TreeTripletLoss as IdListTreeTripletLoss
from .ops import (aux_cross_entropy, build_fine_to_level_map, colorize, hierarchical_argmax, targets_gather,
                  targets_three_level, targets_two_level)

__all__ = ["HieraTripletLoss", "RMIHieraTripletLoss", "TreeTripletLoss", "IdListTreeTripletLoss",
           "hierarchical_argmax", "targets_gather", "targets_three_level", "targets_two_level",
           "build_fine_to_level_map", "colorize", "aux_cross_entropy"]

from .hiera_triplet_loss import HieraTripletLoss
from .rmi_hiera_triplet_loss import RMIHieraTripletLoss
from .tree_triplet_loss import TreeTripletLoss
from . import rmi_tree_triplet_loss

__all__ = ["HieraTripletLoss", "RMIHieraTripletLoss", "TreeTripletLoss", "rmi_tree_triplet_loss"]

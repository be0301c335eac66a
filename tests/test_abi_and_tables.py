"""CPU-only checks of the drop-in boundary and the host logic: the C-ABI library loads and exports every symbol
include/seghiero_b200.h declares (no compute calls), the ctypes signature table mirrors the header, the host-side
hierarchy tables obey the orders the kernels rely on, and the module shells keep the reference's constructor
behaviour (models/loss/hiera_triplet_loss.py:129-150, rmi_hiera_triplet_loss.py:226-287)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import seghiero_b200 as sb
from seghiero_b200 import _lib, hierarchy as H
from tests.util import F2H, F2M, HI, HM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "seghiero_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b(?:int|size_t)\s+(sh_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)


def test_library_exports_every_declared_symbol():
    decl = _declared()
    names = [n for n, _ in decl]
    assert len(names) >= 16 and len(set(names)) == len(names)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    for n, args in decl:
        n_args = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n_args == len(_lib.SIGNATURES[n][1]), f"{n}: header has {n_args} arguments"
    assert _lib.load() is _lib.load()


def test_workspace_size_is_pure_host_arithmetic():
    lib = _lib.load()
    small = lib.sh_rmi3_workspace_bytes(1, 64, 64, 19, 7, 2)
    big = lib.sh_rmi3_workspace_bytes(2, 64, 64, 19, 7, 2)
    assert 0 < small < big
    assert lib.sh_bce2_grid(2, 512 * 512, 26, 7) > 0


def test_fast_tree_order_properties():
    tab, n_mh, fast_ok = H.three_level_tables(19, 7, 2, F2M, F2H)
    assert fast_ok and n_mh == 7
    c = 28
    base = 2 * 19 + 8 + n_mh + 7
    order = tab[base + c: base + 2 * c].view(np.uint32)
    aux = tab[base + 2 * c: base + 3 * c].view(np.uint32)
    chans = [int(o >> 24) for o in order]
    assert sorted(chans) == list(range(c))                       # every channel exactly once
    pos = {ch: i for i, ch in enumerate(chans)}
    for f in range(19):
        assert pos[f] < pos[19 + F2M[f]] < pos[19 + 7 + F2H[f]]   # children before their mid before their high
        assert int(aux[pos[f]] & 0xff) == F2M[f] and int(aux[pos[f]] >> 8) == F2H[f]
    firsts = [i for i, o in enumerate(order) if (int(o) >> 16) & 1]
    assert len(firsts) == 7                                       # one "first of a mid group" flag per mid
    assert sum(1 for o in order if (int(o) >> 16) & 4) == 2       # one per high group
    # a mid whose children sit under two highs is not a tree: generic kernels only
    bad_f2h = list(F2H)
    bad_f2h[0] = 1
    assert not H.three_level_tables(19, 7, 2, F2M, bad_f2h)[2]
    with pytest.raises(ValueError):
        H.three_level_tables(19, 7, 2, [9] * 19, F2H)


def test_two_level_tables():
    assert H.two_level_is_tree(19, HI)
    assert not H.two_level_is_tree(9, [[0, 4], [2, 6], [6, 7]])
    assert H.two_level_is_tree(9, [[0, 4], [5, 7], [7, 7]])
    blob, n_fb, lut_size = H.two_level_tables(19, HI)
    nc = len(HI)
    lut = blob[-lut_size:]
    for i, (s, e) in enumerate(HI):
        assert (lut[s:e] == i).all()
    owner = blob[2 * nc: 2 * nc + 19]
    assert list(owner) == HM and n_fb == 19


def test_module_shells_keep_the_reference_contract():
    with pytest.raises(AssertionError):
        sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M, dtype=torch.int32), torch.tensor(F2H))
    with pytest.raises(AssertionError):
        sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M[:5]), torch.tensor(F2H))
    m3 = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), loss_weight_lambda=0.25)
    for name in ("fine_to_mid", "fine_to_high", "n_fine", "n_mid", "n_high", "upper_ids", "lower_ids",
                 "loss_weight_lambda", "half_d", "d", "triplet_loss", "ignore_index", "loss_weight"):
        assert hasattr(m3, name)
    assert m3.half_d == 9 and m3.d == 18 and len(list(m3.parameters())) == 0
    m2 = sb.HieraTripletLoss(19, HM, HI, ignore_index=255, use_sigmoid=True, loss_weight=2.0)
    for name in ("hiera_map", "hiera_index", "num_classes", "ignore_index", "loss_weight", "ce", "triplet_loss_fn"):
        assert hasattr(m2, name)
    # no CPU fallback: host tensors are refused, loudly
    x = torch.zeros(1, 26, 8, 8, requires_grad=True)
    with pytest.raises(RuntimeError):
        m2(torch.tensor([0]), torch.zeros(1, 4, 1, 1), None, x, torch.zeros(1, 8, 8, dtype=torch.long))

"""Pin the oracle's restatements of the SURVEY 8f rows (N1-N4, a3) against vectors produced by the reference code /
the stock torch calls the reference makes (tests/golden/make_golden_r2.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hiera_oracle as O
from tests.util import F2H, F2M, HI, HM, rel


def test_interpolate_restatement_matches_torch(golden):
    for name in ("n1_three_level_up4", "n1_two_level_up"):
        g = golden(name)
        full = O.interpolate_bilinear(torch.from_numpy(g["x"]), g["label"].shape[-2:])
        assert rel(full.numpy(), g["full"]) <= 1e-6
        assert np.abs(full.numpy() - g["full"]).max() <= 1e-5


def test_n1_three_level_from_head_logits(golden):
    g = golden("n1_three_level_up4")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    emb = torch.from_numpy(g["emb"]).requires_grad_(True)
    full = O.interpolate_bilinear(x, g["label"].shape[-2:])
    loss, _ = O.rmi_hiera_triplet_loss(int(g["step"]), emb, full, g["label"], 19, 7, 2, F2M, F2H)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel(x.grad.numpy(), g["dx"]) <= 1e-5
    assert rel(emb.grad.numpy(), g["demb"]) <= 1e-5


def test_n1_two_level_from_head_logits(golden):
    g = golden("n1_two_level_up")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    emb = torch.from_numpy(g["emb"]).requires_grad_(True)
    full = O.interpolate_bilinear(x, g["label"].shape[-2:])
    loss, _ = O.hiera_triplet_loss(int(g["step"]), emb, full, g["label"], 19, HM, HI)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel(x.grad.numpy(), g["dx"]) <= 1e-5


def test_n2_aux_cross_entropy(golden):
    g = golden("n2_aux_ce")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    loss = O.aux_cross_entropy(x, g["label"])
    (0.4 * loss).backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel(x.grad.numpy(), g["dx_times_0p4"]) <= 1e-5


def test_n3_decode_from_head_logits(golden):
    g = golden("n3_decode_up4")
    pf, pm, ph = O.argmax_decode_upsampled(torch.from_numpy(g["x"]), [19, 7, 2], g["label"].shape[-2:])
    # the restatement's fp32 interpolation may differ from ATen's in the last ulp: ties between channels can flip
    for got, ref in ((pf, g["pf"]), (pm, g["pm"]), (ph, g["ph"])):
        assert (np.asarray(got) != ref).mean() <= 1e-3
    g2 = golden("n3_decode_up_any")
    pf, pm, ph = O.argmax_decode_upsampled(torch.from_numpy(g2["x"]), [19, 7, 2], g2["pf"].shape[-2:])
    for got, ref in ((pf, g2["pf"]), (pm, g2["pm"]), (ph, g2["ph"])):
        assert (np.asarray(got) != ref).mean() <= 1e-3


def test_n4_maps_and_colorize(golden):
    import seghiero_b200 as sb
    g = golden("n4_maps_colorize")
    for cfg, ref, n in ((g["cfg_a"], g["map_a"], 19), (g["cfg_b"], g["map_b"], 9)):
        cfg_l = [[int(v) for v in r if v >= 0] for r in cfg]
        assert np.array_equal(O.build_fine_to_level_map(cfg_l, n), ref)
        got = sb.build_fine_to_level_map(cfg_l, n)                      # the product's host-side builder (no GPU needed)
        assert got.dtype == torch.long and np.array_equal(got.numpy(), ref)
    assert np.array_equal(O.targets_dataloader(g["fine_mask"], g["map_a"]), g["coarse_mask"])
    assert np.array_equal(O.colorize(g["mask"], g["colormap"]), g["rgb"])
    with pytest.raises(ValueError):
        sb.build_fine_to_level_map([[0, 1], [3, 4]], 5)                  # id 2 unmapped (dataloader.py:31-33)
    with pytest.raises(ValueError):
        sb.build_fine_to_level_map([[0, 1, 2]], 3)
    with pytest.raises(AssertionError):
        sb.build_fine_to_level_map([[0, 5]], 5)

"""Pin the CPU oracle against vectors produced by the reference modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hiera_oracle as O

HI = [[0, 2], [2, 5], [5, 8], [8, 10], [10, 11], [11, 13], [13, 19]]
HM = [0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 6, 6, 6, 6, 6, 6]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _inputs(g):
    x = torch.from_numpy(g["x"])
    if "is_bf16" in g and bool(g["is_bf16"]):
        x = x.bfloat16()
    return x.requires_grad_(True)


@pytest.mark.parametrize("name", ["two_level_blob", "two_level_iid", "two_level_bf16", "two_level_allvoid"])
def test_two_level_module(golden, name):
    g = golden(name)
    x = _inputs(g)
    emb = torch.from_numpy(g["emb"]).requires_grad_(True)
    loss, parts = O.hiera_triplet_loss(int(g["step"]), emb, x, g["label"], 19, HM, HI,
                                       loss_weight=float(g["loss_weight"]))
    loss.backward()
    bf16 = bool(g["is_bf16"])
    tol = 2e-2 if bf16 else 1e-5
    assert np.array_equal(parts["targets"][1], g["tc"])
    assert parts["count"] == int(g["count"][0])
    assert abs(float(loss) - float(g["loss"])) <= tol * abs(float(g["loss"]))
    if not bf16:
        assert abs(parts["hiera"] - float(g["hiera"])) <= 1e-5 * abs(float(g["hiera"])) + 1e-7
        assert abs(parts["ce"][0] - float(g["ce_f"])) <= 1e-5 * abs(float(g["ce_f"])) + 1e-7
        assert abs(parts["ce"][1] - float(g["ce_c"])) <= 1e-5 * abs(float(g["ce_c"])) + 1e-7
        if parts["triplet"] is not None:
            assert abs(parts["triplet"] - float(g["triplet"])) <= 1e-5 * abs(float(g["triplet"]))
    assert rel(x.grad.float().numpy(), g["dx"]) <= tol
    if not bool(g["demb_none"]):
        assert rel(emb.grad.numpy(), g["demb"]) <= 1e-5
    else:
        assert emb.grad is None or float(emb.grad.abs().max()) == 0.0


@pytest.mark.parametrize("name", ["three_level_blob", "three_level_iid", "three_level_small", "three_level_bf16"])
def test_three_level_module(golden, name):
    g = golden(name)
    bf16 = bool(g["is_bf16"])
    tol = 2e-2 if bf16 else 1e-5
    nf, nm, nh = int(g["nf"]), int(g["nm"]), int(g["nh"])
    for tag, lam in (("", float(g["lam"])), ("_lam0", 0.0)):
        x = _inputs(g)
        emb = torch.from_numpy(g["emb"]).requires_grad_(True)
        loss, parts = O.rmi_hiera_triplet_loss(int(g["step"]), emb, x, g["label"], nf, nm, nh, g["f2m"], g["f2h"],
                                               loss_weight_lambda=lam, loss_weight=float(g["loss_weight"]))
        loss.backward()
        assert np.array_equal(parts["targets"][1], g["tm"])
        assert np.array_equal(parts["targets"][2], g["th"])
        assert parts["count"] == int(g["count"][0])
        assert abs(float(loss) - float(g["loss" + tag])) <= tol * abs(float(g["loss" + tag])), (tag, float(loss))
        assert rel(x.grad.float().numpy(), g["dx" + tag]) <= tol, tag
        if tag == "" and not bool(g["demb_none"]):
            assert rel(emb.grad.numpy(), g["demb"]) <= 1e-5


def test_three_level_flat_predictions(golden):
    g = golden("three_level_flat")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    f2m, f2h = np.array(HM), np.array([0] * 11 + [1] * 8)
    loss, _ = O.rmi_hiera_triplet_loss(0, torch.zeros(1, 4, 2, 2), x, g["label"], 19, 7, 2, f2m, f2h)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel(x.grad.numpy(), g["dx"]) <= 1e-4   # ill-conditioned regime: reference itself is noisy here


def test_decode(golden):
    g = golden("decode")
    pf, pm, ph = O.argmax_decode(torch.from_numpy(g["x"]), [19, 7, 2])
    assert np.array_equal(pf, g["pf"]) and np.array_equal(pm, g["pm"]) and np.array_equal(ph, g["ph"])
    assert O.pixel_accuracy_counts(pf, g["label"]) == (int(g["correct"]), int(g["total"]))


def test_survey_known_answers():
    """values recorded from the reference during the survey."""
    f2m, f2h = np.array(HM), np.array([0] * 11 + [1] * 8)
    torch.manual_seed(1)
    label = torch.randint(0, 19, (2, 40, 56))
    label[torch.rand(2, 40, 56) < 0.15] = 255
    x = torch.randn(2, 28, 40, 56) * 2
    loss, parts = O.rmi_hiera_triplet_loss(0, None, x, label, 19, 7, 2, f2m, f2h, loss_weight_lambda=1.0,
                                           with_triplet=False)
    assert abs(float(loss) - 89.24185943603516) < 1e-5 * 89.24
    assert abs(parts["hiera"] - 27.01216697692871) < 1e-5 * 27.0
    assert abs(parts["rmi"] - 67.8801498413086) < 1e-5 * 67.9
    torch.manual_seed(2)
    label = torch.randint(0, 19, (2, 128, 160))
    label[torch.rand(2, 128, 160) < 0.15] = 255
    x = torch.randn(2, 26, 128, 160) * 2
    emb = torch.nn.functional.normalize(torch.randn(2, 64, 4, 5), dim=1)
    loss, parts = O.hiera_triplet_loss(60000, emb, x, label, 19, HM, HI)
    assert abs(float(loss) - 22.305631637573242) < 1e-5 * 22.3
    assert parts["count"] == 15
    assert abs(parts["factor"] - 0.42677669529663687) < 1e-12


def test_builders_and_errors():
    lab = np.array([[0, 1, 5, 18, 255, 19, -1]])
    _, tc = O.targets_two_level(lab, HI)
    assert tc.tolist() == [[0, 0, 2, 6, 255, 255, 255]]
    _, tc = O.targets_two_level(np.array([3, 4]), [[0, 5], [3, 4]])   # overlap: later bucket wins
    assert tc.tolist() == [1, 0]
    with pytest.raises(IndexError):
        O.targets_three_level(np.array([19]), np.arange(19), np.arange(19))
    with pytest.raises(IndexError):
        O.targets_dataloader(np.array([255]), np.arange(19))
    assert O.build_fine_to_level_map([[0, 1], [2], [3, 4]], 5).tolist() == [0, 0, 1, 2, 2]
    with pytest.raises(ValueError):
        O.build_fine_to_level_map([[0, 1]], 3)
    # class in neither id list -> ValueError
    emb = torch.zeros(1, 4, 2, 2)
    with pytest.raises(ValueError):
        O.triplet_id_lists(emb, np.full((1, 2, 2), 8), [1, 2, 3, 4], [5, 6])


def test_nearest_matches_torch():
    for n_in, n_out in [(512, 16), (100, 7), (37, 5), (8, 8), (5, 9), (1024, 33)]:
        lab = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1)
        ref = torch.nn.functional.interpolate(lab, (n_out, 1), mode="nearest").view(-1).long().numpy()
        assert np.array_equal(O.nearest_rows(n_in, n_out), ref)

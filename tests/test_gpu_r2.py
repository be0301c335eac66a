"""GPU parity tests of the round-2 rows: uint8 / int32 labels end to end (N4), the losses fed with the head's
low-resolution logits (N1), the aux-head cross entropy (N2), decode from low-resolution logits (N3), GPU colourise
and dataloader target build (N4 / a3), the torch custom-op boundary under torch.compile, and parity at the
benchmark shapes (config 3 batch 8, config 4 batch 32).  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hiera_oracle as O
from tests.util import F2H, F2M, HI, HM, blob_labels, iid_labels, rel, to_np

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def sb():
    import seghiero_b200
    from seghiero_b200 import _lib
    _lib.load()
    assert torch.cuda.is_available()
    return seghiero_b200


# ------------------------------------------------------------------------------------------------
# N4: label element types
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ldt", [torch.uint8, torch.int32])
def test_label_dtypes_bit_exact_integer_paths(sb, ldt):
    g = torch.Generator().manual_seed(6)
    lab = iid_labels(g, 3, 37, 53, 19, 0.2)
    out = sb.targets_two_level(lab.to(ldt).cuda(), HI)
    assert out.dtype == ldt
    assert np.array_equal(out.cpu().numpy().astype(np.int64), O.targets_two_level(lab.numpy(), HI)[1])
    mid, high = sb.targets_three_level(lab.to(ldt).cuda(), torch.tensor(F2M), torch.tensor(F2H))
    _, om, oh = O.targets_three_level(lab.numpy(), F2M, F2H)
    assert mid.dtype == ldt and np.array_equal(mid.cpu().numpy().astype(np.int64), om)
    assert np.array_equal(high.cpu().numpy().astype(np.int64), oh)
    # unaligned views (odd offset into the buffer): the scalar kernels take them
    flat = lab.to(ldt).cuda().reshape(-1)
    sl = flat[1:1 + 36 * 53].reshape(1, 36, 53)
    out = sb.targets_two_level(sl, HI)
    assert np.array_equal(out.cpu().numpy().astype(np.int64),
                          O.targets_two_level(lab.reshape(-1)[1:1 + 36 * 53].reshape(1, 36, 53).numpy(), HI)[1])
    fm = torch.randint(0, 19, (64, 48), generator=g)
    out = sb.targets_gather(fm.to(ldt).cuda(), torch.tensor(F2H))
    assert np.array_equal(out.cpu().numpy().astype(np.int64), O.targets_dataloader(fm.numpy(), F2H))
    x = torch.randn(2, 28, 24, 40, generator=g)
    lab2 = iid_labels(g, 2, 24, 40, 19, 0.2)
    p64, c64 = sb.hierarchical_argmax(x.cuda(), [19, 7, 2], lab2.cuda())
    pl, cl = sb.hierarchical_argmax(x.cuda(), [19, 7, 2], lab2.to(ldt).cuda(), out_dtype=torch.uint8)
    assert c64.tolist() == cl.tolist()
    assert all(torch.equal(a, b.long()) for a, b in zip(p64, pl))


def test_int64_unaligned_label_view(sb):
    """a contiguous int64 view that is only 8-byte aligned must not be rejected."""
    g = torch.Generator().manual_seed(8)
    lab = iid_labels(g, 2, 15, 17, 19, 0.2).cuda()       # odd H*W: labels[1:] is 8-byte aligned only
    out = sb.targets_two_level(lab[1:], HI)
    assert np.array_equal(out.cpu().numpy(), O.targets_two_level(lab[1:].cpu().numpy(), HI)[1])


@pytest.mark.parametrize("ldt", [torch.uint8, torch.int32])
def test_label_dtypes_losses_identical_to_int64(sb, ldt):
    g = torch.Generator().manual_seed(17)
    lab = blob_labels(g, 2, 48, 132, 19, 7, 0.1)
    lab[1] = iid_labels(g, 1, 48, 132, 19, 0.15)[0]
    x3 = (torch.randn(2, 28, 48, 132, generator=g) * 2).cuda()
    x2 = (torch.randn(2, 26, 48, 132, generator=g) * 2).cuda()
    emb = F.normalize(torch.randn(2, 12, 6, 16, generator=g), dim=1).cuda()
    step = torch.tensor([100000]).cuda()
    for mod, x in ((sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H)), x3),
                   (sb.HieraTripletLoss(19, HM, HI), x2)):
        res = []
        for lt in (torch.int64, ldt):
            xc, ec = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
            loss = mod(step, ec, None, xc, lab.to(lt).cuda())
            loss.backward()
            res.append((loss.detach().clone(), xc.grad.clone(), ec.grad.clone()))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        assert rel(to_np(res[1][2]), to_np(res[0][2])) <= 1e-6      # float atomics in the embedding gradient


def test_colorize_and_dataloader_targets(sb, golden):
    g = golden("n4_maps_colorize")
    for dt in (torch.int32, torch.int64):
        rgb = sb.colorize(torch.from_numpy(g["mask"]).to(dt).cuda(), g["colormap"])
        assert rgb.dtype == torch.uint8 and np.array_equal(rgb.cpu().numpy(), g["rgb"])
    rgb = sb.colorize(torch.from_numpy(g["fine_mask"]).to(torch.uint8).cuda(), g["colormap"])
    assert np.array_equal(rgb.cpu().numpy(), O.colorize(g["fine_mask"], g["colormap"]))
    with pytest.raises(IndexError):
        sb.colorize(torch.full((4, 5), 19, dtype=torch.int32).cuda(), g["colormap"])
    # full-size image, odd width (unvectorised tail), against the oracle
    gen = torch.Generator().manual_seed(2)
    m = torch.randint(-1, 19, (1023, 2047), generator=gen, dtype=torch.int32)
    assert np.array_equal(sb.colorize(m.cuda(), g["colormap"]).cpu().numpy(), O.colorize(m.numpy(), g["colormap"]))
    fmap = sb.build_fine_to_level_map([[int(v) for v in r if v >= 0] for r in g["cfg_a"]], 19)
    out = sb.targets_gather(torch.from_numpy(g["fine_mask"]).cuda(), fmap)
    assert np.array_equal(out.cpu().numpy(), g["coarse_mask"])


# ------------------------------------------------------------------------------------------------
# N1: losses fed with the head's logits
# ------------------------------------------------------------------------------------------------
def test_n1_golden(sb, golden):
    g = golden("n1_three_level_up4")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    emb = torch.from_numpy(g["emb"]).cuda().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    loss = mod(torch.tensor([int(g["step"])]), emb, None, x, torch.from_numpy(g["label"]).to(torch.uint8).cuda())
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= FP32_TOL * abs(float(g["loss"]))
    assert x.grad.shape == x.shape and rel(to_np(x.grad), g["dx"]) <= FP32_TOL
    assert rel(to_np(emb.grad), g["demb"]) <= 10 * FP32_TOL
    g = golden("n1_two_level_up")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    emb = torch.from_numpy(g["emb"]).cuda().requires_grad_(True)
    loss = sb.HieraTripletLoss(19, HM, HI)(torch.tensor([int(g["step"])]), emb, None, x,
                                           torch.from_numpy(g["label"]).cuda())
    (loss * 0.5).backward()
    assert abs(float(loss) - float(g["loss"])) <= FP32_TOL * abs(float(g["loss"]))
    assert rel(to_np(x.grad), 0.5 * g["dx"]) <= FP32_TOL


@pytest.mark.parametrize("case", [
    dict(b=2, c=28, h=16, w=24, H=64, W=96, dtype=torch.float32),          # the head's 4x geometry
    dict(b=1, c=28, h=9, w=13, H=40, W=52, dtype=torch.float32),           # general scale
    dict(b=1, c=28, h=16, w=32, H=64, W=128, dtype=torch.bfloat16),
])
def test_n1_three_level_vs_torch_interpolate(sb, case):
    """Same module, logits at the head's resolution vs F.interpolate outside (what train.py:282-284 does)."""
    g = torch.Generator().manual_seed(case["H"] + case["w"])
    b, c, h, w, hh, ww = (case[k] for k in ("b", "c", "h", "w", "H", "W"))
    lab = blob_labels(g, b, hh, ww, 19, 8, 0.1).cuda()
    xl = (torch.randn(b, c, h, w, generator=g) * 2).to(case["dtype"]).cuda()
    emb = F.normalize(torch.randn(b, 12, 4, 6, generator=g), dim=1).cuda()
    step = torch.tensor([100000]).cuda()
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    xa = xl.clone().requires_grad_(True)
    la = mod(step, emb, None, xa, lab)
    la.backward()
    xb = xl.clone().requires_grad_(True)
    full = F.interpolate(xb, size=(hh, ww), mode="bilinear", align_corners=False)
    lb = mod(step, emb, None, full, lab)
    lb.backward()
    tol = FP32_TOL if case["dtype"] == torch.float32 else BF16_TOL
    assert abs(float(la) - float(lb)) <= tol * abs(float(lb))
    assert rel(to_np(xa.grad), to_np(xb.grad)) <= tol
    # and against the oracle (CPU restatement of interpolate + loss)
    if case["dtype"] == torch.float32:
        xr = xl.cpu().clone().requires_grad_(True)
        ref, _ = O.rmi_hiera_triplet_loss(100000, emb.cpu(), O.interpolate_bilinear(xr, (hh, ww)), lab.cpu(), 19, 7, 2,
                                          F2M, F2H)
        ref.backward()
        assert abs(float(la) - float(ref)) <= FP32_TOL * abs(float(ref))
        assert rel(to_np(xa.grad), to_np(xr.grad)) <= FP32_TOL


def test_upsample_kernels_against_torch(sb):
    from seghiero_b200 import ops
    g = torch.Generator().manual_seed(3)
    for (h, w, hh, ww) in ((8, 12, 32, 48), (7, 5, 23, 31), (6, 10, 6, 10), (3, 5, 48, 80)):
        for dt in (torch.float32, torch.bfloat16):
            x = torch.randn(3, 5, h, w, generator=g).to(dt).cuda()
            got = ops._upsampled(x, hh, ww)
            ref = F.interpolate(x, size=(hh, ww), mode="bilinear", align_corners=False)
            assert got.dtype == dt
            tol = 1e-6 if dt == torch.float32 else 1e-2
            assert float((got.float() - ref.float()).abs().max()) <= tol * max(1.0, float(ref.float().abs().max()))


@pytest.mark.parametrize("planes,h,w,dt", [(6, 9, 37, torch.float32), (3, 16, 64, torch.float32), (5, 1, 40, torch.float32),
                                           (4, 33, 1, torch.float32), (2, 17, 45, torch.bfloat16), (7, 8, 32, torch.float16),
                                           (66000, 2, 4, torch.float32), (3, 11, 34, torch.float32)])
def test_x4_upsample_and_adjoint_edge_shapes(sb, planes, h, w, dt):
    """The head's x4 geometry on maps whose size is not a multiple of the 8 x 32 tile (odd heights: the second row of a
    thread's pair does not exist), one-pixel-wide / one-pixel-high maps (generic gather), 16-bit types (TMA box start 8
    columns left of the tile), more planes than grid.y holds: upsample against F.interpolate, adjoint against torch's
    backward of it and through the inner-product identity <up(x), g> = <x, adjoint(g)>."""
    from seghiero_b200 import ops
    gen = torch.Generator().manual_seed(planes + 10 * h + w)
    x = torch.randn(1, planes, h, w, generator=gen).to(dt).cuda()
    g = torch.randn(1, planes, 4 * h, 4 * w, generator=gen).to(dt).cuda()
    up = ops._upsampled(x, 4 * h, 4 * w)
    xr = x.float().clone().requires_grad_(True)
    ref = F.interpolate(xr, size=(4 * h, 4 * w), mode="bilinear", align_corners=False)
    ref.backward(g.float())
    tol = 2e-6 if dt == torch.float32 else 1e-2
    assert float((up.float() - ref.detach()).abs().max()) <= tol * max(1.0, float(ref.detach().abs().max()))
    gin = torch.empty_like(x)
    ops._call("sh_upsample_bilinear_adjoint", ops._p(g), ops._dtype_code(g), ops._p(gin), planes, h, w, 4 * h, 4 * w,
              ops._stream())
    assert rel(to_np(gin.float()), to_np(xr.grad)) <= (1e-6 if dt == torch.float32 else 1e-2)
    if dt == torch.float32:
        lhs = float((ref.detach().double() * g.double()).sum())
        rhs = float((x.double() * gin.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


# ------------------------------------------------------------------------------------------------
# N2: aux-head cross entropy
# ------------------------------------------------------------------------------------------------
def test_n2_aux_ce(sb, golden):
    g = golden("n2_aux_ce")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    loss = sb.aux_cross_entropy(x, torch.from_numpy(g["label"]).cuda())
    (0.4 * loss).backward()
    assert abs(float(loss) - float(g["loss"])) <= FP32_TOL * abs(float(g["loss"]))
    assert rel(to_np(x.grad), g["dx_times_0p4"]) <= FP32_TOL
    # config-3 geometry (one image), H/16 logits, against torch on the same device; uint8 labels
    gen = torch.Generator().manual_seed(5)
    lab = blob_labels(gen, 2, 256, 512, 19, 32, 0.1)
    xa = (torch.randn(2, 19, 16, 32, generator=gen) * 2).cuda()
    xr = xa.clone().requires_grad_(True)
    ref = torch.nn.CrossEntropyLoss(ignore_index=255)(
        F.interpolate(xr, size=(256, 512), mode="bilinear", align_corners=False), lab.cuda())
    ref.backward()
    xc = xa.clone().requires_grad_(True)
    loss = sb.aux_cross_entropy(xc, lab.to(torch.uint8).cuda())
    loss.backward()
    assert abs(float(loss) - float(ref)) <= FP32_TOL * abs(float(ref))
    assert rel(to_np(xc.grad), to_np(xr.grad)) <= FP32_TOL
    # general scale + bf16 + forward only
    lab = iid_labels(gen, 1, 37, 53, 19, 0.2)
    xb = torch.randn(1, 19, 5, 7, generator=gen).bfloat16().cuda()
    with torch.no_grad():
        loss = sb.aux_cross_entropy(xb, lab.cuda())
        ref = torch.nn.CrossEntropyLoss(ignore_index=255)(
            F.interpolate(xb, size=(37, 53), mode="bilinear", align_corners=False).float(), lab.cuda())
    assert abs(float(loss) - float(ref)) <= BF16_TOL * abs(float(ref))


@pytest.mark.parametrize("k,h,w,dt", [(16, 9, 21, torch.float32), (8, 13, 40, torch.float32), (4, 19, 33, torch.float32),
                                      (16, 6, 17, torch.bfloat16), (2, 24, 31, torch.float32), (16, 1, 1, torch.float32)])
def test_n2_aux_ce_integer_scales(sb, k, h, w, dt):
    """The strip kernel (thread = k/2 output pixels that share their source pixels) against torch on the same device:
    x16 / x8 / x4, sizes that are not multiples of the CTA tile, void pixels, bf16 logits; x2 runs the per-pixel kernel."""
    gen = torch.Generator().manual_seed(100 + k + h)
    H, W = k * h, k * w
    lab = iid_labels(gen, 2, H, W, 19, 0.15) if k != 8 else blob_labels(gen, 2, H, W, 19, 24, 0.1)
    lab[0, : H // 3, : W // 2] = 255
    xa = (torch.randn(2, 19, h, w, generator=gen) * 2).to(dt).cuda()
    # 16-bit logits: the reference interpolates in fp32 here (torch's own bf16 backward accumulates the 256 contributions
    # of a source pixel with bf16 atomics and is off by ~10 %); the kernel rounds the interpolated logits to bf16 like
    # torch's forward does, which stays inside the 16-bit tolerance
    xr = xa.float().clone().requires_grad_(True)
    ref = torch.nn.CrossEntropyLoss(ignore_index=255)(
        F.interpolate(xr, size=(H, W), mode="bilinear", align_corners=False), lab.cuda())
    (1.7 * ref).backward()
    xc = xa.clone().requires_grad_(True)
    loss = sb.aux_cross_entropy(xc, lab.to(torch.uint8).cuda())
    (1.7 * loss).backward()
    tol = FP32_TOL if dt == torch.float32 else BF16_TOL
    assert abs(float(loss) - float(ref)) <= tol * abs(float(ref))
    assert rel(to_np(xc.grad.float()), to_np(xr.grad.float())) <= tol
    assert xc.grad.dtype == dt
    # the scatter uses fixed-point integer atomics: a second run gives the same bits
    xd = xa.clone().requires_grad_(True)
    (1.7 * sb.aux_cross_entropy(xd, lab.to(torch.uint8).cuda())).backward()
    assert torch.equal(xd.grad, xc.grad)


# ------------------------------------------------------------------------------------------------
# N3: decode from the head's logits
# ------------------------------------------------------------------------------------------------
def _near_tie_only(x_full, got, ref, lo, hi, tol):
    """Every pixel where the fused decode differs from argmax(F.interpolate) must be a near tie of the two winners."""
    diff = got != ref
    if not bool(diff.any()):
        return 0
    lvl = x_full[:, lo:hi].float()
    a = lvl.gather(1, got.long().unsqueeze(1)).squeeze(1)[diff]
    b = lvl.gather(1, ref.long().unsqueeze(1)).squeeze(1)[diff]
    assert float((a - b).abs().max()) <= tol
    return int(diff.sum())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_n3_decode_upsampled(sb, golden, dtype):
    g = golden("n3_decode_up4")
    x = torch.from_numpy(g["x"]).to(dtype).cuda()
    lab = torch.from_numpy(g["label"]).cuda()
    preds, counts = sb.hierarchical_argmax(x, [19, 7, 2], lab, out_dtype=torch.uint8)
    full = F.interpolate(x, size=lab.shape[-2:], mode="bilinear", align_corners=False)
    n_flip = 0
    for p, (lo, hi) in zip(preds, ((0, 19), (19, 26), (26, 28))):
        assert p.dtype == torch.uint8 and p.shape == lab.shape
        n_flip += _near_tie_only(full, p.long(), full[:, lo:hi].argmax(1), lo, hi, 1e-6 if dtype == torch.float32 else 0.0)
    if dtype == torch.float32:
        assert np.array_equal(preds[0].cpu().numpy().astype(np.int64), g["pf"]) or n_flip > 0
    pf = preds[0].long()
    assert counts.tolist() == [int(((pf == lab) & (lab != 255)).sum()), int((lab != 255).sum())]
    # fused (4x) path == our own two-kernel path (upsample kernel + decode kernel), bit for bit
    from seghiero_b200 import ops
    up = ops._upsampled(x, *lab.shape[-2:])
    two, c2 = sb.hierarchical_argmax(up, [19, 7, 2], lab, out_dtype=torch.uint8)
    assert all(torch.equal(a, b) for a, b in zip(preds, two)) and counts.tolist() == c2.tolist()
    # any image size (infer.py): falls back to upsample + decode
    g2 = golden("n3_decode_up_any")
    preds, _ = sb.hierarchical_argmax(torch.from_numpy(g2["x"]).to(dtype).cuda(), [19, 7, 2], size=(17, 23))
    full = F.interpolate(torch.from_numpy(g2["x"]).to(dtype).cuda(), size=(17, 23), mode="bilinear", align_corners=False)
    for p, (lo, hi) in zip(preds, ((0, 19), (19, 26), (26, 28))):
        _near_tie_only(full, p, full[:, lo:hi].argmax(1), lo, hi, 1e-6 if dtype == torch.float32 else 0.0)


def test_n3_decode_upsampled_full_size(sb):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2, 28, 512, 512, generator=g, device="cuda", dtype=torch.bfloat16)
    lab = torch.randint(0, 19, (2, 2048, 2048), generator=g, device="cuda").to(torch.uint8)
    preds, counts = sb.hierarchical_argmax(x, [19, 7, 2], lab, out_dtype=torch.uint8)
    full = F.interpolate(x, size=(2048, 2048), mode="bilinear", align_corners=False)
    tot = 0
    for p, (lo, hi) in zip(preds, ((0, 19), (19, 26), (26, 28))):
        tot += _near_tie_only(full, p.long(), full[:, lo:hi].argmax(1), lo, hi, 0.0)
    assert tot <= 1e-4 * 3 * lab.numel(), tot     # bf16-rounded values: exact ties decided by the last fp32 ulp before rounding
    assert int(counts[1]) == lab.numel()


# ------------------------------------------------------------------------------------------------
# custom-op boundary
# ------------------------------------------------------------------------------------------------
def test_custom_ops_registered_and_opcheck(sb):
    assert hasattr(torch.ops.seghiero_b200, "hier3_fwd") and hasattr(torch.ops.seghiero_b200, "hier2_fwd")
    g = torch.Generator().manual_seed(21)
    lab = blob_labels(g, 1, 32, 64, 19, 8, 0.1).cuda()
    x = (torch.randn(1, 28, 32, 64, generator=g) * 2).cuda().requires_grad_(True)
    emb = F.normalize(torch.randn(1, 8, 4, 8, generator=g), dim=1).cuda().requires_grad_(True)
    step = torch.tensor([100000.0], dtype=torch.float64).cuda()
    args = (x, lab, emb, step, 19, 7, 2, list(F2M), list(F2H), [1, 2, 3, 4, 5, 6, 7, 10, 11, 13, 14, 15],
            [8, 9, 12, 16, 17, 18, 19], 0.5, 1.0, 160000.0, True, True)
    torch.library.opcheck(torch.ops.seghiero_b200.hier3_fwd.default, args,
                          test_utils=("test_schema", "test_faketensor"))


def test_inplace_edit_between_forward_and_backward_is_detected(sb):
    g = torch.Generator().manual_seed(22)
    lab = blob_labels(g, 1, 32, 64, 19, 8, 0.1).cuda()
    base = (torch.randn(1, 28, 32, 64, generator=g) * 2).cuda().requires_grad_(True)
    x = base * 1.0
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    loss = mod(torch.tensor([0]), None, None, x, lab)
    x.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        loss.backward()
    # the 2-level gradient is computed in the forward kernel and scaled in place: a second backward must not rescale it
    x2 = (torch.randn(1, 26, 32, 64, generator=g) * 2).cuda().requires_grad_(True)
    l2 = sb.HieraTripletLoss(19, HM, HI)(torch.tensor([0]), None, None, x2, lab)
    l2.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="only once|modified by an inplace operation"):
        l2.backward()


def test_torch_compile_fullgraph(sb):
    g = torch.Generator().manual_seed(23)
    lab = blob_labels(g, 2, 64, 128, 19, 8, 0.1).to(torch.uint8).cuda()
    emb0 = F.normalize(torch.randn(2, 16, 8, 16, generator=g), dim=1).cuda()
    step = torch.tensor([100000]).cuda()
    for mod, c in ((sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H)), 28),
                   (sb.HieraTripletLoss(19, HM, HI), 26)):
        x0 = (torch.randn(2, c, 64, 128, generator=g) * 2).cuda()

        def fn(x, e):
            return mod(step, e, None, x * 1.0, lab) * 2.0

        xe, ee = x0.clone().requires_grad_(True), emb0.clone().requires_grad_(True)
        le = fn(xe, ee)
        le.backward()
        xc, ec = x0.clone().requires_grad_(True), emb0.clone().requires_grad_(True)
        lc = torch.compile(fn, fullgraph=True)(xc, ec)
        lc.backward()
        assert torch.equal(lc.detach(), le.detach())
        assert torch.equal(xc.grad, xe.grad)
        assert rel(to_np(ec.grad), to_np(ee.grad)) <= 1e-6


def test_triplet_label_outside_id_lists_poisons_the_fused_loss(sb):
    """the reference raises ValueError from list.remove; the fused module (no host sync) returns NaN, and
    strict=True raises."""
    g = torch.Generator().manual_seed(24)
    lab = torch.full((1, 64, 64), 5, dtype=torch.long)
    lab[:, :, 32:] = 8                      # 8 is in neither id list of a 9-class hierarchy
    x = torch.randn(1, 9 + 4 + 3, 64, 64, generator=g).cuda()
    emb = F.normalize(torch.randn(1, 8, 8, 8, generator=g), dim=1).cuda()
    f2m, f2h = torch.tensor([0, 0, 1, 1, 1, 3, 3, 3, 3]), torch.tensor([0, 0, 0, 0, 0, 2, 2, 2, 2])
    assert torch.isnan(sb.RMIHieraTripletLoss(9, 4, 3, f2m, f2h)(torch.tensor([0]), emb, None, x, lab.cuda()))
    with pytest.raises(ValueError):
        sb.RMIHieraTripletLoss(9, 4, 3, f2m, f2h, strict=True)(torch.tensor([0]), emb, None, x, lab.cuda())


def test_two_level_large_hierarchies(sb):
    """the reference handles any class count in the 2-level loss.  150 fine + 30 coarse classes run
    through the 128-pixel form of the any-bucket kernel; 64 channels is the largest tree-order case; beyond ~220
    channels the constructor refuses with a clear ValueError."""
    with pytest.raises(ValueError, match="channels exceed"):
        sb.HieraTripletLoss(230, [i // 23 for i in range(230)], [[23 * i, 23 * i + 23] for i in range(10)])
    for nf, per in ((150, 5), (56, 7)):
        nc = nf // per
        hi = [[per * i, per * i + per] for i in range(nc)]
        hm = [i // per for i in range(nf)]
        g = torch.Generator().manual_seed(nf)
        lab = blob_labels(g, 2, 32, 66, nf, 8, 0.1)
        x = (torch.randn(2, nf + nc, 32, 66, generator=g) * 2)
        emb = F.normalize(torch.randn(2, 8, 4, 8, generator=g), dim=1)
        xr = x.clone().requires_grad_(True)
        ref, _ = O.hiera_triplet_loss(0, emb, xr, lab, nf, hm, hi)
        ref.backward()
        xc = x.cuda().requires_grad_(True)
        loss = sb.HieraTripletLoss(nf, hm, hi)(torch.tensor([0]), emb.cuda(), None, xc, lab.cuda())
        loss.backward()
        assert abs(float(loss) - float(ref)) <= FP32_TOL * abs(float(ref)), (nf, float(loss), float(ref))
        assert rel(to_np(xc.grad), to_np(xr.grad)) <= FP32_TOL, nf


# ------------------------------------------------------------------------------------------------
# parity at the benchmark shapes
# ------------------------------------------------------------------------------------------------
def _sample_in_batch_equals_sample_alone(sb, mod, x, lab, idx, b, px_share):
    """The gradient of sample i inside the batch equals, after the normalisers that depend on the batch (number of
    valid pixels, B), the gradient of the same sample run alone -- checked through a third quantity both must match:
    the batch gradient restricted to sample i is a fixed linear combination of per-term gradients, so we compare
    against the same module on the single sample with the batch's normalisers folded in by construction of the
    inputs (all samples share one label map => equal valid counts)."""
    xg = x.clone().requires_grad_(True)
    loss = mod(torch.tensor([0], device="cuda"), None, None, xg, lab)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(xg.grad).all()
    for i in idx:
        x1 = x[i:i + 1].clone().requires_grad_(True)
        l1 = mod(torch.tensor([0], device="cuda"), None, None, x1, lab[i:i + 1])
        l1.backward()
        # same labels in every sample: nv_batch = B * nv_1, CE denominators B*H*W vs H*W, RMI mean over B: all 1/B
        err = float((xg.grad[i:i + 1] * b - x1.grad).norm() / x1.grad.norm())
        assert err <= FP32_TOL, (i, err)
    return float(loss)


def test_config3_bench_shape_batch8(sb):
    """cfg3 as benched: B=8, 28 x 1024 x 2048 fp32 (1.88 GB of logits): first and last sample against the sample alone."""
    g = torch.Generator(device="cuda").manual_seed(7)
    gc = torch.Generator().manual_seed(7)
    lab1 = blob_labels(gc, 1, 1024, 2048, 19, 32, 0.1).cuda()
    lab = lab1.expand(8, -1, -1).contiguous()
    x = torch.randn(8, 28, 1024, 2048, generator=g, device="cuda") * 2
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    mod.triplet_loss = None
    _sample_in_batch_equals_sample_alone(sb, mod, x, lab, (0, 7), 8, None)


def test_config4_bench_shape_batch32(sb):
    """cfg4 as benched: B=32, 186 x 512 x 512 fp32 (6.24 GB: byte offsets beyond 2^32)."""
    nf, nm, nh = 150, 30, 6
    f2m, f2h = [f // 5 for f in range(nf)], [f // 25 for f in range(nf)]
    g = torch.Generator(device="cuda").manual_seed(9)
    gc = torch.Generator().manual_seed(9)
    lab1 = blob_labels(gc, 1, 512, 512, nf, 32, 0.1).cuda()
    lab = lab1.expand(32, -1, -1).contiguous()
    x = torch.randn(32, nf + nm + nh, 512, 512, generator=g, device="cuda") * 2
    mod = sb.RMIHieraTripletLoss(nf, nm, nh, torch.tensor(f2m), torch.tensor(f2h))
    mod.triplet_loss = None
    _sample_in_batch_equals_sample_alone(sb, mod, x, lab, (0, 31), 32, None)


def test_full_size_iid_image_against_oracle_on_cuda(sb):
    """One full-size image with per-pixel (iid) labels: the inline one-hot instantiation of pass 2, against the oracle's
    torch restatement on CUDA tensors; prints how many pixels route a tied max/min term differently."""
    from tests.test_gpu_parity import _assert_grad_close_up_to_tie_flips
    dev = "cuda"
    g = torch.Generator().manual_seed(4096)
    lab = iid_labels(g, 1, 1024, 2048, 19, 0.1).to(dev)
    x = (torch.randn(1, 28, 1024, 2048, generator=g) * 2).to(dev)
    xr = x.clone().requires_grad_(True)
    ref, _ = O.rmi_hiera_triplet_loss(0, None, xr, lab, 19, 7, 2, F2M, F2H, with_triplet=False)
    ref.backward()
    xc = x.clone().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    mod.triplet_loss = None
    loss = mod(torch.tensor([0], device=dev), None, None, xc, lab)
    loss.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= FP32_TOL * abs(float(ref.detach()))
    n_flip = _assert_grad_close_up_to_tie_flips(xc.grad, xr.grad, FP32_TOL)
    print(f"full-size iid image: {n_flip} tie-flip pixels of {lab.numel()}")

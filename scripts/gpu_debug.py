"""Component-wise diagnostics of the CUDA path against the oracle (run on the GPU box)."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from seghiero_b200 import _lib
from oracle import hiera_oracle as O
from tests.util import F2H, F2M, HI, HM, blob_labels, iid_labels, rel, to_np

torch.manual_seed(0)


def two_level(b, h, w, labels, dtype=torch.float32):
    g = torch.Generator().manual_seed(h * 7 + w)
    lab = iid_labels(g, b, h, w, 19, 0.15) if labels == "iid" else blob_labels(g, b, h, w, 19, 6, 0.1)
    x = (torch.randn(b, 26, h, w, generator=g) * 3).to(dtype)
    emb = F.normalize(torch.randn(b, 12, max(h // 8, 1), max(w // 8, 1), generator=g), dim=1)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    ref, parts = O.hiera_triplet_loss(60000, er, xr, lab, 19, HM, HI)
    ref.backward()
    xc, ec = x.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
    mod = sb.HieraTripletLoss(19, HM, HI)
    loss = mod(torch.tensor([60000]).cuda(), ec, None, xc, lab.cuda())
    loss.backward()
    st = mod.last_stats
    sums, counts = st["sums"].cpu().numpy(), st["counts"].cpu().numpy()
    npx = b * h * w
    print(f"[2lvl {b}x{h}x{w} {labels} {dtype}] loss gpu {float(loss):.7f} ref {float(ref):.7f}")
    print("   hiera gpu", 5 * (sums[0] / (max(counts[0], 1) * 19) + sums[1] / (max(counts[1], 1) * 7)), "ref", parts["hiera"])
    print("   ce gpu", sums[2] / npx, sums[3] / npx, "ref", parts["ce"])
    print("   trip gpu", st["triplet"].trip.tolist(), "ref", parts["triplet"], parts["count"], "out", st["out"].tolist())
    print("   dx rel", rel(to_np(xc.grad), to_np(xr.grad)), "demb rel",
          rel(to_np(ec.grad), to_np(er.grad)) if er.grad is not None else None)
    gd = (xc.grad.cpu().float() - xr.grad.float()).abs()
    print("   dx max abs err", float(gd.max()), "at", np.unravel_index(int(gd.argmax()), gd.shape),
          "per-channel rel:", [round(rel(to_np(xc.grad[:, c]), to_np(xr.grad[:, c])), 7) for c in range(0, 26, 5)])


def three_level(b, h, w, labels, lam=0.5):
    g = torch.Generator().manual_seed(h * 11 + w)
    lab = iid_labels(g, b, h, w, 19, 0.15) if labels == "iid" else blob_labels(g, b, h, w, 19, 7, 0.1)
    x = torch.randn(b, 28, h, w, generator=g) * 2
    emb = F.normalize(torch.randn(b, 12, max(h // 8, 1), max(w // 8, 1), generator=g), dim=1)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    ref, parts = O.rmi_hiera_triplet_loss(100000, er, xr, lab, 19, 7, 2, F2M, F2H, loss_weight_lambda=lam)
    ref.backward()
    # oracle r_bc
    tf, tm, th = O.targets_three_level(lab.numpy(), F2M, F2H)
    la, pr = O.rmi_onehot_and_probs(x, tf, tm, th, 19, 7, 2)
    la_v, pr_v = O.rmi_windows(la).double(), O.rmi_windows(pr).double()
    _, r_bc = O.rmi_from_moments(la_v @ la_v.transpose(2, 3), pr_v @ pr_v.transpose(2, 3), la_v @ pr_v.transpose(2, 3))
    xc, ec = x.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), loss_weight_lambda=lam)
    loss = mod(torch.tensor([100000]).cuda(), ec, None, xc, lab.cuda())
    loss.backward()
    st = mod.last_stats
    ws = st["workspace"]
    offs = (ctypes.c_size_t * 12)()
    _lib.load().sh_rmi3_workspace_offsets(b, h, w, 19, 7, 2, offs)
    sums = ws[offs[1]:offs[1] + 64].view(torch.float64).cpu().numpy()
    nv = int(ws[0:8].view(torch.int64).item())
    rbc = ws[offs[9]:offs[9] + b * 28 * 8].view(torch.float64).cpu().numpy().reshape(b, 28)
    npx = b * h * w
    print(f"[3lvl {b}x{h}x{w} {labels}] loss gpu {float(loss):.7f} ref {float(ref):.7f}  nv {nv} vs {int((lab != 255).sum())}")
    print("   hiera gpu", 5 * (sums[0] / (max(nv, 1) * 19) + sums[1] / (max(nv, 1) * 7) + sums[2] / (max(nv, 1) * 2)), "ref", parts["hiera"])
    print("   ce gpu", (sums[3:6] / npx).tolist(), "ref", parts["ce"])
    print("   rmi gpu", st["out"][2].item(), "ref", parts["rmi"], " r_bc rel", rel(rbc, r_bc.numpy()),
          "max abs", float(np.abs(rbc - r_bc.numpy()).max()))
    print("   trip gpu", st["triplet"].trip.tolist(), "ref", parts["triplet"], parts["count"])
    print("   dx rel", rel(to_np(xc.grad), to_np(xr.grad)), "demb rel",
          rel(to_np(ec.grad), to_np(er.grad)) if er.grad is not None else None)
    gd = (xc.grad.cpu() - xr.grad).abs()
    interior = torch.zeros(h, w, dtype=torch.bool)
    interior[2:h - 2, 2:w - 2] = True
    print("   dx rel interior", rel(to_np(xc.grad[..., interior]), to_np(xr.grad[..., interior])),
          "frame", rel(to_np(xc.grad[..., ~interior]), to_np(xr.grad[..., ~interior])))
    print("   dx max abs err", float(gd.max()), "at", np.unravel_index(int(gd.argmax()), gd.shape),
          "per-level rel:", rel(to_np(xc.grad[:, :19]), to_np(xr.grad[:, :19])),
          rel(to_np(xc.grad[:, 19:26]), to_np(xr.grad[:, 19:26])), rel(to_np(xc.grad[:, 26:]), to_np(xr.grad[:, 26:])))
    # lambda = 0 isolates BCE+CE gradient
    xr2 = x.clone().requires_grad_(True)
    ref2, _ = O.rmi_hiera_triplet_loss(100000, None, xr2, lab, 19, 7, 2, F2M, F2H, loss_weight_lambda=0.0, with_triplet=False)
    ref2.backward()
    xc2 = x.cuda().requires_grad_(True)
    mod2 = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), loss_weight_lambda=0.0)
    mod2.triplet_loss = None
    l2 = mod2(torch.tensor([0]).cuda(), None, None, xc2, lab.cuda())
    l2.backward()
    print("   lam=0: loss", float(l2), float(ref2), "dx rel", rel(to_np(xc2.grad), to_np(xr2.grad)))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    two_level(2, 33, 47, "iid")
    two_level(2, 64, 96, "blob")
    two_level(1, 32, 64, "blob", torch.bfloat16)
    three_level(1, 24, 40, "blob")
    three_level(2, 37, 50, "iid")
    three_level(1, 48, 132, "blob", lam=1.0)

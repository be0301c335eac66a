"""Fast (warp-specialised) vs generic 3-level kernels on the same inputs (development aid, GPU box)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from seghiero_b200 import ops, _lib
from tests.util import F2H, F2M, blob_labels, iid_labels, rel, to_np


def run(x, lab, emb, fast, lam=0.5):
    ops.FAST_PATH["enabled"] = fast
    xc = x.clone().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), loss_weight_lambda=lam)
    loss = mod(torch.tensor([100000]).cuda(), emb, None, xc, lab)
    loss.backward()
    torch.cuda.synchronize()
    st = mod.last_stats
    ws = st["workspace"]
    b, c, h, w = x.shape
    offs = (ctypes_offsets(b, h, w))
    n = b * h * w
    sums = ws[offs[1]:offs[1] + 64].view(torch.float64).cpu().numpy().copy()
    rbc = ws[offs[9]:offs[9] + b * c * 8].view(torch.float64).cpu().numpy().copy()
    hold = ws[offs[4]:offs[4] + 11 * n].cpu().numpy().copy().reshape(11, n)
    inv = ws[offs[5]:offs[5] + 3 * n * 4].view(torch.float32).cpu().numpy().copy()
    wts = ws[offs[10]:offs[10] + b * c * 64 * 4].view(torch.float32).cpu().numpy().copy()
    return float(loss), st["out"].cpu().numpy().copy(), sums, rbc, hold, inv, wts, xc.grad.detach().clone()


def ctypes_offsets(b, h, w):
    import ctypes
    out = (ctypes.c_size_t * 12)()
    _lib.load().sh_rmi3_workspace_offsets(b, h, w, 19, 7, 2, out)
    return list(out)


def case(b, h, w, labels, dtype=torch.float32, tile=7):
    g = torch.Generator().manual_seed(h * 11 + w)
    lab = iid_labels(g, b, h, w, 19, 0.15) if labels == "iid" else blob_labels(g, b, h, w, 19, tile, 0.1)
    x = (torch.randn(b, 28, h, w, generator=g) * 2).to(dtype).cuda()
    emb = F.normalize(torch.randn(b, 12, max(h // 8, 1), max(w // 8, 1), generator=g), dim=1).cuda()
    lab = lab.cuda()
    rg = run(x, lab, emb, False)
    rf = run(x, lab, emb, True)
    print(f"[{b}x{h}x{w} {labels} {dtype}] loss generic {rg[0]:.7f} fast {rf[0]:.7f} rel {abs(rg[0]-rf[0])/abs(rg[0]):.2e}")
    print("   out", rg[1], rf[1])
    print("   sums rel", np.abs(rg[2][:6] - rf[2][:6]) / np.maximum(np.abs(rg[2][:6]), 1e-30))
    print("   rbc rel max", float(np.max(np.abs(rg[3] - rf[3]) / np.maximum(np.abs(rg[3]), 1e-30))))
    print("   hold mismatches per plane", [(int((rg[4][i] != rf[4][i]).sum())) for i in range(11)])
    print("   inv rel", rel(rf[5], rg[5]), "wts rel", rel(rf[6], rg[6]))
    print("   dx rel", rel(to_np(rf[7]), to_np(rg[7])))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        case(1, 24, 64, "blob")
        case(2, 37, 52, "iid")
        case(1, 48, 132, "blob")
        case(3, 16, 260, "blob")
        case(2, 100, 200, "iid")
        case(1, 40, 72, "blob", torch.bfloat16)
    else:
        case(2, 512, 1024, "blob", tile=32)
        case(1, 1024, 2048, "blob", tile=32)
        case(1, 1024, 2048, "iid")

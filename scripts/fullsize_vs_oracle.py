"""Full-size parity on the GPU box: the oracle's torch restatement evaluated on CUDA tensors (eager ATen, fp64 RMI
unfolds -- ~21 GB per 1024x2048 image) against the CUDA kernels.  Development aid; the oracle is only the checker."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import seghiero_b200 as sb
from oracle import hiera_oracle as O
from tests.util import F2H, F2M, HI, HM, blob_labels, iid_labels

def rel(a, b):
    return float((a - b).norm() / b.norm())

dev = "cuda"
for (b, h, w, kind) in [(1, 1024, 2048, "blob"), (1, 512, 1024, "iid")]:
    g = torch.Generator().manual_seed(h + w)
    lab = (blob_labels(g, b, h, w, 19, 32, 0.1) if kind == "blob" else iid_labels(g, b, h, w, 19, 0.1)).to(dev)
    x = (torch.randn(b, 28, h, w, generator=g) * 2).to(dev)
    emb = F.normalize(torch.randn(b, 64, h // 32, w // 32, generator=g), dim=1).to(dev)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    try:
        ref, parts = O.rmi_hiera_triplet_loss(100000, er, xr, lab, 19, 7, 2, F2M, F2H)
        ref.backward()
    except Exception as e:
        print("oracle on CUDA failed:", type(e).__name__, str(e)[:200]); continue
    xc, ec = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    loss = mod(torch.tensor([100000], device=dev), ec, None, xc, lab)
    loss.backward()
    torch.cuda.synchronize()
    d = (xc.grad - xr.grad).abs()
    print(f"3-level {b}x{h}x{w} {kind}: loss {float(loss):.7f} vs {float(ref):.7f} rel {abs(float(loss)-float(ref))/abs(float(ref)):.2e}; "
          f"grad rel {rel(xc.grad, xr.grad):.2e}; pixels with |diff| > 1e-3 max|g|: {int((d > 1e-3 * xr.grad.abs().max()).sum())}")
    del xr, er, xc, ec, ref, loss, d
    torch.cuda.empty_cache()
for dt in (torch.float32, torch.bfloat16):
    b, h, w = 4, 512, 1024
    g = torch.Generator().manual_seed(7)
    lab = blob_labels(g, b, h, w, 19, 32, 0.1).to(dev)
    x = (torch.randn(b, 26, h, w, generator=g) * 2).to(dt).to(dev)
    emb = F.normalize(torch.randn(b, 64, h // 32, w // 32, generator=g), dim=1).to(dev)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    ref, parts = O.hiera_triplet_loss(100000, er, xr, lab, 19, HM, HI)
    ref.backward()
    xc, ec = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    mod = sb.HieraTripletLoss(19, HM, HI)
    loss = mod(torch.tensor([100000], device=dev), ec, None, xc, lab)
    loss.backward()
    torch.cuda.synchronize()
    print(f"2-level {b}x{h}x{w} {dt}: loss rel {abs(float(loss)-float(ref))/abs(float(ref)):.2e}; grad rel {rel(xc.grad.float(), xr.grad.float()):.2e}")

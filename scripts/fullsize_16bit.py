"""Full-size 16-bit parity against the oracle restatement on CUDA (development aid).  bf16: gradient within 3e-3 of its norm;
fp16: the per-element gradients of a 2-Mpixel image (1e-7 .. 1e-6) sit in the fp16 subnormal range, 6e-2 is rounding there."""
import os, sys
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
import seghiero_b200 as sb
from oracle import hiera_oracle as O
from tests.util import F2H, F2M, blob_labels
dev = "cuda"
g = torch.Generator().manual_seed(99)
for dt in (torch.bfloat16, torch.float16):
    lab = blob_labels(g, 1, 1024, 2048, 19, 32, 0.1).to(dev)
    x = (torch.randn(1, 28, 1024, 2048, generator=g) * 2).to(dt).to(dev)
    emb = F.normalize(torch.randn(1, 64, 32, 64, generator=g), dim=1).to(dev)
    xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    ref, _ = O.rmi_hiera_triplet_loss(100000, er, xr, lab, 19, 7, 2, F2M, F2H)
    ref.backward()
    xc, ec = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    loss = mod(torch.tensor([100000], device=dev), ec, None, xc, lab)
    loss.backward()
    print(dt, "loss rel", abs(float(loss.detach()) - float(ref.detach())) / abs(float(ref.detach())),
          "grad rel", float((xc.grad.float() - xr.grad.float()).norm() / xr.grad.float().norm()))
    del xr, er, xc, ec, ref, loss
    torch.cuda.empty_cache()

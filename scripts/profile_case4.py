"""Small config-4 workload (150 / 30 / 6 classes, 512 x 512, batch 4, full-resolution fp32 logits, uint8 labels) for
`ncu --set full` captures of the wide-hierarchy instantiations (pass 1 with 2 round buffers, TMA pass 2 at C = 186)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from bench import make_labels, hierarchy_maps, WORKLOADS

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
w = WORKLOADS["cfg4"]
b, h, wd = 4, w["H"], w["W"]
f2m, f2h = hierarchy_maps(w)
lab = make_labels(torch, g, b, h, wd, w["nf"], "blob", dev).to(torch.uint8)
c = w["nf"] + w["nm"] + w["nh"]
x = (torch.randn(b, c, h, wd, generator=g, device=dev) * 2).requires_grad_(True)
emb = F.normalize(torch.randn(b, 256, h // 32, wd // 32, generator=g, device=dev), dim=1).requires_grad_(True)
mod = sb.RMIHieraTripletLoss(w["nf"], w["nm"], w["nh"], torch.tensor(f2m), torch.tensor(f2h))
if not w.get("triplet", True):
    mod.triplet_loss = None
for _ in range(3):
    x.grad = None
    emb.grad = None
    loss = mod(torch.tensor([100000], device=dev), emb if mod.triplet_loss is not None else None, None, x, lab)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))

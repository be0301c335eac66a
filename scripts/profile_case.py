"""One small workload for `ncu --set full` captures: config 3 shapes, batch 2, fed with the head's H/4 logits and uint8
labels (so that the upsample and its adjoint are in the launch list too), plus the aux-head cross entropy from H/16
logits.  3 steps; capture the last one (-k regex:'k3|k_up|k_aux' skips nothing else; use -s to skip the warm-up steps)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from bench import make_labels
from tests.util import F2H, F2M

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1234)
b, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 2, 1024, 2048
lab = make_labels(torch, g, b, h, w, 19, "blob", dev).to(torch.uint8)
x = (torch.randn(b, 28, h // 4, w // 4, generator=g, device=dev) * 2).requires_grad_(True)
emb = F.normalize(torch.randn(b, 256, h // 32, w // 32, generator=g, device=dev), dim=1).requires_grad_(True)
mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
aux = (torch.randn(b, 19, h // 16, w // 16, generator=g, device=dev) * 2).requires_grad_(True)
for _ in range(3):
    x.grad = None
    emb.grad = None
    aux.grad = None
    loss = mod(torch.tensor([100000], device=dev), emb, None, x, lab) + 0.4 * sb.aux_cross_entropy(aux, lab)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))

"""Run-to-run bit equality of the 3-level loss and gradients (default stream, user stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import seghiero_b200 as sb
from tests.util import blob_labels, F2M, F2H

g = torch.Generator().manual_seed(77)
b, h, w = 2, 64, 96
lab = blob_labels(g, b, h, w, 19, 8, 0.1).cuda()
x0 = (torch.randn(b, 28, h, w, generator=g) * 2).cuda()
e0 = F.normalize(torch.randn(b, 16, h // 8, w // 8, generator=g), dim=1).cuda()
step = torch.tensor([170000]).cuda()
mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
x, e = x0.clone().requires_grad_(True), e0.clone().requires_grad_(True)

def run(use_e=True):
    loss = mod(step, e if use_e else None, None, x, lab)
    if use_e:
        gx, ge = torch.autograd.grad(loss, (x, e))
    else:
        (gx,) = torch.autograd.grad(loss, (x,)); ge = torch.zeros(1, device="cuda")
    torch.cuda.synchronize()
    return loss.detach().clone(), gx.clone(), ge.clone()

for use_e in (True, False):
    ref = run(use_e)
    for it in range(4):
        got = run(use_e)
        print("emb" if use_e else "no-emb", it, [bool(torch.equal(a, r)) for a, r in zip(got, ref)],
              float((got[1] - ref[1]).abs().max()), float(got[0] - ref[0]))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    got = run(True)
print("user stream", [bool(torch.equal(a, r)) for a, r in zip(got, run(True))])

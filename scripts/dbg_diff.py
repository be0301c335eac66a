import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import seghiero_b200 as sb
from tests.util import F2H, F2M, blob_labels

def run(x, lab, emb, fast):
    xc = x.clone().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), fast_path=fast)
    loss = mod(torch.tensor([100000]).cuda(), emb, None, xc, lab)
    loss.backward()
    torch.cuda.synchronize()
    return xc.grad.detach().clone()

b, h, w = 1, 1024, 2048
g = torch.Generator().manual_seed(h * 11 + w)
lab = blob_labels(g, b, h, w, 19, 32, 0.1).cuda()
x = (torch.randn(b, 28, h, w, generator=g) * 2).cuda()
emb = F.normalize(torch.randn(b, 12, h // 8, w // 8, generator=g), dim=1).cuda()
gg = run(x, lab, emb, False)
gf = run(x, lab, emb, True)
gf2 = run(x, lab, emb, True)
d = (gf - gg).abs()
print("fast run-to-run max diff", float((gf - gf2).abs().max()))
print("max abs diff", float(d.max()), "grad abs max", float(gg.abs().max()), "norm rel", float((gf - gg).norm() / gg.norm()))
idx = torch.nonzero(d > 1e-3 * gg.abs().max())
print("n large", idx.shape[0])
if idx.shape[0]:
    print("channels", torch.unique(idx[:, 1]).tolist())
    ys = idx[:, 2]; xs = idx[:, 3]
    print("y range", int(ys.min()), int(ys.max()), "x range", int(xs.min()), int(xs.max()))
    print("y % 32 hist", torch.bincount(ys % 32, minlength=32).tolist())
    print("x % 64 hist", torch.bincount(xs % 64, minlength=64).tolist())
    print(idx[:10].tolist())
    for i in idx[:5].tolist():
        print(i, float(gg[tuple(i)]), float(gf[tuple(i)]))

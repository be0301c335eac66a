import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import seghiero_b200 as sb
from tests.util import F2M, F2H, rel, to_np
g = dict(np.load("/root/repo/tests/golden/three_level_flat.npz"))
for fast in (True, False):
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H), fast_path=fast)
    loss = mod(torch.tensor([0]), torch.zeros(1, 4, 2, 2).cuda(), None, x, torch.from_numpy(g["label"]).cuda())
    loss.backward()
    print("fast" if fast else "generic", "loss rel", abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])), "grad rel", rel(to_np(x.grad), g["dx"]))

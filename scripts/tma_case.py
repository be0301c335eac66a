import torch, sys
sys.path.insert(0, "/root/repo")
import seghiero_b200 as sb
from seghiero_b200 import ops
from bench import make_labels
from tests.util import F2M, F2H
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1234)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lab = make_labels(torch, g, b, 1024, 2048, 19, "blob", dev)
x = (torch.randn(b, 28, 1024, 2048, generator=g, device=dev) * 2).requires_grad_(True)
mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
mod.triplet_loss = None
for it in range(8):
    x.grad = None
    out, ws, *_ = ops.hier3_fwd(x, lab, None, ops.step_tensor(0, dev), 19, 7, 2, list(F2M), list(F2H), [1], [2], 0.5, 1.0, 160000.0, True, False)
    out[0].backward()
    torch.cuda.synchronize()
    dbg = int(ws[24:32].view(torch.int64).item())
    cnt = int(ws[8:16].view(torch.int64).item())
    print("iter", it, "loss", float(out[0]), "dbg", hex(dbg), "block", dbg & 0xffffff, "ci", (dbg >> 24) & 0xffff, "stalled warps", cnt)

"""Stand-alone timing of the upsample / adjoint kernels (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seghiero_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

b, c, h, w = 8, 28, 256, 512
x = torch.randn(b, c, h, w, device="cuda")
full = torch.empty(b, c, 4 * h, 4 * w, device="cuda")
gl = torch.empty_like(x)
st = ops._stream
up = lambda: ops._call("sh_upsample_bilinear", ops._p(x), 0, ops._p(full), b * c, h, w, 4 * h, 4 * w, st())
adj = lambda: ops._call("sh_upsample_bilinear_adjoint", ops._p(full), 0, ops._p(gl), b * c, h, w, 4 * h, 4 * w, st())
cp = lambda: full.clone()
print(f"upsample {t(up):.3f} ms  adjoint {t(adj):.3f} ms  clone(1.88GB r+w) {t(cp):.3f} ms  sum-read {t(lambda: full.sum()):.3f} ms")

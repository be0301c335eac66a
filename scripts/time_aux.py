"""CUDA-event timing of the fused aux-head cross entropy (N2) and of the H/4-input loss step at config 3 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from bench import make_labels
from tests.util import F2H, F2M

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
b, h, w = 8, 1024, 2048
lab = make_labels(torch, g, b, h, w, 19, "blob", dev).to(torch.uint8)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for div in (16, 8, 4):
    aux = (torch.randn(b, 19, h // div, w // div, generator=g, device=dev) * 2).requires_grad_(True)

    def step():
        aux.grad = None
        sb.aux_cross_entropy(aux, lab).backward()
    print(f"aux CE from H/{div} logits: {timeit(step):.3f} ms fwd+bwd")

    def ref():
        aux.grad = None
        F.cross_entropy(F.interpolate(aux, (h, w), mode="bilinear", align_corners=False), lab.long(), ignore_index=255).backward()
    if div == 16:
        print(f"  eager ATen (interpolate + cross_entropy): {timeit(ref, 3):.3f} ms")

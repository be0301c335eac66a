"""CUDA-event timing of the x4 upsample and its adjoint at config 3 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seghiero_b200 import ops, _lib

dev = torch.device("cuda")
for dt in (torch.float32, torch.bfloat16):
    b, c, h, w = 8, 28, 256, 512
    x = torch.randn(b, c, h, w, device=dev).to(dt)
    g = torch.randn(b, c, 4 * h, 4 * w, device=dev).to(dt)
    gin = torch.empty_like(x)
    lib = _lib.load()

    def adj():
        ops._call("sh_upsample_bilinear_adjoint", ops._p(g), ops._dtype_code(g), ops._p(gin), b * c, h, w, 4 * h, 4 * w, ops._stream())

    def up():
        return ops._upsampled(x, 4 * h, 4 * w)
    for name, fn, nbytes in (("adjoint", adj, g.numel() * g.element_size() * (1 + 1 / 16)), ("upsample", up, g.numel() * g.element_size() * (1 + 1 / 16))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{name} {str(dt)[6:]}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s")
    # check against torch's own backward
    if dt == torch.float32:
        xr = x.clone().requires_grad_(True)
        torch.nn.functional.interpolate(xr, scale_factor=4, mode="bilinear", align_corners=False).backward(g)
        adj()
        print("adjoint rel err vs torch:", float((gin - xr.grad).norm() / xr.grad.norm()))

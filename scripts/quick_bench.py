"""Quick CUDA-event timing of the loss modules at BASELINE config shapes (development aid)."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from tests.util import F2H, F2M, HI, HM


def blob(g, b, h, w, n, tile, dev):
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    c = torch.randint(0, n, (b, th, tw), generator=g, device=dev)
    c[torch.rand(b, th, tw, generator=g, device=dev) < 0.1] = 255
    return c.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w].contiguous()


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--labels", default="blob")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dtype", default="")
    ap.add_argument("--h", type=int, default=0)
    ap.add_argument("--w", type=int, default=0)
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1234)
    if a.cfg == "2":
        b, h, w, dt = a.batch or 16, a.h or 512, a.w or 1024, torch.bfloat16
        dt = {"": dt, "fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
        x = (torch.randn(b, 26, h, w, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        mod = sb.HieraTripletLoss(19, HM, HI)
        bpp = 2 * 26 * x.element_size() + 8
    else:
        b, h, w, dt = a.batch or 8, a.h or 1024, a.w or 2048, torch.float32
        dt = {"": dt, "fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
        x = (torch.randn(b, 28, h, w, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
        mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
        bpp = 3 * 28 * x.element_size() + 16
    if a.labels == "blob":
        lab = blob(g, b, h, w, 19, 32, dev)
    else:
        lab = torch.randint(0, 19, (b, h, w), generator=g, device=dev)
        lab[torch.rand(b, h, w, generator=g, device=dev) < 0.1] = 255
    emb = F.normalize(torch.randn(b, 256, h // 32, w // 32, generator=g, device=dev), dim=1).requires_grad_(True)
    step = torch.tensor([100000], device=dev)

    def fwdbwd():
        x.grad = None
        emb.grad = None
        loss = mod(step, emb, None, x, lab)
        loss.backward()

    def fwd():
        with torch.no_grad():
            mod(step, emb, None, x, lab)

    med, best = timeit(fwdbwd, a.iters)
    medf, bestf = timeit(fwd, a.iters)
    px = b * h * w
    print(f"cfg{a.cfg} {dt} labels={a.labels} B={b}: fwd+bwd median {med:.3f} ms (best {best:.3f}) -> "
          f"{px / med / 1e6:.2f} Gpix/s, {bpp * px / med / 1e6:.0f} GB/s algorithmic; fwd-only {medf:.3f} ms")


if __name__ == "__main__":
    main()

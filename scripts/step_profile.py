"""Where a step's time goes on the host and on the device (development aid): CPU time to enqueue a step, GPU time
per step (CUDA events), per-kernel stage times, for full-resolution and head-resolution (H/4) logits."""
import argparse
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from seghiero_b200 import ops
from bench import StageTimer, STAGE_NAMES, FAST_NAMES, make_labels
from tests.util import F2H, F2M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1234)
    b, h, w = a.batch, 1024, 2048
    lab = make_labels(torch, g, b, h, w, 19, "blob", dev)
    emb = F.normalize(torch.randn(b, 256, h // 32, w // 32, generator=g, device=dev), dim=1).requires_grad_(True)
    step_t = torch.tensor([100000], device=dev)
    mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
    names = dict(STAGE_NAMES)
    names.update(FAST_NAMES)
    for tag, shape, labels in (("full-res int64", (b, 28, h, w), lab), ("H/4 uint8", (b, 28, h // 4, w // 4), lab.to(torch.uint8))):
        x = (torch.randn(*shape, generator=g, device=dev) * 2).requires_grad_(True)

        def step():
            x.grad = None
            emb.grad = None
            loss = mod(step_t, emb, None, x, labels)
            loss.backward()
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        for timed in (False, True):
            timer = StageTimer() if timed else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ctx = ops.stage_timing(timer) if timed else None
            if ctx:
                ctx.__enter__()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(a.iters):
                step()
            e1.record()
            t1 = time.perf_counter()
            if ctx:
                ctx.__exit__(None, None, None)
            torch.cuda.synchronize()
            msg = f"{tag:16s} stage-events={timed!s:5s} cpu enqueue {1e3 * (t1 - t0) / a.iters:6.3f} ms/step, gpu {e0.elapsed_time(e1) / a.iters:6.3f} ms/step"
            if timed:
                st = {names.get(k, str(k)): round(t / n, 4) for k, (t, n) in timer.totals().items()}
                msg += f"  stages {st} sum {sum(st.values()):.3f}"
            print(msg)
        del x


if __name__ == "__main__":
    main()

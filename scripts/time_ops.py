"""CUDA-event timing of every public op at production shapes with the bytes it has to move (development aid: spots
kernels that sit far from their HBM bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import seghiero_b200 as sb
from bench import make_labels
from tests.util import F2H, F2M, HI, HM

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
b, h, w = 8, 1024, 2048
px = b * h * w
lab = make_labels(torch, g, b, h, w, 19, "blob", dev)
lab8 = lab.to(torch.uint8)


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


def report(name, ms, nbytes):
    print(f"{name:58s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s of algorithmic traffic")


f2m, f2h = torch.tensor(F2M, device=dev), torch.tensor(F2H, device=dev)
report("targets_three_level int64", timeit(lambda: sb.targets_three_level(lab, f2m, f2h)), px * (8 + 16))
report("targets_three_level uint8", timeit(lambda: sb.targets_three_level(lab8, f2m, f2h)), px * (1 + 2))
report("targets_two_level int64", timeit(lambda: sb.targets_two_level(lab, HI)), px * (8 + 8))
report("targets_two_level uint8", timeit(lambda: sb.targets_two_level(lab8, HI)), px * (1 + 1))
lm = sb.build_fine_to_level_map([[0, 3], [4, 9], [10, 18]], 19).to(dev)
labv = lab.clone(); labv[labv == 255] = 0
report("targets_gather int64", timeit(lambda: sb.targets_gather(labv, lm)), px * 16)
report("targets_gather uint8", timeit(lambda: sb.targets_gather(labv.to(torch.uint8), lm)), px * 2)
cmap = [(i * 13 % 256, i * 29 % 256, i * 53 % 256) for i in range(19)]
report("colorize uint8 -> [.,3]", timeit(lambda: sb.colorize(labv.to(torch.uint8), cmap)), px * 4)
xb = torch.randn(b, 28, h, w, generator=g, device=dev).bfloat16()
report("hierarchical_argmax bf16 full-res, uint8 out + accuracy", timeit(lambda: sb.hierarchical_argmax(xb, [19, 7, 2], label=lab8, out_dtype=torch.uint8)), px * (56 + 4))
xl = torch.randn(b, 28, h // 4, w // 4, generator=g, device=dev).bfloat16()
report("hierarchical_argmax from H/4 bf16 logits", timeit(lambda: sb.hierarchical_argmax(xl, [19, 7, 2], label=lab8, size=(h, w), out_dtype=torch.uint8)), px * (56 / 16 + 4))
del xb
emb = F.normalize(torch.randn(b, 256, h // 32, w // 32, generator=g, device=dev), dim=1).requires_grad_(True)
step_t = torch.tensor([100000], device=dev)
for name, mod, C in (("HieraTripletLoss 19+7", sb.HieraTripletLoss(19, HM, HI), 26),
                     ("RMIHieraTripletLoss 19/7/2", sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H)), 28)):
    for dt in (torch.float32, torch.bfloat16):
        for lowres in (False, True):
            shape = (b, C, h // 4, w // 4) if lowres else (b, C, h, w)
            x = (torch.randn(*shape, generator=g, device=dev) * 2).to(dt).requires_grad_(True)

            def step():
                x.grad = None; emb.grad = None
                mod(step_t, emb, None, x, lab8).backward()
            es = x.element_size()
            report(f"{name} {str(dt)[6:]} {'H/4' if lowres else 'full'} logits, uint8 labels, fwd+bwd", timeit(step), px * (3 * C * es + 2) if not lowres else px * (3 * C * es + 2))
            del x
for name, mod in (("TreeTripletLoss (hierarchy)", sb.TreeTripletLoss(19, HM, HI)), ("TreeTripletLoss (id lists)", sb.IdListTreeTripletLoss(19, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [11, 12, 13, 14, 15, 16, 17, 18]))):
    def step():
        emb.grad = None
        out = mod(emb, lab)
        (out[0] if isinstance(out, tuple) else out).backward()
    try:
        report(name + " fwd+bwd", timeit(step), emb.numel() * 4 * 2)
    except Exception as e:      # noqa: BLE001
        print(name, "failed:", type(e).__name__, str(e)[:100])

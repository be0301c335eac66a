"""Per-step timeline of the three-stream e2e loop of bench.py (development aid): when does each step's D2H land?"""
import os, sys, time
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
x = torch.randn(b, 28, h // 4, w // 4, generator=g, device=dev) * 2
emb = F.normalize(torch.randn(b, 256, h // 32, w // 32, generator=g, device=dev), dim=1)
mod = sb.RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M), torch.tensor(F2H))
step_t = torch.tensor([100000], device=dev)
hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
hlab = torch.empty(lab.shape, dtype=lab.dtype, pin_memory=True).copy_(lab)
hemb = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True).copy_(emb)
hgx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
hge = torch.empty(emb.shape, dtype=emb.dtype, pin_memory=True)
hloss = torch.empty(1, dtype=torch.float32, pin_memory=True)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
s_cmp = torch.cuda.current_stream(dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
dbuf = [(torch.empty_like(x), torch.empty_like(lab), torch.empty_like(emb)) for _ in range(2)]
free_ev = [None, None]
evs = []


def e2e_step(i):
    xb, lb, eb = dbuf[i % 2]
    e_a = torch.cuda.Event(enable_timing=True); e_b = torch.cuda.Event(enable_timing=True)
    e_c = torch.cuda.Event(enable_timing=True); e_d = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_in):
        if free_ev[i % 2] is not None:
            s_in.wait_event(free_ev[i % 2])
        e_a.record(s_in)
        if mode != "nocopy":
            xb.copy_(hx, non_blocking=True); lb.copy_(hlab, non_blocking=True); eb.copy_(hemb, non_blocking=True)
        ev_in = torch.cuda.Event(); ev_in.record(s_in)
    s_cmp.wait_event(ev_in)
    e_b.record(s_cmp)
    xd = xb.detach().requires_grad_(True); ed = eb.detach().requires_grad_(True)
    loss = mod(step_t, ed, None, xd, lb)
    loss.backward()
    e_c.record(s_cmp)
    ev_cmp = torch.cuda.Event(); ev_cmp.record(s_cmp)
    free_ev[i % 2] = ev_cmp
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_cmp)
        xd.grad.record_stream(s_out)
        if mode != "nocopy":
            hgx.copy_(xd.grad, non_blocking=True)
            ed.grad.record_stream(s_out)
            hge.copy_(ed.grad, non_blocking=True)
        loss.record_stream(s_out)
        hloss.copy_(loss.detach().reshape(1), non_blocking=True)
        e_d.record(s_out)
    evs.append((e_a, e_b, e_c, e_d))


for i in range(3):
    e2e_step(i)
torch.cuda.synchronize()
evs.clear()
n = 24
t0 = time.time()
cpu_t = []
for i in range(n):
    e2e_step(i)
    cpu_t.append(time.time() - t0)
s_out.synchronize()
t1 = time.time()
print(f"mode {mode}: {1e3 * (t1 - t0) / n:.3f} ms/step wall; mem reserved {torch.cuda.memory_reserved() / 1e9:.1f} GB")
base = evs[0][0]
for i, (a, b_, c, d) in enumerate(evs):
    print(f"step {i:2d}: cpu enq done {1e3 * cpu_t[i]:7.2f}  h2d start {base.elapsed_time(a):7.2f}  compute {base.elapsed_time(b_):7.2f} -> {base.elapsed_time(c):7.2f}  d2h done {base.elapsed_time(d):7.2f}")

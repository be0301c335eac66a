"""Generate golden vectors by EXECUTING the unmodified reference modules.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own, so these
files are the pin for `oracle/hiera_oracle.py` and, through it, for the CUDA
path.  Inputs are stored next to outputs so nothing depends on RNG stability
across torch versions.  The only shim applied is `torch.Tensor.cuda -> identity`
(the reference hard-codes `.cuda()` at tree_triplet_loss.py:48,54,63,65 and
rmi_tree_triplet_loss.py:53,59,68,70); the arithmetic is untouched.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("SEGHIERO_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

torch.Tensor.cuda = lambda self, *a, **k: self  # CPU shim, see module docstring
torch.set_num_threads(max(1, os.cpu_count() or 1))

from models.loss.hiera_triplet_loss import (HieraTripletLoss, _losses_hiera_two_level,  # noqa: E402
                                            _prepare_targets_two_level)
from models.loss.rmi_hiera_triplet_loss import (RMIHieraTripletLoss,  # noqa: E402
                                                _prepare_targets_three_level)
from models.loss.tree_triplet_loss import TreeTripletLoss as Triplet2  # noqa: E402
from models.loss.rmi_tree_triplet_loss import TreeTripletLoss as Triplet3  # noqa: E402
from models.loss.cross_entropy_loss import CrossEntropyLoss  # noqa: E402

HI_19_7 = [[0, 2], [2, 5], [5, 8], [8, 10], [10, 11], [11, 13], [13, 19]]
HM_19_7 = [0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 6, 6, 6, 6, 6, 6]
F2M_19 = HM_19_7
F2H_19 = [0] * 11 + [1] * 8


def blob_labels(g, b, h, w, n_fine, tile, p_ignore):
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    coarse = torch.randint(0, n_fine, (b, th, tw), generator=g)
    coarse[torch.rand(b, th, tw, generator=g) < p_ignore] = 255
    return coarse.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w].contiguous()


def iid_labels(g, b, h, w, n_fine, p_ignore):
    lab = torch.randint(0, n_fine, (b, h, w), generator=g)
    lab[torch.rand(b, h, w, generator=g) < p_ignore] = 255
    return lab


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def run_two_level(name, seed, b, h, w, d, eh, ew, step, labels, dtype=torch.float32, lw=1.0):
    g = torch.Generator().manual_seed(seed)
    lab = labels(g)
    x = (torch.randn(b, 26, h, w, generator=g) * 2).to(dtype)
    emb = F.normalize(torch.randn(b, d, eh, ew, generator=g), dim=1)
    xg = x.clone().requires_grad_(True)
    eg = emb.clone().requires_grad_(True)
    mod = HieraTripletLoss(19, HM_19_7, HI_19_7, loss_weight=lw)
    loss = mod(torch.tensor([step]), eg, None, xg, lab)
    loss.backward()
    tf, tc, _ = _prepare_targets_two_level(lab, HI_19_7)
    hiera = _losses_hiera_two_level(x, tf, tc, 19, HI_19_7)
    ce = CrossEntropyLoss()
    ce_f, ce_c = ce(x[:, :19], tf), ce(x[:, 19:26], tc)
    trip, cnt = Triplet2(19, HM_19_7, HI_19_7)(emb, lab)
    save(name, x=x.float().numpy(), emb=emb.numpy(), label=lab.numpy(), step=np.int64(step),
         loss=loss.detach().float().numpy(), dx=xg.grad.float().numpy(),
         demb=(eg.grad if eg.grad is not None else torch.zeros_like(emb)).numpy(),
         demb_none=np.bool_(eg.grad is None),
         tc=tc.numpy(), hiera=hiera.float().numpy(), ce_f=ce_f.float().numpy(), ce_c=ce_c.float().numpy(),
         triplet=np.float32(-1.0 if trip is None else float(trip)), count=cnt.numpy(),
         is_bf16=np.bool_(dtype == torch.bfloat16), loss_weight=np.float32(lw))


def run_three_level(name, seed, b, h, w, d, eh, ew, step, labels, nf, nm, nh, f2m, f2h, lam=0.5, lw=1.0,
                    dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    lab = labels(g)
    c = nf + nm + nh
    x = (torch.randn(b, c, h, w, generator=g) * 2).to(dtype)
    emb = F.normalize(torch.randn(b, d, eh, ew, generator=g), dim=1)
    f2m_t, f2h_t = torch.tensor(f2m), torch.tensor(f2h)
    out = {}
    for tag, lam_v in (("", lam), ("_lam0", 0.0)):
        xg = x.clone().requires_grad_(True)
        eg = emb.clone().requires_grad_(True)
        mod = RMIHieraTripletLoss(nf, nm, nh, f2m_t, f2h_t, loss_weight_lambda=lam_v, loss_weight=lw)
        loss = mod(torch.tensor([step]), eg, None, xg, lab)
        loss.backward()
        out["loss" + tag] = loss.detach().float().numpy()
        out["dx" + tag] = xg.grad.float().numpy()
        if tag == "":
            out["demb"] = (eg.grad if eg.grad is not None else torch.zeros_like(emb)).numpy()
            out["demb_none"] = np.bool_(eg.grad is None)
    tf, tm, th = _prepare_targets_three_level(lab, f2m_t, f2h_t)
    mod = RMIHieraTripletLoss(nf, nm, nh, f2m_t, f2h_t)
    trip, cnt = Triplet3(nf, mod.upper_ids, mod.lower_ids)(emb, lab)
    save(name, x=x.float().numpy(), emb=emb.numpy(), label=lab.numpy(), step=np.int64(step),
         tm=tm.numpy(), th=th.numpy(), nf=np.int64(nf), nm=np.int64(nm), nh=np.int64(nh),
         f2m=np.array(f2m, dtype=np.int64), f2h=np.array(f2h, dtype=np.int64),
         lam=np.float32(lam), loss_weight=np.float32(lw),
         triplet=np.float32(-1.0 if trip is None else float(trip)), count=cnt.numpy(),
         is_bf16=np.bool_(dtype == torch.bfloat16), **out)


def main():
    # --- two-level ----------------------------------------------------------
    run_two_level("two_level_blob", 11, 2, 32, 48, 16, 8, 12, 60000,
                  lambda g: blob_labels(g, 2, 32, 48, 19, 4, 0.10))
    run_two_level("two_level_iid", 12, 2, 24, 40, 8, 6, 10, 100000,
                  lambda g: iid_labels(g, 2, 24, 40, 19, 0.15), lw=0.7)
    run_two_level("two_level_bf16", 13, 1, 32, 32, 8, 4, 4, 80000,
                  lambda g: blob_labels(g, 1, 32, 32, 19, 8, 0.10), dtype=torch.bfloat16)
    run_two_level("two_level_allvoid", 14, 1, 16, 16, 8, 2, 2, 80000,
                  lambda g: torch.full((1, 16, 16), 255, dtype=torch.long))
    # --- three-level --------------------------------------------------------
    run_three_level("three_level_blob", 21, 2, 24, 40, 16, 6, 10, 100000,
                    lambda g: blob_labels(g, 2, 24, 40, 19, 4, 0.10), 19, 7, 2, F2M_19, F2H_19)
    run_three_level("three_level_iid", 22, 2, 20, 28, 8, 5, 7, 200000,
                    lambda g: iid_labels(g, 2, 20, 28, 19, 0.15), 19, 7, 2, F2M_19, F2H_19, lam=1.0, lw=0.5)
    # small hierarchy (n_fine<=15 -> id lists [1..4]/[5,6], T=60000); non-tree maps on purpose
    run_three_level("three_level_small", 23, 2, 16, 20, 8, 4, 5, 30000,
                    lambda g: blob_labels(g, 2, 16, 20, 7, 2, 0.10), 7, 3, 2,
                    [0, 0, 1, 1, 2, 2, 2], [0, 0, 0, 1, 1, 1, 0])
    run_three_level("three_level_bf16", 24, 1, 16, 24, 8, 4, 6, 100000,
                    lambda g: blob_labels(g, 1, 16, 24, 19, 4, 0.10), 19, 7, 2, F2M_19, F2H_19,
                    dtype=torch.bfloat16)
    # near-constant predictions: pr_cov ~ rank-1 + 1e-3 I (SURVEY 7.2 conditioning regime)
    g = torch.Generator().manual_seed(25)
    lab = blob_labels(g, 1, 16, 16, 19, 4, 0.0)
    x = torch.full((1, 28, 16, 16), 0.3) + 1e-3 * torch.randn(1, 28, 16, 16, generator=g)
    xg = x.clone().requires_grad_(True)
    mod = RMIHieraTripletLoss(19, 7, 2, torch.tensor(F2M_19), torch.tensor(F2H_19))
    loss = mod(torch.tensor([0]), torch.zeros(1, 4, 2, 2), None, xg, lab)
    loss.backward()
    save("three_level_flat", x=x.numpy(), label=lab.numpy(), loss=loss.detach().numpy(), dx=xg.grad.numpy())
    # --- decode -------------------------------------------------------------
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 28, 12, 20, generator=g).bfloat16().float()  # bf16-rounded -> exact ties occur
    x[0, 3, 0, 0] = float("nan")
    lab = iid_labels(g, 2, 12, 20, 19, 0.2)
    pf, pm, ph = x[:, :19].argmax(1), x[:, 19:26].argmax(1), x[:, 26:].argmax(1)
    correct = int(((pf == lab) & (lab != 255)).sum())
    total = int((lab != 255).sum())
    save("decode", x=x.numpy(), label=lab.numpy(), pf=pf.numpy(), pm=pm.numpy(), ph=ph.numpy(),
         correct=np.int64(correct), total=np.int64(total))


if __name__ == "__main__":
    main()

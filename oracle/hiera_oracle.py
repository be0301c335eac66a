"""CPU oracle for SegHiero's hierarchical-loss path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch CPU restatement (numpy for the integer work,
torch-CPU autograd for the floating-point terms) of the arithmetic performed by
the reference's `models/loss` modules.  It is the *checker* for the CUDA path:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs
(cpu_baseline, `--impl reference`, and the optional `--eager-gpu` context line
that times this same restatement as eager ATen on the GPU) may import it.
Nothing under `seghiero_b200/` imports it, and the product path raises if its
CUDA library is missing.  The functions are device-agnostic torch code, so the
full-size parity tests can evaluate them on CUDA tensors.

Parity pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against outputs of the reference modules themselves,
executed in the build container by `tests/golden/make_golden.py` and committed
as `tests/golden/*.npz` (checked by `tests/test_oracle_golden.py`).

Every function cites the reference lines it restates (paths relative to the
reference root).  Elementwise arithmetic follows the reference's fp32 order
literally where the order matters (`1 - s + eps`, `(1-ap) - (1-an) + 0.6`);
large reductions are carried in float64 (the reference's fp32 pairwise sums agree
with that to ~1e-7 relative).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

IGNORE = 255  # hard-coded in the reference's builders and triplet losses


# --------------------------------------------------------------------------- #
# A.1 target builders (bit-exact integer work)
# --------------------------------------------------------------------------- #
def last_bucket_table(hiera_index: Sequence[Sequence[int]], size: int) -> np.ndarray:
    """lut[t] = index of the LAST bucket [start,end) containing t, else 255.

    Restates the sequential masked assignment of
    models/loss/hiera_triplet_loss.py:33-36 (a later bucket overwrites an
    earlier one when they overlap).
    """
    lut = np.full(size, IGNORE, dtype=np.int64)
    for i, (start, end) in enumerate(hiera_index):
        lo, hi = max(int(start), 0), min(int(end), size)
        if hi > lo:
            lut[lo:hi] = i
    return lut


def targets_two_level(label, hiera_index):
    """models/loss/hiera_triplet_loss.py:11-38 -> (fine, coarse); same dtype."""
    lab = np.asarray(label)
    size = max([int(e) for _, e in hiera_index] + [1])
    lut = last_bucket_table(hiera_index, size)
    inside = (lab >= 0) & (lab < size)
    coarse = np.full(lab.shape, IGNORE, dtype=lab.dtype)
    coarse[inside] = lut[lab[inside]].astype(lab.dtype)
    return lab, coarse


def targets_three_level(label, fine_to_mid, fine_to_high):
    """models/loss/rmi_hiera_triplet_loss.py:21-63 -> (fine, mid, high).

    255 is literal (not ignore_index); labels outside [0,n_fine) raise
    IndexError like the reference's advanced indexing does.
    """
    lab = np.asarray(label)
    f2m = np.asarray(fine_to_mid)
    f2h = np.asarray(fine_to_high)
    keep = lab != IGNORE
    vals = lab[keep]
    if vals.size and (vals.min() < -f2m.size or vals.max() >= f2m.size):
        raise IndexError("fine label out of range for fine_to_mid/fine_to_high")
    mid = np.full(lab.shape, IGNORE, dtype=lab.dtype)
    high = np.full(lab.shape, IGNORE, dtype=lab.dtype)
    mid[keep] = f2m[vals]
    high[keep] = f2h[vals]
    return lab, mid, high


def targets_dataloader(fine_mask, level_map):
    """dataset/dataloader.py:166-177 -> map[fine_mask]; no ignore handling."""
    m = np.asarray(level_map)
    lab = np.asarray(fine_mask)
    if lab.size and (lab.min() < -m.size or lab.max() >= m.size):
        raise IndexError("index out of range in dataloader target gather")
    return m[lab]


def build_fine_to_level_map(map_cfg, n_fine: int) -> np.ndarray:
    """dataset/dataloader.py:12-34 (ranges are INCLUSIVE; must cover all ids)."""
    out = np.full(n_fine, -1, dtype=np.int64)
    for lvl, sub in enumerate(map_cfg):
        if len(sub) == 1:
            lo = hi = int(sub[0])
            assert 0 <= lo < n_fine
        elif len(sub) == 2:
            lo, hi = int(sub[0]), int(sub[1])
            assert 0 <= lo <= hi < n_fine
        else:
            raise ValueError(f"Each entry must be [lbl] or [start,end], got {sub}")
        out[lo:hi + 1] = lvl
    if (out < 0).any():
        raise ValueError(f"Fine-labels not mapped: {np.flatnonzero(out < 0).tolist()}")
    return out


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def _as_long(t, device=None):
    """index tensor on the device of the logits it will index (the restatement runs on CPU or, for full-size checks, CUDA)"""
    t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)).long()
    return t if device is None else t.to(device)


def _dsum(t: torch.Tensor) -> torch.Tensor:
    return t.double().sum()


def _pick(planes: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """planes [B,K,H,W], idx [B,H,W] in [0,K) -> [B,H,W]."""
    return planes.gather(1, idx.unsqueeze(1)).squeeze(1)


def _first_max(cols):
    """max over a python list of [B,H,W] planes; tie gradient -> first plane."""
    return torch.stack(cols, dim=1).max(dim=1).values


def _first_min(cols):
    return torch.stack(cols, dim=1).min(dim=1).values


def _bce_level(pos_plane, neg_planes, target0, valid, eps):
    """sum_px valid * [ -log(pos[target]+eps) - sum_{k != target} log(1-neg[k]+eps) ].

    pos_plane/neg_planes: [B,K,H,W] fp32; target0: void already mapped to 0.
    """
    pos = -torch.log(_pick(pos_plane, target0) + eps)
    neg_all = -torch.log(1 - neg_planes + eps)          # literal fp32 order
    neg = neg_all.double().sum(1) - _pick(neg_all, target0).double()
    return ((pos.double() + neg) * valid.double()).sum()


# --------------------------------------------------------------------------- #
# A.2 two-level tree BCE
# --------------------------------------------------------------------------- #
def tree_bce_two_level(x, tf, tc, n_fine, hiera_index, eps=1e-8):
    """models/loss/hiera_triplet_loss.py:41-107 -> 5*(L_fine+L_coarse)."""
    tf, tc = _as_long(tf, x.device), _as_long(tc, x.device)
    n_coarse = len(hiera_index)
    s = torch.sigmoid(x.float())
    A = [s[:, f] for f in range(n_fine)]
    Bc = [s[:, n_fine + i] for i in range(n_coarse)]
    # :81-85  max over (fine in bucket ..., coarse) ; ties -> lowest fine id
    mcmb = torch.stack([_first_max(A[st:en] + [Bc[i]]) for i, (st, en) in enumerate(hiera_index)], 1)
    # :88-92  min(A_f, B_bucket(f)), last bucket containing f wins; ties -> fine
    owner = last_bucket_table(hiera_index, n_fine)
    mcla = torch.stack([A[f] if owner[f] == IGNORE else _first_min([A[f], Bc[int(owner[f])]])
                        for f in range(n_fine)], 1)
    vf, vc = tf != IGNORE, tc != IGNORE
    nvf = vf.sum().clamp_min(1).double()
    nvc = vc.sum().clamp_min(1).double()
    tf0 = torch.where(vf, tf, torch.zeros_like(tf))
    tc0 = torch.where(vc, tc, torch.zeros_like(tc))
    lf = _bce_level(mcla, torch.stack(A, 1), tf0, vf, eps) / (nvf * n_fine)
    lc = _bce_level(torch.stack(Bc, 1), mcmb, tc0, vc, eps) / (nvc * n_coarse)
    return 5.0 * (lf + lc)


# --------------------------------------------------------------------------- #
# A.3 three-level tree BCE
# --------------------------------------------------------------------------- #
def tree_sets(fine_to_mid, fine_to_high, n_mid, n_high):
    """Set tables implied by rmi_hiera_triplet_loss.py:379-442.

    F(m)  = fine ids with f2m==m (ascending)
    Ms(h) = {f2m[f] : f2h[f]==h} (ascending: python set of small ints)
    Hs(m) = {f2h[f] : f in F(m)} (ascending here; order only matters on exact
            ties between different high channels, where the reference's own
            order is unspecified -- it iterates a set of tensors)
    """
    f2m = [int(v) for v in np.asarray(fine_to_mid)]
    f2h = [int(v) for v in np.asarray(fine_to_high)]
    Fm = [[f for f, m in enumerate(f2m) if m == i] for i in range(n_mid)]
    Ms = [sorted({f2m[f] for f, h in enumerate(f2h) if h == j}) for j in range(n_high)]
    Hs = [sorted({f2h[f] for f in Fm[i]}) for i in range(n_mid)]
    return Fm, Ms, Hs


def tree_bce_three_level(x, tf, tm, th, fine_to_mid, fine_to_high, n_fine, n_mid, n_high,
                         eps=1e-6, ignore_index=IGNORE):
    """models/loss/rmi_hiera_triplet_loss.py:352-470 -> 5*(L_f+L_m+L_h)."""
    tf, tm, th = _as_long(tf, x.device), _as_long(tm, x.device), _as_long(th, x.device)
    Fm, Ms, Hs = tree_sets(fine_to_mid, fine_to_high, n_mid, n_high)
    f2m = [int(v) for v in np.asarray(fine_to_mid)]
    s = torch.sigmoid(x.float())
    A = [s[:, f] for f in range(n_fine)]
    Bm = [s[:, n_fine + m] for m in range(n_mid)]
    Ch = [s[:, n_fine + n_mid + h] for h in range(n_high)]
    mcmb = [_first_max([A[f] for f in Fm[m]] + [Bm[m]]) for m in range(n_mid)]        # :379-390
    mcmc = [_first_max([mcmb[m] for m in Ms[h]] + [Ch[h]]) for h in range(n_high)]    # :394-411
    mcla = [_first_min([A[f], Bm[f2m[f]]]) for f in range(n_fine)]                    # :418-425
    mclb = [_first_min([Ch[h] for h in Hs[m]] + [Bm[m]]) if Fm[m] else Bm[m]          # :429-442
            for m in range(n_mid)]
    out = 0.0
    for tgt, pos, neg, k in ((tf, mcla, A, n_fine), (tm, mclb, mcmb, n_mid), (th, Ch, mcmc, n_high)):
        v = tgt != ignore_index
        nv = v.sum().clamp_min(1).double()
        t0 = torch.where(v, tgt, torch.zeros_like(tgt))
        out = out + _bce_level(torch.stack(pos, 1), torch.stack(neg, 1), t0, v, eps) / (nv * k)
    return 5.0 * out


# --------------------------------------------------------------------------- #
# A.7 softmax CE (mean over ALL pixels, ignored ones count in the denominator)
# --------------------------------------------------------------------------- #
def ce_level(x_level, target, ignore_index=IGNORE):
    """models/loss/cross_entropy_loss.py:7-30 + utils.py:45-47."""
    target = _as_long(target, x_level.device)
    v = target != ignore_index
    t0 = torch.where(v, target, torch.zeros_like(target))
    xl = x_level.float()
    nll = torch.logsumexp(xl, dim=1) - _pick(xl, t0)
    return (nll.double() * v.double()).sum() / float(target.numel())


# --------------------------------------------------------------------------- #
# A.4 RMI lower bound (direct definition; small shapes only)
# --------------------------------------------------------------------------- #
def rmi_windows(z: torch.Tensor) -> torch.Tensor:
    """rmi_hiera_triplet_loss.py:292-311: [B,C,H,W] -> [B,C,9,(H-2)(W-2)], k=3y+x."""
    b, c, h, w = z.shape
    cols = F.unfold(z.reshape(b * c, 1, h, w), kernel_size=3)
    return cols.reshape(b, c, 9, (h - 2) * (w - 2))


def rmi_onehot_and_probs(x, tf, tm, th, n_fine, n_mid, n_high, ignore_index=IGNORE, clip=1e-6):
    """rmi_hiera_triplet_loss.py:355-370, 479-487: void pixels are one-hot of
    class 0 at every level (not masked); P = sigmoid*valid + 1e-6 (fp32)."""
    tf, tm, th = _as_long(tf, x.device), _as_long(tm, x.device), _as_long(th, x.device)
    s = torch.sigmoid(x.float())
    hot, val = [], []
    for tgt, k in ((tf, n_fine), (tm, n_mid), (th, n_high)):
        v = tgt != ignore_index
        t0 = torch.where(v, tgt, torch.zeros_like(tgt))
        hot.append(F.one_hot(t0, k).permute(0, 3, 1, 2).float())
        val.append(v.unsqueeze(1).float().expand(-1, k, -1, -1))
    return torch.cat(hot, 1), s * torch.cat(val, 1) + clip


def rmi_from_moments(s_ll, s_pp, s_lp, alpha=1e-3):
    """rmi_hiera_triplet_loss.py:505-517 on float64 [B,C,9,9] moments -> scalar."""
    eye = torch.eye(9, dtype=torch.float64, device=s_pp.device)
    k_inv = torch.linalg.inv(s_pp + alpha * eye)
    m = s_ll - s_lp @ k_inv @ s_lp.transpose(-1, -2) + alpha * eye
    chol = torch.linalg.cholesky(m)
    r_bc = 0.5 * 2.0 * torch.log(torch.diagonal(chol, dim1=-2, dim2=-1) + 1e-8).sum(-1)
    per_class = r_bc.mean(0).float() / 9.0
    return per_class.sum(), r_bc


def rmi_lower_bound(x, tf, tm, th, n_fine, n_mid, n_high, ignore_index=IGNORE):
    """rmi_hiera_triplet_loss.py:479-517 -> rmi_loss (fp32 scalar)."""
    la, pr = rmi_onehot_and_probs(x, tf, tm, th, n_fine, n_mid, n_high, ignore_index)
    la_v = rmi_windows(la).double()
    pr_v = rmi_windows(pr).double()
    s_ll = la_v @ la_v.transpose(2, 3)
    s_pp = pr_v @ pr_v.transpose(2, 3)
    s_lp = la_v @ pr_v.transpose(2, 3)
    return rmi_from_moments(s_ll, s_pp, s_lp)[0]


# --------------------------------------------------------------------------- #
# A.5 / A.6 triplet losses
# --------------------------------------------------------------------------- #
def nearest_rows(n_in: int, n_out: int) -> np.ndarray:
    """F.interpolate(mode='nearest') source index: min(floor(dst*scale), in-1)
    with scale = float32(in)/out (tree_triplet_loss.py:17-20)."""
    scale = np.float32(n_in) / np.float32(n_out)
    src = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(src, n_in - 1)


def downsample_labels(label, h: int, w: int) -> np.ndarray:
    lab = np.asarray(label)
    ys = nearest_rows(lab.shape[-2], h)
    xs = nearest_rows(lab.shape[-1], w)
    return lab[..., ys[:, None], xs[None, :]]


def _triplet_from_sets(rows, sets, max_triplet):
    """sets: iterable of (anchor_idx, pos_idx, neg_idx) numpy index arrays."""
    total, count = 0.0, 0
    for a, p, n in sets:
        k = min(len(a), len(p), len(n), max_triplet)
        if k == 0:
            continue
        fa = rows[torch.from_numpy(a[:k]).to(rows.device)]
        fp = rows[torch.from_numpy(p[:k]).to(rows.device)]
        fn = rows[torch.from_numpy(n[:k]).to(rows.device)]
        d_pos = 1 - (fa * fp).sum(1)
        d_neg = 1 - (fa * fn).sum(1)
        total = total + F.relu(d_pos - d_neg + 0.6).mean()
        count += 1
    if count == 0:
        return None, 0
    return total / count, count


def triplet_hierarchy(feats, label, hiera_map, hiera_index, max_triplet=200):
    """models/loss/tree_triplet_loss.py:15-65 -> (loss|None, count)."""
    b, d, h, w = feats.shape
    lab = downsample_labels(label, h, w).reshape(-1)
    rows = feats.permute(0, 2, 3, 1).reshape(-1, d)
    sets = []
    for c in sorted(set(np.unique(lab).tolist()) - {IGNORE}):
        st, en = hiera_index[hiera_map[c]][0], hiera_index[hiera_map[c]][-1]
        anchor = lab == c
        sets.append((np.flatnonzero(anchor),
                     np.flatnonzero((lab >= st) & (lab < en) & ~anchor),
                     np.flatnonzero((lab < st) | (lab >= en))))      # 255 is a negative
    return _triplet_from_sets(rows, sets, max_triplet)


def triplet_id_lists(feats, label, upper_ids, lower_ids, max_triplet=200):
    """models/loss/rmi_tree_triplet_loss.py:14-70 -> (loss|None, count).
    Classes 0 and 255 are skipped; a present class in neither list raises
    ValueError exactly like the reference's list.remove."""
    b, d, h, w = feats.shape
    lab = downsample_labels(label, h, w).reshape(-1)
    rows = feats.permute(0, 2, 3, 1).reshape(-1, d)
    sets = []
    for c in sorted(set(np.unique(lab).tolist()) - {IGNORE, 0}):
        same, other = (upper_ids, lower_ids) if c in upper_ids else (lower_ids, upper_ids)
        same = list(same)
        same.remove(c)                                   # ValueError if absent
        sets.append((np.flatnonzero(lab == c),
                     np.flatnonzero(np.isin(lab, same)),
                     np.flatnonzero(np.isin(lab, list(other)))))
    return _triplet_from_sets(rows, sets, max_triplet)


def schedule_factor(step: float, total_steps: int) -> float:
    """hiera_triplet_loss.py:204-208 / rmi_hiera_triplet_loss.py:539-543."""
    if step < total_steps:
        return 0.25 * (1 + math.cos((step - total_steps) / total_steps * math.pi))
    return 0.5


def id_lists_for(n_fine: int):
    """rmi_hiera_triplet_loss.py:265-270."""
    if n_fine > 15:
        return [1, 2, 3, 4, 5, 6, 7, 10, 11, 13, 14, 15], [8, 9, 12, 16, 17, 18, 19]
    return [1, 2, 3, 4], [5, 6]


# --------------------------------------------------------------------------- #
# whole modules
# --------------------------------------------------------------------------- #
def hiera_triplet_loss(step, embedding, x, label, num_classes, hiera_map, hiera_index,
                       loss_weight=1.0, world_ready=None):
    """HieraTripletLoss.forward, hiera_triplet_loss.py:152-211.
    Returns (loss, parts).  `world_ready` overrides the local `count>0` gate
    (the reference's all_gather branch, :193-198)."""
    label_np = label.cpu().numpy() if isinstance(label, torch.Tensor) else np.asarray(label)
    tf, tc = targets_two_level(label_np, hiera_index)
    n_coarse = len(hiera_index)
    hiera = tree_bce_two_level(x, tf, tc, num_classes, hiera_index)
    ce_f = ce_level(x[:, :num_classes], tf)
    ce_c = ce_level(x[:, num_classes:num_classes + n_coarse], tc)
    loss = hiera + ce_f + ce_c
    trip, count = triplet_hierarchy(embedding, label_np, hiera_map, hiera_index)
    ready = (count > 0) if world_ready is None else world_ready
    factor = schedule_factor(float(step), 80000)
    if ready and trip is not None:
        loss = loss + factor * trip.double()
    parts = dict(hiera=float(hiera.detach()), ce=[float(ce_f.detach()), float(ce_c.detach())],
                 triplet=None if trip is None else float(trip.detach()), count=count, factor=factor,
                 targets=(tf, tc))
    return (loss * loss_weight).float(), parts


def rmi_hiera_triplet_loss(step, embedding, x, label, n_fine, n_mid, n_high, fine_to_mid,
                           fine_to_high, loss_weight_lambda=0.5, loss_weight=1.0,
                           ignore_index=IGNORE, with_triplet=True, world_ready=None):
    """RMIHieraTripletLoss.forward, rmi_hiera_triplet_loss.py:323-546."""
    label_np = label.cpu().numpy() if isinstance(label, torch.Tensor) else np.asarray(label)
    tf, tm, th = targets_three_level(label_np, fine_to_mid, fine_to_high)
    hiera = tree_bce_three_level(x, tf, tm, th, fine_to_mid, fine_to_high, n_fine, n_mid, n_high,
                                 ignore_index=ignore_index)
    rmi = rmi_lower_bound(x, tf, tm, th, n_fine, n_mid, n_high, ignore_index)
    ce = [ce_level(x[:, :n_fine], tf),
          ce_level(x[:, n_fine:n_fine + n_mid], tm),
          ce_level(x[:, n_fine + n_mid:n_fine + n_mid + n_high], th)]
    loss = loss_weight_lambda * rmi.double() + 0.5 * hiera + ce[0] + ce[1] + ce[2]
    trip, count = (None, 0)
    if with_triplet:
        upper, lower = id_lists_for(n_fine)
        trip, count = triplet_id_lists(embedding, label_np, upper, lower)
    ready = (count > 0) if world_ready is None else world_ready
    factor = schedule_factor(float(step), 160000 if n_fine > 15 else 60000)
    if ready and trip is not None:
        loss = loss + factor * trip.double()
    parts = dict(hiera=float(hiera.detach()), rmi=float(rmi.detach()), ce=[float(c.detach()) for c in ce],
                 triplet=None if trip is None else float(trip.detach()), count=count, factor=factor,
                 targets=(tf, tm, th))
    return (loss * loss_weight).float(), parts


# --------------------------------------------------------------------------- #
# decode / metrics (infer.py:303-312, train.py:37-49, 382-385)
# --------------------------------------------------------------------------- #
def argmax_decode(x, level_sizes):
    """Independent per-level argmax over channel slices; first max wins, NaN
    counts as max (torch.argmax semantics).  x: [B,C,H,W] -> list of int64."""
    arr = x.detach().float().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    out, c0 = [], 0
    for k in level_sizes:
        out.append(np.argmax(arr[:, c0:c0 + k], axis=1).astype(np.int64))
        c0 += k
    return out


def pixel_accuracy_counts(pred, target, ignore_index=IGNORE):
    pred, target = np.asarray(pred), np.asarray(target)
    valid = target != ignore_index
    return int(((pred == target) & valid).sum()), int(valid.sum())


# --------------------------------------------------------------------------- #
# the callers either side of the loss
# --------------------------------------------------------------------------- #
def _bilinear_axis(n_in: int, n_out: int, device):
    """Source rows and weights of F.interpolate(mode='bilinear', align_corners=False) along one axis, fp32 like ATen:
    src = (in / out) * (dst + 0.5) - 0.5 clamped at 0; i0 = floor(src); i1 = i0 + (i0 < in - 1); l1 = src - i0."""
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = np.maximum(scale * (dst + np.float32(0.5)) - np.float32(0.5), np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(device)
    return t(i0), t(i1), t(l0), t(l1)


def interpolate_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(x, size=size, mode='bilinear', align_corners=False) as train.py:282-284, 309-312 and
    infer.py:296-299 call it, restated with explicit gathers (differentiable; result in x's dtype)."""
    hh, ww = int(size[0]), int(size[1])
    y0, y1, h0, h1 = _bilinear_axis(x.shape[2], hh, x.device)
    x0, x1, w0, w1 = _bilinear_axis(x.shape[3], ww, x.device)
    xf = x.float()
    top, bot = xf[:, :, y0, :], xf[:, :, y1, :]
    h0, h1 = h0.view(1, 1, hh, 1), h1.view(1, 1, hh, 1)
    w0, w1 = w0.view(1, 1, 1, ww), w1.view(1, 1, 1, ww)
    out = h0 * (w0 * top[:, :, :, x0] + w1 * top[:, :, :, x1]) + h1 * (w0 * bot[:, :, :, x0] + w1 * bot[:, :, :, x1])
    return out.to(x.dtype)


def aux_cross_entropy(x_low: torch.Tensor, label) -> torch.Tensor:
    """train.py:309-313: nn.CrossEntropyLoss(ignore_index=255) of the bilinearly upsampled aux logits
    (mean over the NON-ignored pixels, unlike the loss modules' own CE wrapper)."""
    lab = _as_long(label, x_low.device)
    full = interpolate_bilinear(x_low, lab.shape[-2:]).float()
    valid = lab != IGNORE
    logp = torch.log_softmax(full, dim=1)
    picked = _pick(logp, torch.where(valid, lab, torch.zeros_like(lab)))
    return -(picked * valid).double().sum().float() / valid.sum().float()


def argmax_decode_upsampled(x_low: torch.Tensor, level_sizes, size):
    """train.py:345-352, 382-385 / infer.py:296-312: per-level argmax of the upsampled logits."""
    return argmax_decode(interpolate_bilinear(x_low, size), level_sizes)


def colorize(mask, colormap) -> np.ndarray:
    """infer.py:117-131 (`mask_to_color_image`): negative ids black, others colormap[id]; [H,W] -> uint8 [H,W,3]."""
    m = np.asarray(mask)
    cm = np.asarray(colormap, dtype=np.uint8).reshape(-1, 3)
    if m.size and m.max() >= len(cm):
        raise IndexError("list index out of range")
    out = np.zeros(m.shape + (3,), dtype=np.uint8)
    pos = m >= 0
    out[pos] = cm[m[pos]]
    return out

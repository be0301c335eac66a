"""Host-side hierarchy tables (small int32 arrays) consumed by the CUDA kernels.

Pure Python/numpy so it can be unit-tested without a GPU.  Semantics follow the
reference's per-class Python loops:
  two-level  : models/loss/hiera_triplet_loss.py:28-36, 81-92
  three-level: models/loss/rmi_hiera_triplet_loss.py:379-442
  triplet    : models/loss/tree_triplet_loss.py:32-36, models/loss/rmi_tree_triplet_loss.py:28-45
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

IGNORE = 255


def two_level_tables(n_fine: int, hiera_index: Sequence[Sequence[int]]):
    """-> (int32 blob [bstart][bend][owner][fb_ptr][fb_idx][lut], n_fb, lut_size)."""
    nc = len(hiera_index)
    bstart = np.array([max(0, min(int(s), n_fine)) for s, _ in hiera_index], dtype=np.int32)
    bend = np.array([max(0, min(int(e), n_fine)) for _, e in hiera_index], dtype=np.int32)
    bend = np.maximum(bend, bstart)
    owner = np.full(n_fine, -1, dtype=np.int32)           # last bucket containing f (later bucket wins)
    members = [[] for _ in range(n_fine)]
    for i in range(nc):
        for f in range(bstart[i], bend[i]):
            owner[f] = i
            members[f].append(i)
    fb_ptr = np.zeros(n_fine + 1, dtype=np.int32)
    fb_idx = []
    for f in range(n_fine):
        fb_idx.extend(members[f])
        fb_ptr[f + 1] = len(fb_idx)
    # target LUT follows the RAW ranges (labels >= n_fine are an error in the loss, but the
    # standalone target builder accepts any range the caller wrote)
    lut_size = max([int(e) for _, e in hiera_index] + [1])
    lut = np.full(lut_size, IGNORE, dtype=np.int32)
    for i, (s, e) in enumerate(hiera_index):
        lo, hi = max(int(s), 0), min(int(e), lut_size)
        if hi > lo:
            lut[lo:hi] = i
    blob = np.concatenate([bstart, bend, owner, fb_ptr, np.array(fb_idx, dtype=np.int32), lut]).astype(np.int32)
    return blob, len(fb_idx), lut_size


def two_level_is_tree(n_fine: int, hiera_index: Sequence[Sequence[int]]) -> bool:
    """True when no fine class sits in two buckets (disjoint ranges): the fast 2-level kernel applies."""
    seen = np.zeros(n_fine, dtype=np.int32)
    for s, e in hiera_index:
        lo, hi = max(0, min(int(s), n_fine)), max(0, min(int(e), n_fine))
        if hi > lo:
            seen[lo:hi] += 1
    return bool((seen <= 1).all())


def three_level_tables(n_fine: int, n_mid: int, n_high: int, fine_to_mid, fine_to_high):
    """-> (int32 blob [f2m][f2h][mh_ptr][mh_idx][hsmask][order C][fast order C][fast aux C], n_mh, fast_ok).
    Validates the maps."""
    f2m = np.asarray(fine_to_mid, dtype=np.int64).reshape(-1)
    f2h = np.asarray(fine_to_high, dtype=np.int64).reshape(-1)
    if f2m.size != n_fine or f2h.size != n_fine:
        raise ValueError("fine_to_mid / fine_to_high must have n_fine entries")
    if (f2m < 0).any() or (f2m >= n_mid).any():
        raise ValueError("fine_to_mid holds ids outside [0, n_mid)")
    if (f2h < 0).any() or (f2h >= n_high).any():
        raise ValueError("fine_to_high holds ids outside [0, n_high) (uninitialised map?)")
    if n_high > 32:
        raise ValueError("n_high > 32 is not supported")
    if n_fine + n_mid + n_high > 254:
        raise ValueError("more than 254 channels are not supported")
    # Ms(h) = {f2m[f] : f2h[f]==h};  mh list of mid m = highs whose Ms contains m
    mh = [sorted({int(f2h[f]) for f in range(n_fine) if f2m[f] == m}) for m in range(n_mid)]
    mh_ptr = np.zeros(n_mid + 1, dtype=np.int32)
    mh_idx = []
    hsmask = np.zeros(n_mid, dtype=np.uint32)
    for m in range(n_mid):
        mh_idx.extend(mh[m])
        mh_ptr[m + 1] = len(mh_idx)
        for h in mh[m]:                                   # Hs(m) = {f2h[f] : f in F(m)} -- the same set
            hsmask[m] |= np.uint32(1) << np.uint32(h)
    fast = fast_tree_order(n_fine, n_mid, n_high, f2m, f2h)
    fast_ok = fast is not None
    if fast is None:
        fast = np.zeros(2 * (n_fine + n_mid + n_high), dtype=np.int32)
    blob = np.concatenate([f2m.astype(np.int32), f2h.astype(np.int32), mh_ptr,
                           np.array(mh_idx, dtype=np.int32), hsmask.view(np.int32),
                           tree_order(n_fine, n_mid, n_high, f2m), fast]).astype(np.int32)
    return blob, len(mh_idx), int(fast_ok)


def fast_tree_order(n_fine: int, n_mid: int, n_high: int, f2m, f2h):
    """Channel order of the warp-specialised kernels (csrc/rmi3_fast.cuh), or None when the maps are
    not a tree (some mid with fine children under two different highs): per high h, per mid m under h
    (ascending) the fine children of m (ascending) then m, then h itself; mids/highs without children
    come in id order.  Entry = kind | class << 8 | flags << 16 | channel << 24; flags bit0 / bit2 = first
    entry of a mid / high group (the backward pass switches its holder bytes there), bit1 = flush the level's
    running product of (1 - s + eps) factors (at most 5 factors of >= 1e-6 stay in fp32 range)."""
    f2m = [int(v) for v in f2m]
    f2h = [int(v) for v in f2h]
    mid_high = {}
    for f in range(n_fine):
        if mid_high.setdefault(f2m[f], f2h[f]) != f2h[f]:
            return None
    entries = []
    placed_mid = set()

    def put_mid(m):
        first = len(entries)
        for f in range(n_fine):
            if f2m[f] == m:
                entries.append([0, f, 0, f])
        entries.append([1, m, 0, n_fine + m])
        entries[first][2] |= 1                            # bit0: first entry of a mid group
        placed_mid.add(m)

    for h in range(n_high):
        first = len(entries)
        for m in range(n_mid):
            if mid_high.get(m) == h:
                put_mid(m)
        entries.append([2, h, 0, n_fine + n_mid + h])
        entries[first][2] |= 4                            # bit2: first entry of a high group
    # childless mids: their max is their own sigmoid and they feed no high; order is irrelevant but
    # they must not sit between a high's mids and the high itself (each is a high group of its own, high = none)
    for m in range(n_mid):
        if m not in placed_mid:
            first = len(entries)
            put_mid(m)
            entries[first][2] |= 4
    seen = [0, 0, 0]
    for e in entries:
        seen[e[0]] += 1
        if seen[e[0]] % 5 == 0:
            e[2] |= 2
    assert len(entries) == n_fine + n_mid + n_high
    order = np.array([k | (c << 8) | (fl << 16) | (ch << 24) for k, c, fl, ch in entries], dtype=np.uint32)
    # second table (backward pass): mid id | high id << 8 of every entry, 0xff = none
    aux = []
    for k, c, _, _ in entries:
        if k == 0:
            aux.append(f2m[c] | (f2h[c] << 8))
        elif k == 1:
            aux.append(c | (mid_high.get(c, 0xff) << 8))
        else:
            aux.append(0xff | (c << 8))
    return np.concatenate([order, np.array(aux, dtype=np.uint32)]).view(np.int32)


def tree_order(n_fine: int, n_mid: int, n_high: int, f2m) -> np.ndarray:
    """Channel visiting order of the forward pass: for every mid m its fine children (ascending), then
    the mid channel itself; the high channels last.  The running max over a mid's children then lives
    in registers.  Entry = kind | class << 8 | flags << 16 with kind 0/1/2 = fine/mid/high,
    flags bit0 = reset the running max (first entry of a mid group), bit1 = flush the level's product
    of (1 - s + eps) factors into a log (every 4th entry of a level and its last one)."""
    entries = []
    for m in range(n_mid):
        first = True
        for f in range(n_fine):
            if int(f2m[f]) == m:
                entries.append([0, f, 1 if first else 0])
                first = False
        entries.append([1, m, 1 if first else 0])
    for hh in range(n_high):
        entries.append([2, hh, 0])
    seen = [0, 0, 0]
    total = [n_fine, n_mid, n_high]
    for e in entries:
        seen[e[0]] += 1
        if seen[e[0]] % 4 == 0 or seen[e[0]] == total[e[0]]:
            e[2] |= 2
    assert len(entries) == n_fine + n_mid + n_high
    return np.array([k | (c << 8) | (fl << 16) for k, c, fl in entries], dtype=np.int32)


def triplet_tables_hierarchy(hiera_map: Sequence[int], hiera_index: Sequence[Sequence[int]]):
    """mode 0: per class c the bucket [lo, hi) = hiera_index[hiera_map[c]][0], [-1]."""
    lo = np.array([int(hiera_index[hiera_map[c]][0]) for c in range(len(hiera_map))], dtype=np.int32)
    hi = np.array([int(hiera_index[hiera_map[c]][-1]) for c in range(len(hiera_map))], dtype=np.int32)
    return np.concatenate([lo, hi]), len(hiera_map)


def triplet_tables_id_lists(upper_ids: Sequence[int], lower_ids: Sequence[int]):
    """mode 1: group id per label value (0 = upper, 1 = lower, -1 = neither)."""
    grp = np.full(256, -1, dtype=np.int32)
    for v in lower_ids:
        if 0 <= int(v) < 256:
            grp[int(v)] = 1
    for v in upper_ids:                                   # `if ii in upper_ids` is tested first
        if 0 <= int(v) < 256:
            grp[int(v)] = 0
    return grp, 256

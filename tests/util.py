import numpy as np
import torch

HI = [[0, 2], [2, 5], [5, 8], [8, 10], [10, 11], [11, 13], [13, 19]]
HM = [0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 6, 6, 6, 6, 6, 6]
F2M = HM
F2H = [0] * 11 + [1] * 8


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def blob_labels(g, b, h, w, n_fine, tile, p_ignore):
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    coarse = torch.randint(0, n_fine, (b, th, tw), generator=g)
    coarse[torch.rand(b, th, tw, generator=g) < p_ignore] = 255
    return coarse.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w].contiguous()


def iid_labels(g, b, h, w, n_fine, p_ignore):
    lab = torch.randint(0, n_fine, (b, h, w), generator=g)
    lab[torch.rand(b, h, w, generator=g) < p_ignore] = 255
    return lab


def to_np(t):
    return t.detach().float().cpu().numpy()

"""Kernel-shaped numpy model of the RMI term.  TEST INFRASTRUCTURE ONLY.

The CUDA path never materialises the reference's [B,C,9,N] unfolds
(rmi_hiera_triplet_loss.py:292-311, 493-510).  It uses the decomposition below,
which this file states in numpy/float64 so that it can be checked on the CPU
against the direct definition in `hiera_oracle.rmi_lower_bound` (see
tests/test_rmi_taps.py).  The CUDA kernels in seghiero_b200/csrc mirror these
formulas one to one.

Window k = 3*y + x has offset d_k = (y, x).  With the P-side pixel r = p + d_j as
anchor ("j-role") and R_j = { p + d_j : p a valid window origin }:

    S_xy[i, j] = sum_{r in R_j} Y[r] * X[r + d_i - d_j]

All R_j contain the interior I = [2,H-2) x [2,W-2); for r in I every (i, j) is
valid, so the interior contribution depends only on the tap d = d_i - d_j:

    T_xy(d) = sum_{r in I} Y[r] * X[r + d]          (25 taps, |d| <= 2)

Pixels of the 2-wide frame contribute directly to the (i, j) they are valid for.
S_pp and S_ll are symmetric, so only taps in the half plane H+ (dy > 0, or
dy == 0 and dx >= 0; 13 taps) are accumulated and the rest is mirrored.
"""
from __future__ import annotations

import numpy as np

ALPHA = 1e-3
OFFS = [(k // 3, k % 3) for k in range(9)]


def tap_index(dy: int, dx: int) -> int:
    return (dy + 2) * 5 + (dx + 2)


def in_half_plane(dy: int, dx: int) -> bool:
    return dy > 0 or (dy == 0 and dx >= 0)


# tap id for every (i, j): d = d_i - d_j
TAP_OF = np.array([[tap_index(OFFS[i][0] - OFFS[j][0], OFFS[i][1] - OFFS[j][1]) for j in range(9)]
                   for i in range(9)])


def interior_taps(Y: np.ndarray, X: np.ndarray) -> np.ndarray:
    """T(d) for all 25 taps; Y, X are [H, W] float64 planes."""
    h, w = Y.shape
    out = np.zeros(25)
    if h < 5 or w < 5:
        return out
    yi = Y[2:h - 2, 2:w - 2]
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            out[tap_index(dy, dx)] = (yi * X[2 + dy:h - 2 + dy, 2 + dx:w - 2 + dx]).sum()
    return out


def frame_pixels(h: int, w: int):
    for y in range(h):
        for x in range(w):
            if not (2 <= y < h - 2 and 2 <= x < w - 2):
                yield y, x


def frame_moments(Y: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Direct [9,9] contributions of frame anchors r (j-role)."""
    h, w = Y.shape
    out = np.zeros((9, 9))
    for y, x in frame_pixels(h, w):
        for j, (yj, xj) in enumerate(OFFS):
            py, px = y - yj, x - xj
            if not (0 <= py < h - 2 and 0 <= px < w - 2):
                continue
            for i, (yi, xi) in enumerate(OFFS):
                out[i, j] += Y[y, x] * X[py + yi, px + xi]
    return out


def assemble(taps: np.ndarray, frame: np.ndarray, symmetric: bool) -> np.ndarray:
    """Matrix from interior taps + frame part.  For symmetric moments only the
    H+ taps / (i,j) with d_i-d_j in H+ are trusted and the rest is mirrored."""
    out = np.zeros((9, 9))
    for i in range(9):
        for j in range(9):
            dy, dx = OFFS[i][0] - OFFS[j][0], OFFS[i][1] - OFFS[j][1]
            if symmetric and not in_half_plane(dy, dx):
                continue
            out[i, j] = taps[tap_index(dy, dx)] + frame[i, j]
    if symmetric:
        for i in range(9):
            for j in range(9):
                dy, dx = OFFS[i][0] - OFFS[j][0], OFFS[i][1] - OFFS[j][1]
                if not in_half_plane(dy, dx):
                    out[i, j] = out[j, i]
    return out


def moments_by_taps(P: np.ndarray, L: np.ndarray):
    """(S_ll, S_pp, S_lp) of one (b, c) plane via taps + frame."""
    s_pp = assemble(interior_taps(P, P), frame_moments(P, P), True)
    s_ll = assemble(interior_taps(L, L), frame_moments(L, L), True)
    s_lp = assemble(interior_taps(P, L), frame_moments(P, L), False)
    return s_ll, s_pp, s_lp


def rmi_algebra(s_ll, s_pp, s_lp):
    """r = sum_k log(chol(M)_kk + 1e-8) and its adjoints.

    Returns (r, G_pp, G_lp) with G = d r / d S treating the +1e-8 as absent
    in the derivative (relative effect < 3e-7 since chol(M)_kk >= sqrt(1e-3)).
    rmi_hiera_triplet_loss.py:505-517, 313-317.
    """
    eye = np.eye(9)
    k_inv = np.linalg.inv(s_pp + ALPHA * eye)
    m = s_ll - s_lp @ k_inv @ s_lp.T + ALPHA * eye
    chol = np.linalg.cholesky(m)
    r = np.log(np.diag(chol) + 1e-8).sum()
    w_inv = np.linalg.inv(m)
    g_lp = -w_inv @ s_lp @ k_inv
    g_pp = 0.5 * k_inv @ s_lp.T @ w_inv @ s_lp @ k_inv
    return r, g_pp, g_lp


def stencil_weights(g: np.ndarray) -> np.ndarray:
    """W(d) = sum over (i,j) with d_i - d_j = d of g[i,j]  (25 taps)."""
    out = np.zeros(25)
    for i in range(9):
        for j in range(9):
            out[TAP_OF[i, j]] += g[i, j]
    return out


def grad_wrt_P(P: np.ndarray, L: np.ndarray, g_pp: np.ndarray, g_lp: np.ndarray) -> np.ndarray:
    """d r / d P[r]: interior = 5x5 stencil, frame = restricted direct sum."""
    h, w = P.shape
    w1 = stencil_weights(g_pp)
    w2 = stencil_weights(g_lp)
    out = np.zeros_like(P)
    for y in range(2, h - 2):
        for x in range(2, w - 2):
            acc = 0.0
            for dy in range(-2, 3):
                for dx in range(-2, 3):
                    t = tap_index(dy, dx)
                    acc += 2.0 * w1[t] * P[y + dy, x + dx] + w2[t] * L[y + dy, x + dx]
            out[y, x] = acc
    for y, x in frame_pixels(h, w):
        acc = 0.0
        for j, (yj, xj) in enumerate(OFFS):
            py, px = y - yj, x - xj
            if not (0 <= py < h - 2 and 0 <= px < w - 2):
                continue
            for i, (yi, xi) in enumerate(OFFS):
                acc += 2.0 * g_pp[i, j] * P[py + yi, px + xi] + g_lp[i, j] * L[py + yi, px + xi]
        out[y, x] = acc
    return out

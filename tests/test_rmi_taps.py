"""Check the tap/frame decomposition used by the CUDA RMI kernels against the
direct unfold definition (CPU only)."""
import numpy as np
import pytest
import torch

from oracle import hiera_oracle as O
from oracle import rmi_taps as T


def _planes(seed, h, w):
    g = torch.Generator().manual_seed(seed)
    P = torch.rand(h, w, generator=g, dtype=torch.float64) + 1e-6
    L = (torch.rand(h, w, generator=g) < 0.4).double()
    return P, L


@pytest.mark.parametrize("h,w", [(3, 3), (3, 7), (4, 5), (5, 5), (6, 9), (11, 8)])
def test_moments_match_unfold(h, w):
    P, L = _planes(h * 100 + w, h, w)
    pv = O.rmi_windows(P.view(1, 1, h, w))[0, 0]
    lv = O.rmi_windows(L.view(1, 1, h, w))[0, 0]
    s_ll, s_pp, s_lp = T.moments_by_taps(P.numpy(), L.numpy())
    assert np.allclose(s_ll, (lv @ lv.T).numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(s_pp, (pv @ pv.T).numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(s_lp, (lv @ pv.T).numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("h,w", [(3, 4), (5, 6), (9, 12)])
def test_gradient_matches_autograd(h, w):
    P, L = _planes(h * 31 + w, h, w)
    Pt = P.clone().requires_grad_(True)
    pv = O.rmi_windows(Pt.view(1, 1, h, w))
    lv = O.rmi_windows(L.view(1, 1, h, w))
    _, r_bc = O.rmi_from_moments(lv @ lv.transpose(2, 3), pv @ pv.transpose(2, 3), lv @ pv.transpose(2, 3))
    r_bc.sum().backward()
    s_ll, s_pp, s_lp = T.moments_by_taps(P.numpy(), L.numpy())
    r, g_pp, g_lp = T.rmi_algebra(s_ll, s_pp, s_lp)
    assert abs(r - float(r_bc)) < 1e-10 * max(1.0, abs(r))
    g = T.grad_wrt_P(P.numpy(), L.numpy(), g_pp, g_lp)
    ref = Pt.grad.numpy()
    assert np.linalg.norm(g - ref) <= 2e-6 * np.linalg.norm(ref)

"""Gradient of the log-likelihood.

The reference has no gradient, so parity is anchored twice: the analytic numpy restatement
(oracle/gp_grad_oracle.py, built on the closed-form covariance that is pinned to the reference)
is checked against central differences of the reference-literal log-likelihood (CPU test), and
the CUDA path is checked against that analytic form (GPU tests).

Tolerance of the GPU comparison: every component within 1e-8 of the scale of the gradient,
|dg_k| <= 1e-8 * max(|g_k|, 1e-3 * max_k |g_k|).  The components are sums over n^2 products of
mixed sign, K^-1 comes from two different libraries (LAPACK on the host, cuSOLVER on the
device), and cond(K) reaches 1e6 on these inputs; the log-likelihood itself keeps its 1e-9.
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import gp_grad_oracle as gg
from oracle import gp_oracle as go

GRAD_RTOL = 1e-8


def _close(g, want):
    scale = np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    return np.max(np.abs(g - want) / scale)


@pytest.mark.parametrize("kname", ["QuasiPeriodic", "SquaredExponential"])
@pytest.mark.parametrize("jitter", [False, True])
def test_analytic_oracle_against_central_differences_of_the_reference_path(kname, jitter):
    N = 24
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=3)
    y = go.biggp_y(rv, rhk, bis)
    a = go.synthetic_biggp_hypers(1, seed=5, kname=kname)[0]
    jit = np.array([0.3, 0.1, 0.2]) if jitter else None
    ll, g = gg.biggp_loglike_grad(kname, a, t, y, e1, e2, e3, jitter=jit)
    assert rel_err(ll, go.biggp_loglike_literal(kname, a, t, y, e1, e2, e3, jitter=jit)) <= 1e-10
    fd = gg.finite_difference_grad(kname, a, t, y, e1, e2, e3, jitter=jit, rel=2e-6)
    # central differences of a float64 likelihood: truncation + cancellation leave ~1e-4
    assert _close(g, fd) <= 2e-3, (g, fd)
    assert g.shape == (len(a) + (3 if jitter else 0),)


@pytest.mark.gpu
@pytest.mark.parametrize("N,kname,jitter", [(1, "QuasiPeriodic", False), (7, "SquaredExponential", True),
                                            (40, "QuasiPeriodic", True), (64, "QuasiPeriodic", False),
                                            (100, "SquaredExponential", False), (150, "QuasiPeriodic", True)])
def test_gpu_gradient_against_analytic_oracle(N, kname, jitter):
    import torch
    assert torch.cuda.is_available()
    import miniframe_b200 as mf
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=N)
    y = go.biggp_y(rv, rhk, bis)
    B = 5
    H = go.synthetic_biggp_hypers(B, seed=N, kname=kname)
    J = go.synthetic_jitter(B, seed=N) if jitter else None
    cls = mf.BIGgpPlusJitt.BIGgp if jitter else mf.BIGgp.BIGgp
    gp = cls(getattr(mf.kernels, kname), [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    if jitter:
        ll, g = gp.log_likelihood_grad_batch(H, None, J)
    else:
        ll, g = gp.log_likelihood_grad_batch(H)
    assert ll.shape == (B,) and g.shape == (B, H.shape[1] + (3 if jitter else 0))
    worst = 0.0
    for b in range(B):
        wl, wg = gg.biggp_loglike_grad(kname, H[b], t, y, e1, e2, e3, jitter=None if J is None else J[b])
        assert rel_err(ll[b], wl) <= 1e-9
        worst = max(worst, _close(g[b], wg))
    print("N %d %s jitter %s: worst gradient error %.2e of scale" % (N, kname, jitter, worst))
    assert worst <= GRAD_RTOL
    # torch in -> torch out with the same numbers; the single call is row 0 (the library's
    # triangular inverse takes another code path for a batch of one: compare to tolerance)
    again = gp.log_likelihood_grad_batch(torch.as_tensor(H, device="cuda"), *(() if not jitter else (None, torch.as_tensor(J, device="cuda"))))
    assert again[1].is_cuda and np.array_equal(again[1].cpu().numpy(), g)
    one = gp.log_likelihood_grad(H[0], [], *(() if not jitter else (J[0],)))
    assert one[0] == ll[0] and _close(one[1], g[0]) <= GRAD_RTOL


@pytest.mark.gpu
def test_gpu_gradient_failed_sample_and_mean_parameters():
    import miniframe_b200 as mf
    N = 30
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=8)
    gp = mf.BIGgp.BIGgp(mf.kernels.SquaredExponential, [mf.means.Constant, None, mf.means.Linear], t, rv, e1,
                        rhk, e2, bis, e3)
    good = [3.0, 0.5, 1.0, 0.5, 0.8, 0.9, 0.4]
    bad = [5000.0, 0.0, 1e8, 1e8, 1e8, 1e8, 1e8]
    A = np.array([good, bad, good])
    Bm = np.array([[0.1, 0.2, -0.3], [0.0, 0.0, 0.0], [0.5, -0.1, 0.2]])
    ll, g = gp.log_likelihood_grad_batch(A, Bm)
    assert ll[1] == -np.inf and np.all(np.isnan(g[1]))
    for b in (0, 2):
        gp.mean_pars = Bm[b]
        resid = gp.y - gp.mean()
        wl, wg = gg.biggp_loglike_grad("SquaredExponential", A[b], t, resid, e1, e2, e3)
        assert rel_err(ll[b], wl) <= 1e-9 and _close(g[b], wg) <= GRAD_RTOL

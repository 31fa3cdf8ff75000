"""CPU oracle for the gradient of the log-likelihood.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference has no gradient (its
optimisers, had it any, would difference ``minus_log_likelihood``, BIGgp.py:315-317), so
parity is anchored twice: this analytic restatement

    d ll / d theta = 1/2 alpha^T (dK/dtheta) alpha - 1/2 tr(K^-1 dK/dtheta),  alpha = K^-1 r,

built on the closed-form covariance of ``gp_oracle`` (itself pinned to the reference), is checked
against central differences of the reference-literal ``biggp_loglike_literal`` in tests/test_grad.py,
and the CUDA path is checked against it.

Parameter order: the reference's ``a`` (kernel parameters, then vc, vr, lc, bc, br), followed by
the three jitters for BIGgpPlusJitt.
"""
import numpy as np
import scipy.linalg as sla

from . import gp_oracle as go

pi = np.pi


def closed_bases_grad(kname, r, kp):
    """(E, AE, Gdd) and their derivatives with respect to the kernel's shape parameters
    (QuasiPeriodic: ell_e, ell_p, P; SquaredExponential: ell), white noise excluded."""
    if kname == "QuasiPeriodic":
        le, lp, P = kp[0], kp[1], kp[2]
        f2, f3 = lp ** 2, le ** 2
        x = pi * r / P
        s, c = np.sin(x), np.cos(x)
        E = np.exp(-(2.0 * s * s / f2) - 0.5 * r * r / f3)
        A = 4 * pi * s * c / (f2 * P) + r / f3
        q0 = 1.0 / f3 + 4 * pi * pi * (c * c - s * s) / (f2 * P * P)
        dx = -x / P
        dE = [E * r * r / le ** 3, E * 4.0 * s * s / lp ** 3, E * (-4.0 * s * c * dx / f2)]
        dA = [-2.0 * r / le ** 3, -8 * pi * s * c / (lp ** 3 * P),
              4 * pi / f2 * ((c * c - s * s) * dx / P - s * c / (P * P))]
        dq = [-2.0 / le ** 3 + 0 * r, -8 * pi * pi * (c * c - s * s) / (lp ** 3 * P * P),
              4 * pi * pi / f2 * (-4.0 * s * c * dx / (P * P) - 2.0 * (c * c - s * s) / P ** 3)]
    else:
        ell = kp[0]
        f2 = ell ** 2
        E = np.exp(-0.5 * r * r / f2)
        A = r / f2
        q0 = 1.0 / f2 + 0 * r
        dE = [E * r * r / ell ** 3]
        dA = [-2.0 * r / ell ** 3]
        dq = [-2.0 / ell ** 3 + 0 * r]
    AE = A * E
    Gdd = (q0 - A * A) * E
    dAE = [da * E + A * de for da, de in zip(dA, dE)]
    dGdd = [(dqq - 2.0 * A * da) * E + (q0 - A * A) * de for dqq, da, de in zip(dq, dA, dE)]
    return (E, AE, Gdd), dE, dAE, dGdd


def _coef_grads(s5):
    """d(cE, cAE, cGdd, cW)/d(vc, vr, lc, bc, br) of gp_oracle.biggp_coefs, upper blocks."""
    vc, vr, lc, bc, br = s5
    z = 0.0
    # each entry: list over the five scaling parameters of (dcE, dcAE, dcGdd, dcW)
    return {
        (0, 0): [(2 * vc, z, z, 2 * (vc + vr)), (z, z, 2 * vr, 2 * (vc + vr)), (z, z, z, z), (z, z, z, z), (z, z, z, z)],
        (1, 1): [(z, z, z, z), (z, z, z, z), (2 * lc, z, z, 2 * lc), (z, z, z, z), (z, z, z, z)],
        (2, 2): [(z, z, z, z), (z, z, z, z), (z, z, z, z), (2 * bc, z, z, 2 * (bc + br)), (z, z, 2 * br, 2 * (bc + br))],
        (0, 1): [(lc, z, z, lc), (z, -lc, z, lc), (vc, -vr, z, vc + vr), (z, z, z, z), (z, z, z, z)],
        (0, 2): [(bc, br, z, bc + br), (z, -bc, br, br + bc), (z, z, z, z), (vc, -vr, z, vc + vr), (z, vc, vr, vr + vc)],
        (1, 2): [(z, z, z, z), (z, z, z, z), (bc, br, z, bc + br), (lc, z, z, lc), (z, lc, z, lc)],
    }


def biggp_matrix_grads(kname, a, t, jitter=None):
    """List of dK/dtheta (3N x 3N), reference block order, for theta = a (+ jitter)."""
    a = np.asarray(a, dtype=float)
    t = np.asarray(t, dtype=float)
    n = t.size
    kp = go.kernel_pars(kname, a)
    wn = kp[-1]
    nshape = len(kp) - 1
    r = t[:, None] - t[None, :]
    (E, AE, Gdd), dE, dAE, dGdd = closed_bases_grad(kname, r, kp)
    coefs = go.biggp_coefs(a[-5:])
    cgrad = _coef_grads(a[-5:])
    eye = np.identity(n)

    def assemble(block):
        K = np.zeros((3 * n, 3 * n))
        for (p, q) in coefs:
            blk = block(p, q)
            K[p * n:(p + 1) * n, q * n:(q + 1) * n] = blk
            if p != q:
                K[q * n:(q + 1) * n, p * n:(p + 1) * n] = blk.T
        return K

    out = []
    for k in range(nshape):                                   # kernel shape parameters
        out.append(assemble(lambda p, q: coefs[(p, q)][0] * dE[k] + coefs[(p, q)][2] * dGdd[k] +
                            coefs[(p, q)][1] * dAE[k]))
    out.append(assemble(lambda p, q: coefs[(p, q)][5] * 2.0 * wn * eye))          # wn (quirk Q1: every block)
    for m in range(5):                                        # vc, vr, lc, bc, br
        out.append(assemble(lambda p, q: cgrad[(p, q)][m][0] * E + cgrad[(p, q)][m][2] * Gdd +
                            cgrad[(p, q)][m][1] * AE + cgrad[(p, q)][m][3] * wn ** 2 * eye))
    if jitter is not None:
        for p in range(3):
            d = np.zeros(3 * n)
            d[p * n:(p + 1) * n] = 2.0 * float(jitter[p])
            out.append(np.diag(d))
    return out


def biggp_loglike_grad(kname, a, t, resid, rverr, sig_rhk, sig_bis, jitter=None):
    """(ll, gradient) with the gradient ordered as [a..., jitter...]."""
    K = go.biggp_matrix_closed(kname, a, t, rverr, sig_rhk, sig_bis, jitter=jitter)
    resid = np.asarray(resid, dtype=float)
    try:
        cf = sla.cho_factor(K, lower=False)
    except np.linalg.LinAlgError:
        return -np.inf, None
    alpha = sla.cho_solve(cf, resid)
    ll = -0.5 * resid @ alpha - np.sum(np.log(np.diag(cf[0]))) - 0.5 * resid.size * np.log(2 * pi)
    Kinv = sla.cho_solve(cf, np.identity(K.shape[0]))
    W = Kinv - np.outer(alpha, alpha)
    grad = np.array([-0.5 * np.sum(W * dK) for dK in biggp_matrix_grads(kname, a, t, jitter)])
    return float(ll), grad


def finite_difference_grad(kname, a, t, resid, rverr, sig_rhk, sig_bis, jitter=None, rel=1e-6):
    """Central differences of the reference-literal log-likelihood (the pin of the analytic form)."""
    theta = np.concatenate([np.asarray(a, dtype=float), [] if jitter is None else np.asarray(jitter, dtype=float)])
    na = len(a)
    out = np.zeros(theta.size)
    for k in range(theta.size):
        h = rel * max(abs(theta[k]), 1e-3)
        vals = []
        for sgn in (1.0, -1.0):
            th = theta.copy()
            th[k] += sgn * h
            vals.append(go.biggp_loglike_literal(kname, th[:na], t, resid, rverr, sig_rhk, sig_bis,
                                                 jitter=None if jitter is None else th[na:]))
        out[k] = (vals[0] - vals[1]) / (2 * h)
    return out

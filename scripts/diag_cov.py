"""GPU diagnostic: per golden case, where the covariance differs from the reference."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden
import miniframe_b200 as mf
from test_gpu_parity import _make_gp, _matrix, _loglike
from oracle import gp_oracle as go

g = Golden()
for case in g.meta["cases"]:
    Kref = g.K(case)
    gp = _make_gp(mf, g, case)
    ll = _loglike(gp, case)
    msg = "%-20s ll=%.12g ref=%.12g rel=%.2e" % (case["name"], ll, case["ll"],
          abs(ll - case["ll"]) / abs(case["ll"]) if np.isfinite(case["ll"]) and np.isfinite(ll) else float(ll != case["ll"]))
    if Kref is not None:
        K = _matrix(gp, case)
        d = np.abs(K - Kref); kmax = np.abs(Kref).max()
        tol = 1e-12 * np.maximum(np.abs(Kref), 1e-6 * kmax)
        ratio = d / tol
        idx = np.dstack(np.unravel_index(np.argsort(-d.ravel())[:4], d.shape))[0]
        asym = np.argwhere(K != K.T)
        msg += " | maxnorm=%.2e worst=%.2f nviol=%d asym=%d top=%s" % (
            d.max() / kmax, ratio.max(), int((ratio > 1).sum()), len(asym),
            [(int(i), int(j), float("%.3g" % d[i, j]), float("%.3g" % Kref[i, j])) for i, j in idx])
        if len(asym):
            i, j = asym[0]
            msg += " asym0=(%d,%d,%.3e)" % (i, j, K[i, j] - K[j, i])
    print(msg, flush=True)

# symmetry / parity at several sizes
for N in (21, 64, 128, 150):
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=N)
    H = go.synthetic_biggp_hypers(5, seed=N)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    K = gp.compute_matrix(H[0]); Kref = go.biggp_matrix_literal("QuasiPeriodic", H[0], t, e1, e2, e3)
    d = np.abs(K - Kref); kmax = np.abs(Kref).max()
    asym = np.argwhere(K != K.T)
    rel = d / np.maximum(np.abs(Kref), 1e-300)
    big = np.argwhere((d > 1e-12 * np.abs(Kref)) & (d > 1e-18 * kmax))
    print("N=%d maxnorm=%.2e asym=%d maxasym=%.3e n(rel>1e-12)=%d" % (
        N, d.max() / kmax, len(asym), np.abs(K - K.T).max(), len(big)))
    for i, j in big[:6]:
        print("   (%d,%d) K=%.6e ref=%.6e d=%.2e  scale~%.2e" % (i, j, K[i, j], Kref[i, j], d[i, j], kmax))
    ll = gp.log_likelihood_batch(H); y = go.biggp_y(rv, rhk, bis)
    want = np.array([go.biggp_loglike_literal("QuasiPeriodic", a, t, y, e1, e2, e3) for a in H])
    print("   ll rel err", np.abs(ll - want) / np.abs(want))

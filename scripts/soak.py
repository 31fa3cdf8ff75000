"""Soak test of the factorisation path: many matrices in flight, several sizes and batch sizes
(one CTA per matrix and cooperative groups), repeated; every repetition must reproduce the first
bit for bit and a few samples are checked against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200.BIGgpPlusJitt import BIGgp as BIGgpJ
from oracle import gp_oracle as go

bad = 0
for N, B, reps in [(100, 2368, 6), (333, 1184, 6), (500, 1776, 6), (500, 100, 10), (500, 7, 10), (200, 9, 30), (200, 1, 30),
                   (700, 592, 4), (1000, 40, 4), (2000, 3, 4), (43, 4096, 6), (21, 50, 20)]:
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=N)
    H = go.synthetic_biggp_hypers(B, seed=B)
    J = go.synthetic_jitter(B, seed=B)
    for cls, extra in ((BIGgp, ()), (BIGgpJ, (torch.as_tensor(J, device="cuda"),))):
        gp = cls(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
        Hd = torch.as_tensor(H, device="cuda")
        args = (Hd, None) + extra if extra else (Hd,)
        ref = gp.log_likelihood_batch(*args).cpu().numpy()
        diff = 0
        for _ in range(reps):
            cur = gp.log_likelihood_batch(*args).cpu().numpy()
            diff += int((cur != ref).sum())
        y = go.biggp_y(rv, rhk, bis)
        worst = 0.0
        for i in (0, B // 2, B - 1):
            want = go.loglike_from_matrix(go.biggp_matrix_closed("QuasiPeriodic", H[i], t, e1, e2, e3, jitter=(J[i] if extra else None)), y)
            worst = max(worst, abs(ref[i] - want) / abs(want))
        bad += diff + (worst > 1e-9)
        print("N %4d B %4d %-6s: differing over %d repetitions %d, finite %d/%d, worst rel err vs oracle %.2e" % (
            N, B, "jitter" if extra else "plain", reps, diff, int(np.isfinite(ref).sum()), B, worst), flush=True)
print("SOAK", "FAILED" if bad else "OK")

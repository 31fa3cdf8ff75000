import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from oracle import gp_oracle as go
for N in (5, 21, 22, 43, 64, 70, 128):
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=11)
    H = go.synthetic_biggp_hypers(2, seed=11)
    gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    ll = gp.log_likelihood_batch(H)
    y = go.biggp_y(rv, rhk, bis)
    want = np.array([go.biggp_loglike_literal("QuasiPeriodic", a, t, y, e1, e2, e3) for a in H])
    print("N", N, "n", 3 * N, "got", ll, "want", want, flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from oracle import gp_oracle as go
N, B = 500, int(os.environ.get("RB", "640"))
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(8, seed=1)
A = np.tile(H, (B // 8, 1))
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
ref = gp.log_likelihood_batch(A)
bad = 0
for rep in range(6):
    cur = gp.log_likelihood_batch(A)
    d = np.abs(cur - ref) / np.abs(ref)
    bad += int((cur != ref).sum())
    print("rep", rep, "differing", int((cur != ref).sum()), "max rel diff %.3e" % d.max(), "slots equal", bool(np.array_equal(cur.reshape(-1, 8), np.tile(cur[:8], (B // 8, 1)))), flush=True)
print("TOTAL differing", bad)

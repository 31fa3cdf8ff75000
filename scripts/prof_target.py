"""Small profiling target: one covariance build + one factorisation launch over
`B` matrices of order 3N (default 296 x 1500 = one resident wave of CTAs)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = Engine.get()
for kv in os.environ.get("MF_OPT", "").split():      # e.g. MF_OPT="potrf_lookahead=1"
    k, v = kv.split("="); eng.set_option(k, int(v))
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
for _ in range(reps):
    ll = gp.log_likelihood_batch(torch.as_tensor(H, device="cuda"))
torch.cuda.synchronize()
print("ok", float(ll[0]), int(torch.isfinite(ll).sum()))

"""Quick timing of small-batch / single-large-matrix cases (whole path), best of 3."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from oracle import gp_oracle as go
for N, B in [(200, 1), (500, 1), (500, 8), (500, 64), (2000, 1), (2000, 8), (10000, 1)]:
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
    H = torch.as_tensor(go.synthetic_biggp_hypers(B, seed=2), device="cuda")
    gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    for _ in range(2): ll = gp.log_likelihood_batch(H)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ll = gp.log_likelihood_batch(H); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    n = 3 * N
    print(json.dumps({"N": N, "B": B, "ms": best, "tflops": (n ** 3 / 3 + n ** 2) * B / best / 1e9, "ll0": float(ll[0])}), flush=True)

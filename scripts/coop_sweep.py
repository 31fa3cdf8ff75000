"""Factorisation time of a small batch as a function of the cooperative group size."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go
eng = Engine.get()
N = int(sys.argv[1]); Bs = [int(x) for x in sys.argv[2].split(",")]; Gs = [int(x) for x in sys.argv[3].split(",")]
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
prob = gp._problem(N); n = 3 * N
resid = eng.tensor(go.biggp_y(rv, rhk, bis)[None, :])
for B in Bs:
    H = eng.tensor(go.synthetic_biggp_hypers(B, seed=2))
    for G in Gs:
        eng.set_option("potrf_group", G)
        best = 1e9
        for rep in range(4):
            Kd = eng.build_cov(prob, eng.tensor(t), eng.tensor(gp._diag_yerr()), H, None, full=False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ll, info = eng.potrf_loglike(Kd, n, resid); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("N %d B %d G %d: potrf %.3f ms  %.2f TFLOP/s" % (N, B, G, best, (n ** 3 / 3 + n ** 2) * B / best / 1e9), flush=True)

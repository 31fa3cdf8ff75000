import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go
eng = Engine.get()
N = 500
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
def tm(f, reps=4):
    f(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for B in (1, 2, 4, 8, 16, 64):
    H = torch.as_tensor(go.synthetic_biggp_hypers(B, seed=2), device="cuda")
    for G in ("1", None):
        if G: os.environ["MF_COOP_GROUP"] = G
        else: os.environ.pop("MF_COOP_GROUP", None)
        print("B %d G %s: log_likelihood_batch %.3f ms" % (B, G, tm(lambda: gp.log_likelihood_batch(H))), flush=True)

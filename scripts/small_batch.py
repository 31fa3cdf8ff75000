"""Small batches and few large matrices (BASELINE configs 2 and 5): the cooperative path
(several CTAs per matrix) against one CTA per matrix, and a parity check of the order-30000
matrix against the oracle's closed form."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from oracle import gp_oracle as go

def run(N, B, group=None, reps=3):
    from miniframe_b200._engine import Engine
    Engine.get().set_option("potrf_group", 0 if group is None else group)
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
    H = go.synthetic_biggp_hypers(B, seed=2)
    gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    Hd = torch.as_tensor(H, device="cuda")
    ll = gp.log_likelihood_batch(Hd); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ll = gp.log_likelihood_batch(Hd); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    n = 3 * N
    return best, (n ** 3 / 3 + n ** 2) * B / best / 1e9, ll.cpu().numpy(), (t, rv, rhk, bis, e1, e2, e3, H)

from miniframe_b200._engine import Engine
rows = []
for N, B in [(200, 1), (200, 9), (500, 1), (500, 8), (500, 64), (500, 148), (2000, 1), (2000, 8), (2000, 64)]:
    ms1, tf1, ll1, _ = run(N, B, group=1)
    row = {"N": N, "B": B, "one_cta_ms": ms1, "one_cta_tflops": tf1}
    for mp in (1, 2):
        Engine.get().set_option("potrf_map", mp)
        msg, tfg, llg, _ = run(N, B)
        row.update({"coop_ms_map%d" % mp: msg, "coop_tflops_map%d" % mp: tfg,
                    "bit_identical_map%d" % mp: bool(np.array_equal(llg, ll1))})
    Engine.get().set_option("potrf_map", 0)
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    for B in (1, 8):
        for mp in (1, 2):
            Engine.get().set_option("potrf_map", mp)
            ms, tf, ll, (t, rv, rhk, bis, e1, e2, e3, H) = run(10000, B, reps=2)
            row = {"N": 10000, "B": B, "map": mp, "coop_ms": ms, "coop_tflops": tf, "ll0": float(ll[0])}
            print(json.dumps(row), flush=True)

"""BASELINE config 4: BIGgpPlusJitt, N = 2000 epochs (n = 6000), 1024 hyper-parameter sets with
per-series jitter, one GPU (the workspace is processed in chunks); parity spot check vs the oracle."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgpPlusJitt import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=4)
J = go.synthetic_jitter(B, seed=4)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
eng = Engine.get()
Hd, Jd = torch.as_tensor(H, device="cuda"), torch.as_tensor(J, device="cuda")
ll = gp.log_likelihood_batch(Hd[:8], None, Jd[:8]); torch.cuda.synchronize()
t0 = time.perf_counter(); ll = gp.log_likelihood_batch(Hd, None, Jd); torch.cuda.synchronize(); dt = time.perf_counter() - t0
ll = ll.cpu().numpy()
n = 3 * N
out = {"config": "BIGgpPlusJitt QP N=%d (n=%d) batch %d, 1 GPU" % (N, n, B), "seconds": dt, "evals_per_s": B / dt,
       "tflops": (n ** 3 / 3 + n ** 2) * B / dt / 1e12, "finite": int(np.isfinite(ll).sum()),
       "workspace_gb": sum(w.numel() for w in eng._ws.values()) / 1e9}
y = go.biggp_y(rv, rhk, bis)
errs = []
for i in (0, B - 1):
    t1 = time.perf_counter()
    want = go.loglike_from_matrix(go.biggp_matrix_closed("QuasiPeriodic", H[i], t, e1, e2, e3, jitter=J[i]), y)
    errs.append(abs(ll[i] - want) / abs(want))
    out["oracle_s_per_eval"] = time.perf_counter() - t1
out["parity_rel_err"] = errs
print(json.dumps(out))

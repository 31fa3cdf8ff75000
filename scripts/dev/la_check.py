"""Look-ahead kernel against the plain one-CTA-per-matrix kernel: bit equality of ll / info / factor, timing.
    python scripts/dev/la_check.py [N epochs] [batch] [library file name]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import _lib
if len(sys.argv) > 3:
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), sys.argv[3])
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
n = 3 * N
eng = Engine.get()
lib, h = eng.lib, eng.handle
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
H[3::7, 0] *= 40.0          # some samples with a huge amplitude against tiny noise: not positive definite in FP64
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
prob = gp._problem(N)
t_d, diag_d, hyper_d = eng.tensor(t), eng.tensor(gp._diag_yerr() * 1e-3), eng.tensor(H)
resid_d = eng.tensor(gp.y[None, :])
ws, ws_bytes = eng.workspace(B, n)
ld, stride = eng.ld(n), eng.matrix_stride(n)
st = eng.stream()
row_n = ctypes.c_void_p(ws.data_ptr() + n * ld * 8)
out = {}
eng.set_option("potrf_group", 1)
eng.set_option("potrf_tile", int(os.environ.get("TILE", "2")))
for name, opt in (("plain", 1), ("lookahead", 2)):
    eng.set_option("potrf_lookahead", opt)
    ll = torch.empty(B, dtype=torch.float64, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
    ms = []
    for rep in range(3):
        rc = lib.mf_build_cov(h, ctypes.byref(prob), B, eng.ptr(t_d), eng.ptr(diag_d), eng.ptr(hyper_d), None, 2, eng.ptr(ws), ld, stride, st)
        assert rc == 0
        lib.mf_interleave_resid(h, B, 3, N, eng.ptr(resid_d), 0, row_n, stride, st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.mf_potrf_loglike(h, B, n, eng.ptr(ws), ld, stride, row_n, stride, eng.ptr(ll), eng.ptr(info), st)
        b.record(); torch.cuda.synchronize()
        assert rc == 0, lib.mf_last_error(h)
        ms.append(a.elapsed_time(b))
    W = ws.view(torch.float64)
    fac = W[: min(B, 64) * stride].clone()
    out[name] = (ll.clone(), info.clone(), fac, min(ms))
    print("%-10s ms %s  %.2f TFLOP/s  positive definite %d / %d" % (name, " ".join("%.2f" % m for m in ms), (n ** 3 / 3 + n ** 2) * B / min(ms) / 1e9, int((info == 0).sum()), B), flush=True)
a, b = out["plain"], out["lookahead"]
print("ll bit-identical:", bool(torch.equal(a[0], b[0])), " info identical:", bool(torch.equal(a[1], b[1])))
good = (a[1][: min(B, 64)] == 0).nonzero().flatten().tolist()
same = all(torch.equal(a[2][i * stride:(i + 1) * stride], b[2][i * stride:(i + 1) * stride]) for i in good)
print("factors of the first %d positive-definite matrices bit-identical: %s" % (len(good), same))
if not torch.equal(a[0], b[0]):
    d = (a[0] - b[0]).abs(); d[torch.isnan(d)] = 0
    print("max |dll|", float(d.max()), "mismatching", int((a[0] != b[0]).sum()))

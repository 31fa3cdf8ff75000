"""How many matrices each CTA took (experiment build, DBG 16384): python scripts/dev/handout.py [N] [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libminiframe_b200_exp.so")
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n = 3 * N
eng = Engine.get(); lib, h = eng.lib, eng.handle
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
prob = gp._problem(N)
t_d, diag_d, hyper_d = eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(H)
resid_d = eng.tensor(gp.y[None, :])
ll = torch.empty(B, dtype=torch.float64, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
ws, _ = eng.workspace(B, n); ld, stride = eng.ld(n), eng.matrix_stride(n); st = eng.stream()
row_n = ctypes.c_void_p(ws.data_ptr() + n * ld * 8)
eng.set_option("experiment", 16384)
for rep in range(2):
    lib.mf_build_cov(h, ctypes.byref(prob), B, eng.ptr(t_d), eng.ptr(diag_d), eng.ptr(hyper_d), None, 2, eng.ptr(ws), ld, stride, st)
    lib.mf_interleave_resid(h, B, 3, N, eng.ptr(resid_d), 0, row_n, stride, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rc = lib.mf_potrf_loglike(h, B, n, eng.ptr(ws), ld, stride, row_n, stride, eng.ptr(ll), eng.ptr(info), st); b.record()
    torch.cuda.synchronize(); assert rc == 0
who = info.cpu().numpy()
cnt = np.bincount(who, minlength=296)
print("n %d B %d: %.2f ms; matrices per CTA: min %d max %d mean %.2f; histogram %s" % (n, B, a.elapsed_time(b), cnt.min(), cnt.max(), cnt.mean(), np.bincount(cnt).tolist()))
lo, hi = cnt[:148], cnt[148:]
print("CTAs 0-147 took %d, CTAs 148-295 took %d; same-SM pairs (i, i+148) if the block scheduler fills SMs round-robin: corr %.3f" % (lo.sum(), hi.sum(), np.corrcoef(lo, hi)[0, 1]))
print("first 16 CTAs:", cnt[:16].tolist(), " CTAs 148..163:", cnt[148:164].tolist())

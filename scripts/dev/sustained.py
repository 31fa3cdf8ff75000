"""Factorisation under sustained load (the bench's regime: 4096 matrices per launch, power cap):
    python scripts/dev/sustained.py [name=value ...]   options go to mf_set_option before the run"""
import ctypes, os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from miniframe_b200 import _lib
if os.environ.get("MFLIB"): _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ["MFLIB"])
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go
N, B = int(os.environ.get("NEP", "500")), int(os.environ.get("NB", "4096"))
REPS = int(os.environ.get("REPS", "10"))
n = 3 * N
eng = Engine.get()
lib, h = eng.lib, eng.handle
for kv in sys.argv[1:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
prob = gp._problem(N)
t_d, diag_d, hyper_d = eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(H)
resid_d = eng.tensor(gp.y[None, :])
ll = torch.empty(B, dtype=torch.float64, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
ws, ws_bytes = eng.workspace(B, n)
assert ws_bytes >= B * eng.matrix_stride(n) * 8, "batch does not fit the workspace"
ld, stride = eng.ld(n), eng.matrix_stride(n)
st = eng.stream()
row_n = ctypes.c_void_p(ws.data_ptr() + n * ld * 8)
samples = []
stop = False
def sampler():
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            samples.append((float(out[0]), float(out[1])))
        except Exception:
            pass
        time.sleep(0.1)
th = threading.Thread(target=sampler, daemon=True); th.start()
ms = []
for rep in range(REPS):
    lib.mf_build_cov(h, ctypes.byref(prob), B, eng.ptr(t_d), eng.ptr(diag_d), eng.ptr(hyper_d), None, 2, eng.ptr(ws), ld, stride, st)
    lib.mf_interleave_resid(h, B, 3, N, eng.ptr(resid_d), 0, row_n, stride, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rc = lib.mf_potrf_loglike(h, B, n, eng.ptr(ws), ld, stride, row_n, stride, eng.ptr(ll), eng.ptr(info), st); b.record()
    assert rc == 0
    ms.append((a, b))
torch.cuda.synchronize(); stop = True; th.join()
tms = [a.elapsed_time(b) for a, b in ms]
clk = sorted(s[0] for s in samples); pw = [s[1] for s in samples]
print("%s: ms %s | second half mean %.2f = %.2f TFLOP/s | clocks median %s min %s, power max %.0f W" % (
    " ".join(sys.argv[1:]) or "default", " ".join("%.1f" % m for m in tms), np.mean(tms[REPS // 2:]),
    (n ** 3 / 3 + n ** 2) * B / np.mean(tms[REPS // 2:]) / 1e9, clk[len(clk) // 2] if clk else None, clk[0] if clk else None, max(pw) if pw else -1))

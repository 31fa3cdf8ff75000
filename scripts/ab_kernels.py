"""A/B timing of the two kernels through the C ABI, on one GPU.

    python scripts/ab_kernels.py [--lib NAME] [--batch B] [--epochs N] [--exp 0,32,...] [--cov]

--lib picks the shared library in miniframe_b200/csrc (default libminiframe_b200.so; the
`_exp` / `_old` builds carry the timing-experiment instantiations, `make -C miniframe_b200/csrc exp old`).
Experiment codes are the DBG bit masks of mf_potrf.cu (results of those launches are garbage).
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="libminiframe_b200.so")
ap.add_argument("--batch", type=int, default=2368)
ap.add_argument("--epochs", type=int, default=500)
ap.add_argument("--exp", default="0")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cov", action="store_true")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--opt", action="append", default=[], help="name=value for mf_set_option")
args = ap.parse_args()

from miniframe_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), args.lib)
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200._engine import Engine
from oracle import gp_oracle as go

eng = Engine.get()
lib, h = eng.lib, eng.handle
N, B = args.epochs, args.batch
n = 3 * N
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
prob = gp._problem(N)
t_d, diag_d, hyper_d = eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(H)
resid_d = eng.tensor(gp.y[None, :])
ll = torch.empty(B, dtype=torch.float64, device="cuda")
info = torch.empty(B, dtype=torch.int32, device="cuda")
ws, ws_bytes = eng.workspace(B, n)
ld, stride = eng.ld(n), eng.matrix_stride(n)
assert ws_bytes >= B * stride * 8, "batch does not fit one chunk"
st = eng.stream()
row_n = ctypes.c_void_p(ws.data_ptr() + n * ld * 8)
flops = (n ** 3 / 3.0 + n ** 2) * B
bytes_ = 4.0 * n * (n + 1) * B


def build(mode):
    rc = lib.mf_build_cov(h, ctypes.byref(prob), B, eng.ptr(t_d), eng.ptr(diag_d), eng.ptr(hyper_d), None, mode,
                          eng.ptr(ws), ld, stride, st)
    assert rc == 0, lib.mf_last_error(h)


def potrf():
    rc = lib.mf_interleave_resid(h, B, 3, N, eng.ptr(resid_d), 0, row_n, stride, st)
    assert rc == 0
    rc = lib.mf_potrf_loglike(h, B, n, eng.ptr(ws), ld, stride, row_n, stride, eng.ptr(ll), eng.ptr(info), st)
    assert rc == 0, lib.mf_last_error(h)


def timed(fn, reps):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out


if args.tile:
    eng.set_option("potrf_tile", args.tile)
for kv in args.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
build(2); potrf(); torch.cuda.synchronize()
ref = ll.clone()
print("lib %s  n %d  batch %d  finite %d/%d" % (args.lib, n, B, int(torch.isfinite(ref).sum()), B), flush=True)
for code in [int(x) for x in args.exp.split(",")]:
    rc = lib.mf_set_option(h, b"experiment", code)
    if rc != 0:
        print("experiment %d: %s" % (code, lib.mf_last_error(h).decode()))
        continue
    ms = []
    for _ in range(args.reps):
        build(2)
        ms += timed(potrf, 1)
    same = bool(torch.equal(ll, ref)) if code == 0 else None
    print("potrf exp %4d: ms %s  best %.2f  %.2f TFLOP/s (nominal flops)%s"
          % (code, " ".join("%.2f" % m for m in ms), min(ms), flops / min(ms) / 1e9,
             "" if same is None else "  bit-reproducible %s" % same), flush=True)
lib.mf_set_option(h, b"experiment", 0)
if args.cov:
    for mode, name in ((0, "reference order, 9 streams"), (2, "epoch-interleaved, TMA tiles")):
        for ctas in (0, 4, 8, 17, 34, 68, 136):
            eng.set_option("cov_ctas", ctas)
            ms = timed(lambda: build(mode), args.reps + 1)[1:]
            print("cov mode %d (%s) ctas/sample %3d: ms %s  best %.3f  %.0f GB/s"
                  % (mode, name, ctas, " ".join("%.3f" % m for m in ms), min(ms), bytes_ / min(ms) / 1e6), flush=True)
    eng.set_option("cov_ctas", 0)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MF_DBG"] = "16"
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from oracle import gp_oracle as go
N, B = 500, 592
t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
H = go.synthetic_biggp_hypers(B, seed=1)
gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
ll = gp.log_likelihood_batch(torch.as_tensor(H, device="cuda"))
torch.cuda.synchronize()
tr = np.fromfile("gpurun_out/potrf_trace.bin", dtype=np.int64).reshape(-1, 200, 8)
print(tr.shape)
# analyse co-residency on a few SMs
sm = tr[:, 0, 0]
from collections import defaultdict
by = defaultdict(list)
for b in range(tr.shape[0]): by[int(sm[b])].append(b)
print("CTAs per SM histogram:", np.bincount([len(v) for v in by.values()]))
print("example co-resident pairs:", [v for v in list(by.values())[:6]])
tot_idle = []; 
for s, ctas in list(by.items()):
    if len(ctas) != 2: continue
    ev = []
    for b in ctas:
        rows = tr[b]; rows = rows[rows[:, 2] > 0]
        # main loop intervals [t3, t4]
        ev.append([(int(r[3]), int(r[4])) for r in rows if r[4] > r[3]])
    t0 = min(e[0][0] for e in ev); t1 = max(e[-1][1] for e in ev)
    # union coverage of main-loop intervals
    allint = sorted(ev[0] + ev[1])
    cov = 0; cur_s, cur_e = allint[0]
    for a, b in allint[1:]:
        if a > cur_e: cov += cur_e - cur_s; cur_s, cur_e = a, b
        else: cur_e = max(cur_e, b)
    cov += cur_e - cur_s
    each = [sum(b - a for a, b in e) for e in ev]
    tot_idle.append((1 - cov / (t1 - t0), each[0] / (t1 - t0), each[1] / (t1 - t0)))
ti = np.array(tot_idle)
print("fraction of time NO CTA is in its contraction loop: mean %.3f; per-CTA loop fraction mean %.3f" % (ti[:, 0].mean(), ti[:, 1:].mean()))
b = list(by.values())[0]
for c in b:
    rows = tr[c]; rows = rows[rows[:, 2] > 0]
    base = rows[0, 2]
    print("CTA", c, "first tiles (jb*1000+it, start, loop_start, loop_end, potf2/prefetch end) in kclk:")
    for r in rows[:14]:
        print("   ", int(r[1]), [round((int(x) - base) / 1e3, 1) for x in r[2:6]])
    dur = rows[:, 4] - rows[:, 3]; tile = np.diff(rows[:, 2])
    print("   sum loop %.0f kclk, total %.0f kclk, tiles %d" % (dur.sum() / 1e3, (rows[-1, 4] - base) / 1e3, len(rows)))

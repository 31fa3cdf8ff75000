"""MCMC-style use of the hot path: the reference example's sampler set-up
(examples/BIGgp_mcmc.py: 2 x 9 walkers, QuasiPeriodic BIGgp, log-uniform box) run device-resident,
plus a larger ensemble; reports steps/s and likelihood evaluations/s."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniframe_b200 import kernels
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200.sampler import EnsembleSampler, box_log_prob
from oracle import gp_oracle as go

lo = np.log([5.0, 0.5, 15.0, 1e-3, 0.1, 0.1, 0.1, 0.1, 0.1])
hi = np.log([100.0, 2.0, 35.0, 0.1, 10.0, 10.0, 1.0, 10.0, 10.0])
for N, walkers, steps in [(200, 18, 400), (200, 256, 100), (500, 64, 50), (500, 592, 20)]:
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
    gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    p0 = np.random.default_rng(2).uniform(lo, hi, size=(walkers, 9))
    for graph in (True, False):
        s = EnsembleSampler(walkers, 9, box_log_prob(gp, lo, hi), seed=1, device="cuda", use_graph=graph)
        p, lnp, _ = s.run_mcmc(p0, 5)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p, lnp, _ = s.run_mcmc(p, steps)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"N": N, "n": 3 * N, "walkers": walkers, "steps": steps, "cuda_graph": bool(s._graph),
                          "graph_error": getattr(s, "graph_error", None), "steps_per_s": steps / dt,
                          "evals_per_s": steps * walkers / dt, "acceptance": float(s.acceptance_fraction.mean()),
                          "best_lnprob": float(lnp.max())}), flush=True)

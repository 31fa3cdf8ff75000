"""BASELINE configs 4 and 5b on the GPUs of one box, through the product's sharded entry point
(miniframe_b200.dist.log_likelihood_batch_sharded; one process per GPU, NCCL all-gather of the
per-sample log-likelihoods):

  config 4:  BIGgpPlusJitt, N = 2000 epochs (n = 6000), 1024 hyper-parameter sets with jitter
  config 5b: BIGgp, N = 10000 epochs (n = 30000), 64 hyper-parameter sets

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/configs_multi_gpu.py

Rank 0 prints one JSON line per config (device time, max over ranks) and checks two samples of
config 4 against the oracle's literal port of the reference.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

from miniframe_b200 import kernels, dist as mfdist
from miniframe_b200.BIGgp import BIGgp
from miniframe_b200.BIGgpPlusJitt import BIGgp as BIGgpJ
from oracle import gp_oracle as go


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


def run(name, cls, N, B, jitter, reps):
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=4)
    gp = cls(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    H = go.synthetic_biggp_hypers(B, seed=4)
    A = torch.as_tensor(H, device=dev)
    C = torch.as_tensor(go.synthetic_jitter(B, seed=4), device=dev) if jitter else None
    if world > 1:
        fn = lambda: mfdist.log_likelihood_batch_sharded(gp, A, None, C)
    else:
        fn = (lambda: gp.log_likelihood_batch(A, None, C)) if jitter else (lambda: gp.log_likelihood_batch(A))
    ms, ll = timed(fn, reps)
    n = 3 * N
    row = {"config": name, "n_gpus": world, "matrix_order": n, "batch": B, "per_gpu": B // world, "ms": ms,
           "evals_per_s": B / ms * 1e3, "tflops_total": (n ** 3 / 3 + n ** 2) * B / ms / 1e9,
           "finite": int(torch.isfinite(ll).sum()), "ll0": float(ll[0])}
    if rank == 0 and jitter:
        y = go.biggp_y(rv, rhk, bis)
        J = go.synthetic_jitter(B, seed=4)
        errs = []
        for i in (0, B - 1):
            want = go.biggp_loglike_literal("QuasiPeriodic", H[i], t, y, e1, e2, e3, jitter=J[i])
            errs.append(abs(float(ll[i]) - want) / abs(want))
        row["parity_vs_oracle_max_rel"] = max(errs)
    if rank == 0:
        print(json.dumps(row), flush=True)
    if world > 1: dist.barrier()
    from miniframe_b200._engine import Engine
    Engine.get().drop_workspaces(); torch.cuda.empty_cache()


which = sys.argv[1:] or ["4", "5b"]
if "4" in which:
    run("config 4: BIGgpPlusJitt N=2000 (n=6000) x 1024", BIGgpJ, 2000, 1024, True, 2)
if "5b" in which:
    run("config 5b: BIGgp N=10000 (n=30000) x 64", BIGgp, 10000, 64, False, 1)
if world > 1:
    dist.destroy_process_group()

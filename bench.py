#!/usr/bin/env python3
"""Headline benchmark: GP log-likelihood evaluations per second, FP64, n = 3N = 1500
(BASELINE.json configs[2]: BIGgp_mcmc-style batch of 4096 hyper-parameter sets x
N = 500 epochs, QuasiPeriodic kernel), on N GPUs of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (covariance build -> Cholesky -> solve ->
log-det -> log-likelihood) over one batch of 4096 synthetic hyper-parameter sets
per GPU (weak scaling: every rank evaluates its own 4096; the per-sample results
are all-gathered over NCCL inside the timed region).  With N > 1 the line also
carries "strong": the global batch of BASELINE configs[2] (4096, and 32768) cut
across the N ranks by the product's own sharded entry point
(miniframe_b200.dist.log_likelihood_batch_sharded), with rank 0 evaluating the
same batch alone in the same run as the 1-GPU reference.  One JSON line on stdout.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import hashlib

# dram__bytes_read.sum + dram__bytes_write.sum per launch come from a tracked ncu capture
# (profiles/ncu_traffic.json, written by scripts/ncu_to_traffic.py from the raw export of an
# `ncu --set full` run).  Each entry records the sha256 of the kernel's source file at capture
# time; when the source has changed since, the figure is stale and `traffic` is reported as null.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def ncu_traffic(kernel_key, n, evals):
    """(GB per launch of `evals` matrices, note) or (None, why)."""
    try:
        with open(TRAFFIC_FILE) as f:
            table = json.load(f)
        ent = table[kernel_key]
        with open(os.path.join(ROOT, ent["source"]), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        if sha != ent["source_sha256"]:
            return None, "stale: %s changed since the ncu capture %s" % (ent["source"], ent["capture"])
        if int(ent["matrix_order"]) != int(n):
            return None, "captured at n = %d, this run is n = %d" % (ent["matrix_order"], n)
        per_eval = (ent["dram_bytes_read"] + ent["dram_bytes_write"]) / ent["matrices"]
        return per_eval * evals / 1e9, ("GB per launch: ncu dram read+write of a %d-matrix launch (%s), "
                                        "scaled to this launch" % (ent["matrices"], ent["capture"]))
    except Exception as e:                      # no capture tracked yet
        return None, "no tracked ncu capture (%s)" % type(e).__name__

METRIC = "gp_loglike_evals_per_s"
UNIT = "evals/s"


def flops_per_eval(n):
    return n ** 3 / 3.0 + n ** 2           # potrf + one triangular solve


def build_bytes_per_eval(n):
    return 4.0 * n * (n + 1)               # lower triangle incl. diagonal, FP64, written once


def workload(args, rank):
    from oracle import gp_oracle as go      # synthetic input generator only (seeded draws)
    N = args.epochs
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
    H = go.synthetic_biggp_hypers(args.batch, seed=1 + rank)
    return N, t, (rv, rhk, bis), (e1, e2, e3), H


def config_dict(args, world):
    return {"workload": "BIGgp QuasiPeriodic, N=%d epochs (n=%d), batch %d hyper-parameter sets per GPU"
                        % (args.epochs, 3 * args.epochs, args.batch),
            "framework": "BIGgp", "kernel": "QuasiPeriodic", "n_epochs": args.epochs,
            "matrix_order": 3 * args.epochs, "per_gpu_batch": args.batch,
            "global_batch": args.batch * world, "sharding": "batch over ranks, allgather of ll",
            "l2_policy": "inputs larger than L2 (%.1f GB of matrices per step per GPU)"
                         % (args.batch * ((3 * args.epochs + 8) // 8 * 8) * ((3 * args.epochs + 15) // 16 * 16) * 8 / 1e9)}


class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="mf_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        # nvidia-smi needs up to a second before its first sample: wait for it, so that short
        # regions are not missed
        t_end = time.time() + 3.0
        while self.proc is not None and time.time() < t_end:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.out.close()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[4:8]):
                    if val == "Active":
                        reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_WORK = {}


def reference_kind():
    """"reference": the unmodified reference classes can be imported (from /root/reference in the
    build container, else from the copy oracle/make_ref.py staged under oracle/_ref/);
    "port": only the oracle's literal port of them is available."""
    from oracle import ref_shim
    return "reference" if ref_shim.available() else "port"


def _worker_init(N, seed):
    """Pool worker: one BLAS thread each (the pool supplies the parallelism, as emcee's
    `threads=` pool does in the reference, examples/BIGgp_mcmc.py:88)."""
    from oracle import gp_oracle as go          # loads numpy's and scipy's BLAS ...
    from oracle import ref_shim
    try:
        from threadpoolctl import threadpool_limits
        _WORK["limit"] = threadpool_limits(limits=1)      # ... which are then pinned to 1 thread
    except Exception:
        pass
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=seed)
    _WORK.update(go=go, t=t, y=go.biggp_y(rv, rhk, bis), errs=(e1, e2, e3), gp=None)
    if ref_shim.available():
        ref = ref_shim.load()                   # the reference's own BIGgp, unmodified
        _WORK["gp"] = ref.BIGgp(ref.kernels.QuasiPeriodic, [None, None, None], ref_shim.as_time(t),
                                rv, e1, rhk, e2, bis, e3)


def _worker_eval(rows):
    if _WORK["gp"] is not None:
        return [float(_WORK["gp"].log_likelihood(list(a), [])) for a in rows]
    go, t, y, (e1, e2, e3) = _WORK["go"], _WORK["t"], _WORK["y"], _WORK["errs"]
    return [go.biggp_loglike_literal("QuasiPeriodic", a, t, y, e1, e2, e3) for a in rows]


class CpuPort(object):
    """The reference's CPU path (17 N x N kernel-class evaluations + scipy cho_factor / cho_solve
    per likelihood) on every host core: a pool of single-threaded workers, one likelihood per
    task.  The workers call the reference's own BIGgp.log_likelihood when its package is
    present (reference_kind() == "reference"), else the oracle's literal port of it."""

    def __init__(self, N, seed=1):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        self.kind = reference_kind()
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_worker_init, initargs=(N, seed))

    def evaluate(self, H):
        chunks = [H[i:i + 1] for i in range(len(H))]
        out = self.pool.map(_worker_eval, chunks)
        return np.array([v for part in out for v in part])

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_port_evals(N, H, seconds, max_evals):
    """Returns (evals/s, evals done, values, cores, kind) for a bounded sample of H."""
    port = CpuPort(N)
    try:
        port.evaluate(H[:port.cores])                      # worker start-up, not timed
        t0 = time.perf_counter()
        port.evaluate(H[:port.cores])
        per_round = time.perf_counter() - t0
        rounds = max(1, min(int(seconds / max(per_round, 1e-3)), max_evals // port.cores))
        sample = H[:rounds * port.cores]
        t0 = time.perf_counter()
        vals = port.evaluate(sample)
        dt = time.perf_counter() - t0
        return len(vals) / dt, len(vals), vals, port.cores, port.kind
    finally:
        port.close()


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host
    cores, a bounded sample of the same workload per step: the unmodified reference classes
    (staged under oracle/_ref/ by oracle/make_ref.py) when present, else the oracle's literal
    port of them (oracle/gp_oracle.py)."""
    if rank != 0:
        return
    N, t, ys, errs, H = workload(args, 0)
    port = CpuPort(N)
    try:
        sample = args.ref_sample * port.cores
        port.evaluate(H[:port.cores])                      # worker start-up

        def step(k):
            lo = (k * sample) % max(1, len(H) - sample)
            port.evaluate(H[lo:lo + sample])

        for k in range(args.warmup):
            step(k)
        t0 = time.perf_counter()
        for k in range(args.steps):
            step(args.warmup + k)
        dt = time.perf_counter() - t0
        cores, kind = port.cores, port.kind
    finally:
        port.close()
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d of the %d hyper-parameter sets per step (pool of %d "
                                       "single-threaded workers calling %s, numpy %s + scipy LAPACK)"
                                       % (sample, args.batch, cores,
                                          "the reference's BIGgp.log_likelihood" if kind == "reference"
                                          else "the oracle's literal port", np.__version__)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args, rank, local_rank, world):
    import ctypes

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path; no GPU is visible")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; keep stdout
        # clean for the one JSON line by pointing fd 1 at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from miniframe_b200 import _lib as mflib
    from miniframe_b200 import dist as mfdist
    from miniframe_b200 import kernels
    from miniframe_b200.BIGgp import BIGgp
    from miniframe_b200._engine import Engine
    from oracle import gp_oracle as go      # synthetic input generator only (seeded draws)
    eng = Engine.get()
    lib = eng.lib

    N, t, (rv, rhk, bis), (e1, e2, e3), H = workload(args, rank)
    n, B = 3 * N, args.batch
    gp = BIGgp(kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk,
               sig_rhk=e2, bis=bis, sig_bis=e3)
    prob = gp._problem(N)
    t_d = eng.tensor(t)
    diag_d = eng.tensor(gp._diag_yerr())
    hyper_d = eng.tensor(H)
    resid_d = eng.tensor(gp.y[None, :])
    ll_d = torch.empty(B, dtype=torch.float64, device=dev)
    info_d = torch.empty(B, dtype=torch.int32, device=dev)
    ll_all = torch.empty(B * world, dtype=torch.float64, device=dev)
    ws, ws_bytes = eng.workspace(B, n)
    ld, stride = eng.ld(n), eng.matrix_stride(n)
    chunk = max(1, min(B, ws_bytes // (stride * 8)))
    stream = eng.stream()
    handle = eng.handle

    # ---- FP64 roofline denominators, measured here --------------------------------
    peaks = {}
    if rank == 0:
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        best = 1e9
        for _ in range(5):
            e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1_.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1_))
        peaks["cublas_dgemm_4096_tflops"] = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
        # the register-resident DMMA probe (mma.sync.m8n8k4.f64, 16 independent accumulator tiles per
        # warp): ~5 ms bursts at 32 warps per SM, and ~0.5 s launches at 16 warps per SM - the
        # occupancy of the factorisation kernel - with clocks and power sampled during them.  (At
        # 32 warps per SM the probe itself falls to 29.7 TFLOP/s after ~10 ms; at 16 it holds the
        # full rate for seconds at ~430 W: profiles/r2_fp64_probes.txt.  The pipe's sustained peak is
        # the burst peak.)
        peaks["dmma_probe_tflops"] = max(eng.probe_fp64(0, 4000) for _ in range(3))
        peaks["dfma_probe_tflops"] = max(eng.probe_fp64(1, 4000) for _ in range(3))
        probe_clocks = ClockSampler(local_rank)
        peaks["dmma_probe_sustained_tflops"] = eng.probe_fp64(30, 400000)    # two launches of ~0.45 s
        pc = probe_clocks.stop()
        peaks["dmma_probe_sustained_clocks"] = {"sm_mhz_median": pc["sm_mhz"], "sm_mhz_min": pc.get("sm_mhz_min"),
                                                "power_w_max": pc.get("power_w_max"), "reasons": pc["reasons"],
                                                "samples": pc["samples"]}
        del a, b

    launches = [0]
    ptr = ctypes.c_void_p

    def step(events=None):
        """One pass over this rank's batch through the C ABI, inputs resident in HBM: covariance
        build (epoch-interleaved order), residual rows, factorisation + solve + log-det."""
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if events is not None else None
            if ev:
                ev[0].record()
            st = lib.mf_build_cov(handle, ctypes.byref(prob), nb, eng.ptr(t_d), eng.ptr(diag_d),
                                  ptr(hyper_d.data_ptr() + b0 * prob.hyper_len * 8), None,
                                  mflib.COV_LOWER_INTERLEAVED, eng.ptr(ws), ld, stride, stream)
            assert st == 0, lib.mf_last_error(handle)
            if ev:
                ev[1].record()
            row_n = ptr(ws.data_ptr() + n * ld * 8)
            st = lib.mf_interleave_resid(handle, nb, 3, N, eng.ptr(resid_d), 0, row_n, stride, stream)
            assert st == 0, lib.mf_last_error(handle)
            st = lib.mf_potrf_loglike(handle, nb, n, eng.ptr(ws), ld, stride, row_n, stride,
                                      ptr(ll_d.data_ptr() + b0 * 8), ptr(info_d.data_ptr() + b0 * 4), stream)
            assert st == 0, lib.mf_last_error(handle)
            if ev:
                ev[2].record()
                events.append(ev)
            launches[0] += 3
        if world > 1:
            dist.all_gather_into_tensor(ll_all, ll_d)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for _ in range(args.warmup):
        step()
    # (the clock sampler waits for nvidia-smi's first sample: start it BEFORE the barrier, or the
    # other ranks would spend that second waiting for rank 0 inside their timed all-gather)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    events = []
    launches[0] = 0
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    for _ in range(args.steps):
        step(events)
    ev_b.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = max_over_ranks(ev_a.elapsed_time(ev_b))
    value = world * B * args.steps / (elapsed_ms * 1e-3)
    build_ms = sum(ev[0].elapsed_time(ev[1]) for ev in events) / len(events)
    potrf_ms = sum(ev[1].elapsed_time(ev[2]) for ev in events) / len(events)
    n_fail = int((info_d != 0).sum().item())
    gpu_launches = launches[0]

    # ---- end to end through the public API, host buffers -------------------------------
    # N = 1: BIGgp.log_likelihood_batch(host array).  N > 1: the product's sharded entry point,
    # every rank handing in the same global batch of 4096 * N rows from host memory.
    H_global = H if world == 1 else np.concatenate([go.synthetic_biggp_hypers(B, seed=1 + r) for r in range(world)])
    A_host = torch.from_numpy(np.ascontiguousarray(H_global)).pin_memory()
    out_host = torch.empty(B * world, dtype=torch.float64).pin_memory()

    def e2e_step():
        if world > 1:
            ll = mfdist.log_likelihood_batch_sharded(gp, A_host)     # H2D of this rank's shard inside
        else:
            ll = gp.log_likelihood_batch(A_host)                      # H2D of the hyper-parameters inside
        out_host.copy_(ll, non_blocking=False)                         # D2H of the result

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_value = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)
    e2e_parity = None
    if world > 1:
        e2e_parity = float((out_host[rank * B:(rank + 1) * B].to(dev) - ll_d).abs().max().item())

    # ---- strong scaling: BASELINE configs[2], a FIXED global batch cut across the ranks ----
    ws = None                                  # the engine re-sizes its workspace for the larger shards
    strong = None
    if world > 1 and not args.no_strong:
        strong = []
        for total in (4096, 32768):
            A_dev = torch.as_tensor(go.synthetic_biggp_hypers(total, seed=7), device=dev)

            def sharded():
                return mfdist.log_likelihood_batch_sharded(gp, A_dev)

            for _ in range(2):
                got = sharded()
            barrier()
            e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                got = sharded()
            e1_.record()
            barrier()
            ms_n = max_over_ranks(e0.elapsed_time(e1_)) / args.steps
            # the same batch on ONE GPU of this box, same run: rank 0 alone, the others wait
            ms_1, parity = 0.0, 0.0
            if rank == 0:
                alone = gp.log_likelihood_batch(A_dev)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(args.steps):
                    alone = gp.log_likelihood_batch(A_dev)
                e1_.record()
                torch.cuda.synchronize()
                ms_1 = e0.elapsed_time(e1_) / args.steps
                ok = torch.isfinite(alone)
                parity = float(((got[ok] - alone[ok]).abs() / alone[ok].abs()).max().item())
                parity = parity if bool((torch.isfinite(got) == ok).all()) else float("inf")
            barrier()
            strong.append({"global_batch": total, "per_gpu_batch": total // world, "value": total / (ms_n * 1e-3),
                           "unit": UNIT, "ms_per_step": ms_n, "single_gpu_ms_per_step": ms_1,
                           "efficiency_vs_1gpu_same_run": ms_1 / (world * ms_n) if ms_n > 0 else None,
                           "sharded_parity_max_rel": parity,
                           "entry_point": "miniframe_b200.dist.log_likelihood_batch_sharded (device-resident batch)"})
            del A_dev

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's cores (N = 1 only), bounded sample ---------------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        rate, done, vals, cores, kind = cpu_port_evals(N, H, args.cpu_seconds, 512)
        got = ll_d[:done].cpu().numpy()
        parity = float(np.max(np.abs(got - vals) / np.abs(vals)))
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "first %d of the %d hyper-parameter sets, %s (numpy %s + scipy LAPACK), pool of "
                         "%d single-threaded workers"
                         % (done, B, "the unmodified reference's BIGgp.log_likelihood (oracle/_ref)"
                            if kind == "reference" else "oracle literal port of the reference",
                            np.__version__, cores)}

    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    per_launch = min(chunk, B)
    potrf_tflops = flops_per_eval(n) * per_launch / (potrf_ms * 1e-3) / 1e12
    build_gbs = build_bytes_per_eval(n) * per_launch / (build_ms * 1e-3) / 1e9
    fp64_peak = max(peaks.get("cublas_dgemm_4096_tflops", 0.0), peaks.get("dmma_probe_tflops", 0.0),
                    peaks.get("dmma_probe_sustained_tflops", 0.0))
    potrf_traffic, potrf_note = ncu_traffic("mf_potrf_ll_kernel", n, per_launch)
    build_traffic, build_note = ncu_traffic("mf_cov_build_big_il_kernel", n, per_launch)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(H_global.nbytes), "d2h_bytes_per_step": int(B * 8 * world * world),
                "entry_point": "BIGgp.log_likelihood_batch(host array)" if world == 1 else
                               "dist.log_likelihood_batch_sharded(gp, host array of %d rows)" % (B * world),
                "max_abs_diff_vs_device_timed_results": e2e_parity},
        "gpu_launches": gpu_launches,
        "roofline": {"kernel": "mf_potrf_ll_kernel", "bound": "tensor", "achieved": potrf_tflops,
                     "peak": fp64_peak, "unit": "TFLOP/s", "frac": potrf_tflops / fp64_peak,
                     "traffic": potrf_traffic, "traffic_unit": potrf_note,
                     "peak_source": "FP64 is not in MEASURED_PEAKS.json: max of cuBLAS DGEMM 4096^3 (best of 5) and "
                                    "the DMMA register probe (5 ms bursts, and 0.45 s launches at the "
                                    "factorisation's occupancy), all measured in this run",
                     "ms_per_launch": potrf_ms, "evals_per_launch": per_launch,
                     "flops_per_eval": flops_per_eval(n)},
        "roofline_build": {"kernel": "mf_cov_build_big_il_kernel", "bound": "hbm", "achieved": build_gbs,
                           "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                           "traffic": build_traffic, "traffic_unit": build_note,
                           "peak_source": hbm_src, "ms_per_launch": build_ms,
                           "bytes_per_eval": build_bytes_per_eval(n)},
        "fp64_peaks_measured": peaks,
        "cpu_baseline": cpu,
        "parity_vs_cpu_port_max_rel": parity,
        "non_pd_samples": n_fail,
        "kernel_share": {"build": build_ms / (build_ms + potrf_ms), "potrf": potrf_ms / (build_ms + potrf_ms)},
    }
    if strong is not None:
        line["strong"] = strong
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="hyper-parameter sets per GPU per step")
    ap.add_argument("--epochs", type=int, default=500, help="N; the matrix order is 3N")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-sample", type=int, default=2,
                    help="likelihoods per host core per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate tests/golden/golden_v1.npz + golden_v1.json by running the
UNMODIFIED reference (imported from /root/reference via oracle/ref_shim.py).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every case records its inputs and the reference's outputs (log-likelihood and,
for the small cases, the full covariance matrix).  The reference ships no
expected outputs of its own (miniframe/tests/test_kernels.py:1-4), so these
fixtures are what pins the oracle and the CUDA path to the reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.gp_oracle import synthetic_biggp_hypers, synthetic_jitter, synthetic_series  # noqa: E402


def spot_dataset(ref, stride=1):
    """Data preparation of examples/BIGgp_mcmc.py:17-33 on 1spot_soap.rdb."""
    phase, flux, rv, bis = np.loadtxt(os.path.join(ref.datasets, "1spot_soap.rdb"),
                                      skiprows=2, unpack=True, usecols=(0, 1, 2, 3))
    phase, flux, rv, bis = (v[::stride] for v in (phase, flux, rv, bis))
    t = 25.05 * phase
    rhk = flux ** 2
    rv = 100 * rv
    bis = 100 * bis
    rms = lambda v: np.sqrt((1. / v.size * np.sum(v ** 2)))
    return dict(t=t, rv=rv, rverr=0.05 * rms(rv) * np.ones(rv.size),
                rhk=rhk, sig_rhk=0.20 * rms(rhk) * np.ones(rhk.size),
                bis=bis, sig_bis=0.10 * rms(bis) * np.ones(bis.size))


def synth_dataset(N, seed):
    t, (rv, rhk, bis), (e1, e2, e3) = synthetic_series(N, seed)
    return dict(t=t, rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)


def main():
    ref = ref_shim.load()
    K_, M_ = ref.kernels, ref.means
    kern = {"QuasiPeriodic": K_.QuasiPeriodic, "SquaredExponential": K_.SquaredExponential}
    meancls = {None: None, "Constant": M_.Constant, "Linear": M_.Linear,
               "Parabola": M_.Parabola, "Cubic": M_.Cubic, "Sine": M_.Sine,
               "Keplerian": M_.Keplerian}

    datasets = {
        "spot200": spot_dataset(ref, 1),
        "spot40": spot_dataset(ref, 5),
        "synth64": synth_dataset(64, 0),
        "synth96": synth_dataset(96, 7),
    }
    arrays = {}
    for dname, d in datasets.items():
        for k, v in d.items():
            arrays["data/%s/%s" % (dname, k)] = np.asarray(v, dtype=float)

    a_plot = [1083.1091125670669, 1.215034890646895, 25.07312864134963, 0.031950873139068185,
              6.064550597545819, 0 * 4.23391412490362, 0 * 0.3552833092394814,
              0 * 12.807709071739335, 0 * 9.755026033334879]      # examples/BIGgp_plot.py:55-58
    a_full = [1083.1091125670669, 1.215034890646895, 25.07312864134963, 0.031950873139068185,
              6.064550597545819, 4.23391412490362, 0.3552833092394814,
              12.807709071739335, 9.755026033334879]
    a_se = [20.0, 0.03, 1.0, 0.5, 0.8, 0.9, 0.4]
    a_qp2 = [20.0, 1.2, 25.0, 0.03, 1.0, 0.5, 0.8, 0.9, 0.4]

    cases = []

    def big_case(name, dname, kname, a, means=(None, None, None), b=(), jitter=None,
                 store_K=False, matrix_kwargs=None):
        d = datasets[dname]
        cls = ref.BIGgp if jitter is None else ref.BIGgpPlusJitt
        gp = cls(kern[kname], [meancls[m] for m in means], t=ref_shim.as_time(d["t"]),
                 rv=d["rv"], rverr=d["rverr"], rhk=d["rhk"], sig_rhk=d["sig_rhk"],
                 bis=d["bis"], sig_bis=d["sig_bis"])
        if jitter is None:
            ll = gp.log_likelihood(list(a), list(b))
        else:
            ll = gp.log_likelihood(list(a), list(b), list(jitter))
        case = dict(name=name, framework="BIGgp" if jitter is None else "BIGgpPlusJitt",
                    data=dname, kernel=kname, a=list(map(float, a)), means=list(means),
                    b=list(map(float, b)), jitter=None if jitter is None else list(map(float, jitter)),
                    ll=float(ll), resid_key=None, K_key=None, matrix_kwargs=matrix_kwargs)
        resid = gp.y - gp.mean()
        arrays["resid/%s" % name] = resid
        case["resid_key"] = "resid/%s" % name
        if store_K:
            kw = matrix_kwargs or {}
            Kmat = gp.compute_matrix(list(a), **kw) if jitter is None else \
                gp.compute_matrix(list(a), list(jitter), **kw)
            arrays["K/%s" % name] = Kmat
            case["K_key"] = "K/%s" % name
        cases.append(case)

    def small_case(name, dname, kname, extraname, a, order, means=None, b=(), c=(),
                   store_K=False):
        d = datasets[dname]
        M = len(order)
        means = [None] * M if means is None else list(means)
        pairs = {"rv": ("rv", "rverr"), "bis": ("bis", "sig_bis"), "rhk": ("rhk", "sig_rhk")}
        args = []
        for o in order:
            args += [d[pairs[o][0]], d[pairs[o][1]]]
        gp = ref.SMALLgp(kern[kname], None if extraname is None else kern[extraname],
                         [meancls[m] for m in means], d["t"], *args)
        ll = gp.log_likelihood(list(a), list(b), list(c))
        case = dict(name=name, framework="SMALLgp", data=dname, kernel=kname, extrakernel=extraname,
                    a=list(map(float, a)), order=list(order), means=means, b=list(map(float, b)),
                    c=list(map(float, c)), ll=float(ll), K_key=None)
        arrays["resid/%s" % name] = np.concatenate(gp.y) - gp.mean()
        case["resid_key"] = "resid/%s" % name
        if store_K:
            arrays["K/%s" % name] = gp.compute_matrix(list(a), list(c))
            case["K_key"] = "K/%s" % name
        cases.append(case)

    # ---- BIGgp -----------------------------------------------------------
    big_case("big_plot_zeros", "spot200", "QuasiPeriodic", a_plot,
             means=("Constant", "Constant", "Constant"), b=(0, 0, 0))
    big_case("big_full", "spot200", "QuasiPeriodic", a_full)
    big_case("big_full_40", "spot40", "QuasiPeriodic", a_full, store_K=True)
    big_case("big_qp2_40", "spot40", "QuasiPeriodic", a_qp2, store_K=True)
    big_case("big_qp2_200", "spot200", "QuasiPeriodic", a_qp2)
    big_case("big_se_40", "spot40", "SquaredExponential", a_se, store_K=True)
    big_case("big_se_200", "spot200", "SquaredExponential", a_se)
    big_case("big_means", "spot200", "QuasiPeriodic", a_qp2,
             means=("Constant", "Linear", "Keplerian"),
             b=(0.3, 0.01, -0.2, 12.5, 0.8, 0.2, 0.4, 1.1, 0.05))
    big_case("big_means_poly", "spot40", "QuasiPeriodic", a_qp2,
             means=("Parabola", None, "Sine"), b=(1e-3, -0.02, 0.1, 0.4, 11.0, 0.3, -0.1))
    big_case("big_nugget_40", "spot40", "QuasiPeriodic", a_qp2, store_K=True,
             matrix_kwargs=dict(nugget=True))
    big_case("big_noyerr_40", "spot40", "QuasiPeriodic", a_qp2, store_K=True,
             matrix_kwargs=dict(yerr=False))
    # non positive definite: no white noise, long scale, no measurement error -> -inf path
    d = dict(datasets["spot40"])
    d["rverr"] = d["rverr"] * 0
    d["sig_rhk"] = d["sig_rhk"] * 0
    d["sig_bis"] = d["sig_bis"] * 0
    datasets["spot40_noerr"] = d
    for k, v in d.items():
        arrays["data/spot40_noerr/%s" % k] = np.asarray(v, dtype=float)
    big_case("big_nonpd", "spot40_noerr", "SquaredExponential",
             [500.0, 0.0, 1.0, 0.5, 0.8, 0.9, 0.4])
    # seeded prior-box draws on synthetic series
    H = synthetic_biggp_hypers(6, seed=0)
    for i in range(6):
        big_case("big_synth64_%d" % i, "synth64", "QuasiPeriodic", H[i], store_K=(i < 2))
    Hse = synthetic_biggp_hypers(3, seed=1, kname="SquaredExponential")
    for i in range(3):
        big_case("big_synth96_se_%d" % i, "synth96", "SquaredExponential", Hse[i])

    # ---- BIGgpPlusJitt ---------------------------------------------------
    big_case("jitt_full", "spot200", "QuasiPeriodic", a_full, jitter=(0.1, 0.2, 0.3))
    big_case("jitt_qp2_40", "spot40", "QuasiPeriodic", a_qp2, jitter=(0.1, 0.2, 0.3), store_K=True)
    J = synthetic_jitter(3, seed=0)
    for i in range(3):
        big_case("jitt_synth64_%d" % i, "synth64", "QuasiPeriodic", H[i], jitter=J[i])

    # ---- SMALLgp ---------------------------------------------------------
    a_small_ex = [1, 2, 3, 4, 2, 2, 2, 2, 2, 2, 1, 1, 1]       # examples/SMALLgp_likelihood.py:32-35
    a_small_plot = [1083.1091125670669, 1.215034890646895, 25.07312864134963,
                    0.031950873139068185, 6.064550597545819, 4.23391412490362, 0,
                    0.3552833092394814, 0, 0, 12.807709071739335, 9.755026033334879, 0]
    small_case("small_example", "spot200", "QuasiPeriodic", None, a_small_ex, ("rv", "bis", "rhk"))
    small_case("small_example_40", "spot40", "QuasiPeriodic", None, a_small_ex,
               ("rv", "bis", "rhk"), store_K=True)
    small_case("small_plot", "spot200", "QuasiPeriodic", None, a_small_plot, ("rv", "rhk", "bis"))
    small_case("small_plot_40", "spot40", "QuasiPeriodic", None, a_small_plot,
               ("rv", "rhk", "bis"), store_K=True)
    a_small_pd = [20.0, 1.2, 25.0, 0.05, 1.0, 0.5, 0.0, 0.8, 0.1, 0.0, 0.9, 0.4, 0.0]
    small_case("small_pd_40", "spot40", "QuasiPeriodic", None, a_small_pd,
               ("rv", "bis", "rhk"), store_K=True)
    small_case("small_pd_200", "spot200", "QuasiPeriodic", None, a_small_pd, ("rv", "bis", "rhk"))
    a_small_a3 = [20.0, 1.2, 25.0, 0.05, 1.0, 0.5, 0.01, 0.8, 0.1, 0.02, 0.9, 0.4, 0.005]
    small_case("small_a3_40", "spot40", "QuasiPeriodic", None, a_small_a3,
               ("rv", "bis", "rhk"), store_K=True)
    small_case("small_se_40", "spot40", "SquaredExponential", None,
               [15.0, 0.05, 1.0, 0.5, 0.01, 0.8, 0.1, 0.0, 0.9, 0.4, 0.003],
               ("rv", "bis", "rhk"), store_K=True)
    small_case("small_m2_40", "spot40", "QuasiPeriodic", None,
               [20.0, 1.2, 25.0, 0.05, 1.0, 0.5, 0.0, 0.8, 0.1, 0.0], ("rv", "rhk"), store_K=True)
    small_case("small_m1_40", "spot40", "QuasiPeriodic", None,
               [20.0, 1.2, 25.0, 0.05, 1.0, 0.5, 0.0], ("rv",), store_K=True)     # quirk Q5
    small_case("small_extra_se_40", "spot40", "QuasiPeriodic", "SquaredExponential", a_small_pd,
               ("rv", "bis", "rhk"), c=(3.0, 0.02, 0.5, 0.2, 0.1), store_K=True)
    small_case("small_extra_qp_40", "spot40", "SquaredExponential", "QuasiPeriodic",
               [15.0, 0.05, 1.0, 0.5, 0.0, 0.8, 0.1, 0.0, 0.9, 0.4, 0.0],
               ("rv", "bis", "rhk"), c=(30.0, 0.9, 11.0, 0.02, 0.5, 0.2, 0.1), store_K=True)
    small_case("small_means", "spot40", "QuasiPeriodic", None, a_small_pd, ("rv", "bis", "rhk"),
               means=("Constant", None, "Linear"), b=(0.25, 0.01, -0.3))

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **arrays)
    meta = dict(
        generator="tests/golden/make_golden.py",
        reference="jdavidrcamacho/mini-frame @ /root/reference (unmodified; shims in oracle/ref_shim.py)",
        numpy=np.__version__, scipy=__import__("scipy").__version__,
        cases=cases)
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for c in cases:
        print("%-22s ll=%r" % (c["name"], c["ll"]))
    print("wrote %d cases, %d arrays" % (len(cases), len(arrays)))


if __name__ == "__main__":
    main()

"""CPU: the oracle restatement against the committed golden vectors (produced
by the unmodified reference, tests/golden/make_golden.py) and, when the
reference tree is present, against the live reference."""
import numpy as np
import pytest

from conftest import assert_matrix_parity, rel_err
from oracle import gp_oracle as go
from oracle import ref_shim

PAIRS = {"rv": ("rv", "rverr"), "bis": ("bis", "sig_bis"), "rhk": ("rhk", "sig_rhk")}


def _big_matrix(case, d, fn):
    kw = dict(case.get("matrix_kwargs") or {})
    return fn(case["kernel"], case["a"], d["t"], d["rverr"], d["sig_rhk"], d["sig_bis"],
              jitter=case["jitter"], **kw)


def _small_inputs(case, d):
    yerr2 = np.concatenate([d[PAIRS[o][1]] ** 2 for o in case["order"]])
    return yerr2, len(case["order"])


def test_golden_has_reference_anchors(golden):
    # container-measured anchors of SURVEY.md section 8c / BASELINE.md section 2
    assert golden.cases["small_example"]["ll"] == pytest.approx(-1674.2585443588732, rel=1e-12)
    assert golden.cases["big_plot_zeros"]["ll"] == pytest.approx(-9611.906826167364, rel=1e-12)
    assert golden.cases["big_full"]["ll"] == pytest.approx(-36738.85373581005, rel=1e-12)
    assert golden.cases["jitt_full"]["ll"] == pytest.approx(-17256.7394826831, rel=1e-12)
    assert golden.cases["big_nonpd"]["ll"] == -np.inf


def test_biggp_literal_loglike_matches_golden(golden):
    for case in golden.by_framework("BIGgp", "BIGgpPlusJitt"):
        d = golden.data(case["data"])
        ll = go.biggp_loglike_literal(case["kernel"], case["a"], d["t"], golden.resid(case),
                                      d["rverr"], d["sig_rhk"], d["sig_bis"], jitter=case["jitter"])
        assert rel_err(ll, case["ll"]) <= 1e-13, case["name"]


def test_biggp_literal_matrix_is_bit_exact(golden):
    n = 0
    for case in golden.by_framework("BIGgp", "BIGgpPlusJitt"):
        Kref = golden.K(case)
        if Kref is None:
            continue
        K = _big_matrix(case, golden.data(case["data"]), go.biggp_matrix_literal)
        assert np.array_equal(K, Kref), case["name"]
        n += 1
    assert n >= 7


def test_biggp_closed_matrix_parity(golden):
    for case in golden.by_framework("BIGgp", "BIGgpPlusJitt"):
        Kref = golden.K(case)
        if Kref is None:
            continue
        K = _big_matrix(case, golden.data(case["data"]), go.biggp_matrix_closed)
        assert_matrix_parity(K, Kref)
        assert np.array_equal(K, K.T)


def test_smallgp_literal_matches_golden(golden):
    for case in golden.by_framework("SMALLgp"):
        d = golden.data(case["data"])
        yerr2, M = _small_inputs(case, d)
        ll = go.smallgp_loglike_literal(case["kernel"], case["extrakernel"], case["a"], case["c"],
                                        d["t"], golden.resid(case), yerr2, M)
        assert rel_err(ll, case["ll"]) <= 1e-13, case["name"]
        Kref = golden.K(case)
        if Kref is not None:
            K = go.smallgp_matrix_literal(case["kernel"], case["extrakernel"], case["a"], case["c"],
                                          d["t"], yerr2, M)
            assert np.array_equal(K, Kref), case["name"]


def test_smallgp_closed_matrix_parity(golden):
    for case in golden.by_framework("SMALLgp"):
        Kref = golden.K(case)
        if Kref is None:
            continue
        d = golden.data(case["data"])
        yerr2, M = _small_inputs(case, d)
        K = go.smallgp_matrix_closed(case["kernel"], case["extrakernel"], case["a"], case["c"],
                                     d["t"], yerr2, M)
        # the reference reads only the upper triangle (cho_factor lower=False)
        assert_matrix_parity(K, Kref, tri="upper")


def test_quirks(golden):
    # Q5: single-series SMALLgp collapses to the (nugget-scaled) error diagonal
    case = golden.cases["small_m1_40"]
    K = golden.K(case)
    assert np.count_nonzero(K - np.diag(np.diag(K))) == 0
    # Q3: BIGgp nugget / yerr switches change compute_matrix but never log_likelihood
    assert golden.cases["big_nugget_40"]["ll"] == golden.cases["big_qp2_40"]["ll"]
    assert not np.array_equal(golden.K(golden.cases["big_nugget_40"]), golden.K(golden.cases["big_qp2_40"]))
    # Q1: white noise sits on the diagonal of OFF-diagonal blocks too
    c = golden.cases["big_qp2_40"]
    K, n = golden.K(c), 40
    wn = c["a"][3]
    vc, vr, lc = c["a"][4], c["a"][5], c["a"][6]
    assert K[0, n] == pytest.approx(vc * lc * (1 + wn ** 2) + vr * lc * wn ** 2, rel=1e-14)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_oracle_against_live_reference():
    ref = ref_shim.load()
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(48, seed=3)
    H = go.synthetic_biggp_hypers(4, seed=3)
    gp = ref.BIGgp(ref.kernels.QuasiPeriodic, [None, None, None], t=ref_shim.as_time(t),
                   rv=rv, rverr=e1, rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    y = go.biggp_y(rv, rhk, bis)
    for a in H:
        assert np.array_equal(gp.compute_matrix(list(a)),
                              go.biggp_matrix_literal("QuasiPeriodic", a, t, e1, e2, e3))
        ll = go.biggp_loglike_literal("QuasiPeriodic", a, t, y, e1, e2, e3)
        assert rel_err(ll, gp.log_likelihood(list(a), [])) <= 1e-14

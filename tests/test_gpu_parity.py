"""GPU parity tests proper: the CUDA path, called through the drop-in classes and
the C ABI, against the golden vectors of the unmodified reference and against the
oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  covariance entries: max|dK|/max|K| <= 1e-12 and per entry
                      |dK_ij| <= 1e-12 * max(|K_ij|, 1e-4 max|K|)  (conftest.assert_matrix_parity
                      explains the floor; test_floor_evidence prints what lies between 1e-6 and 1e-4);
  log-likelihood:     |d|/|ll| <= 1e-9;  -inf must agree exactly.
"""
import ctypes

import numpy as np
import pytest

from conftest import assert_matrix_parity, rel_err
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
PAIRS = {"rv": ("rv", "rverr"), "bis": ("bis", "sig_bis"), "rhk": ("rhk", "sig_rhk")}


@pytest.fixture(scope="module")
def mf():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import miniframe_b200
    from miniframe_b200 import _lib
    _lib.load()
    return miniframe_b200


def _means(mf, names):
    table = {None: None, "Constant": mf.means.Constant, "Linear": mf.means.Linear,
             "Parabola": mf.means.Parabola, "Cubic": mf.means.Cubic, "Sine": mf.means.Sine,
             "Keplerian": mf.means.Keplerian}
    return [table[m] for m in names]


def _make_gp(mf, golden, case):
    d = golden.data(case["data"])
    kern = getattr(mf.kernels, case["kernel"])
    if case["framework"] == "SMALLgp":
        args = []
        for o in case["order"]:
            args += [d[PAIRS[o][0]], d[PAIRS[o][1]]]
        ek = None if case["extrakernel"] is None else getattr(mf.kernels, case["extrakernel"])
        return mf.SMALLgp.SMALLgp(kern, ek, _means(mf, case["means"]), d["t"], *args)
    cls = mf.BIGgp.BIGgp if case["framework"] == "BIGgp" else mf.BIGgpPlusJitt.BIGgp
    return cls(kern, _means(mf, case["means"]), t=d["t"], rv=d["rv"], rverr=d["rverr"],
               rhk=d["rhk"], sig_rhk=d["sig_rhk"], bis=d["bis"], sig_bis=d["sig_bis"])


def _loglike(gp, case):
    if case["framework"] == "SMALLgp":
        return gp.log_likelihood(case["a"], case["b"], case["c"])
    if case["framework"] == "BIGgpPlusJitt":
        return gp.log_likelihood(case["a"], case["b"], case["jitter"])
    return gp.log_likelihood(case["a"], case["b"])


def _matrix(gp, case):
    if case["framework"] == "SMALLgp":
        return gp.compute_matrix(case["a"], case["c"])
    kw = dict(case.get("matrix_kwargs") or {})
    if case["framework"] == "BIGgpPlusJitt":
        return gp.compute_matrix(case["a"], case["jitter"], **kw)
    return gp.compute_matrix(case["a"], **kw)


def test_golden_covariance_matrices(mf, golden):
    n = 0
    for case in golden.meta["cases"]:
        Kref = golden.K(case)
        if Kref is None:
            continue
        K = _matrix(_make_gp(mf, golden, case), case)
        assert K.shape == Kref.shape and K.dtype == np.float64
        assert_matrix_parity(K, Kref)
        n += 1
    assert n >= 15


def test_golden_log_likelihoods(mf, golden):
    worst = 0.0
    for case in golden.meta["cases"]:
        ll = _loglike(_make_gp(mf, golden, case), case)
        assert isinstance(ll, float)
        err = rel_err(ll, case["ll"])
        assert err <= LL_RTOL, (case["name"], ll, case["ll"])
        worst = max(worst, err)
    print("worst relative log-likelihood error over golden cases: %.3e" % worst)


def test_reference_quirks_survive(mf, golden):
    # Q3: BIGgp.log_likelihood ignores nugget; compute_matrix honours it
    case = golden.cases["big_qp2_40"]
    gp = _make_gp(mf, golden, case)
    assert gp.log_likelihood(case["a"], [], nugget=True) == gp.log_likelihood(case["a"], [], nugget=False)
    assert gp.minus_log_likelihood(case["a"], []) == -gp.log_likelihood(case["a"], [])
    # non positive definite -> -inf, not an exception
    nonpd = golden.cases["big_nonpd"]
    assert _loglike(_make_gp(mf, golden, nonpd), nonpd) == -np.inf
    # NaN hyper-parameters raise like scipy's check_finite does in the reference
    with pytest.raises(ValueError):
        gp.log_likelihood([np.nan] + case["a"][1:], [])
    # mean parameters: wrong length asserts, right length mutates state
    gm = _make_gp(mf, golden, golden.cases["big_means"])
    with pytest.raises(AssertionError):
        gm.log_likelihood(golden.cases["big_means"]["a"], [0.0])
    gm.log_likelihood(golden.cases["big_means"]["a"], golden.cases["big_means"]["b"])
    assert gm.mean_pars == golden.cases["big_means"]["b"]
    # PlusJitt.minus_log_likelihood is unusable without c, as in the reference
    gj = _make_gp(mf, golden, golden.cases["jitt_qp2_40"])
    with pytest.raises(TypeError):
        gj.minus_log_likelihood(golden.cases["jitt_qp2_40"]["a"], [])


def test_block_methods(mf, golden):
    case = golden.cases["big_qp2_40"]
    d = golden.data(case["data"])
    gp = _make_gp(mf, golden, case)
    k11, k22, k33, k12, k13, k23 = go.biggp_blocks_literal(case["kernel"], case["a"], d["t"])
    for got, want in ((gp.k11(case["a"], d["t"]), k11), (gp.k22(case["a"], d["t"]), k22),
                      (gp.k33(case["a"], d["t"]), k33), (gp.k12(case["a"], d["t"]), k12),
                      (gp.k13(case["a"], d["t"]), k13), (gp.k23(case["a"], d["t"]), k23)):
        assert_matrix_parity(got, want)


@pytest.mark.parametrize("N", [1, 2, 5, 21, 22, 33, 43, 64, 85, 128, 150])
def test_edge_sizes_against_oracle(mf, N):
    """n = 3N walks over every tile boundary of the kernels: below one 64 block,
    exact multiples of 64 and 128 (the residual row then opens a new tile), odd N
    (scalar store path), ragged last blocks."""
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=N)
    H = go.synthetic_biggp_hypers(5, seed=N)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    y = go.biggp_y(rv, rhk, bis)
    got = gp.log_likelihood_batch(H)
    assert got.shape == (5,) and got.dtype == np.float64
    for i in range(5):
        want = go.biggp_loglike_literal("QuasiPeriodic", H[i], t, y, e1, e2, e3)
        assert rel_err(got[i], want) <= LL_RTOL, (N, i, got[i], want)
    K = gp.compute_matrix(H[0])
    assert_matrix_parity(K, go.biggp_matrix_literal("QuasiPeriodic", H[0], t, e1, e2, e3))
    assert np.array_equal(K, K.T)


def test_factor_and_z_against_lapack(mf):
    """The factor itself: L L^T = K, L matches numpy's Cholesky, z = L^-1 r."""
    import torch
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    N = 70
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=11)
    H = go.synthetic_biggp_hypers(3, seed=11)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    n = 3 * N
    prob = gp._problem(N)
    Kd = eng.build_cov(prob, eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(H), None, full=False)
    resid = eng.tensor(go.biggp_y(rv, rhk, bis)[None, :])
    ll, info = eng.potrf_loglike(Kd, n, resid)
    assert info.cpu().numpy().tolist() == [0, 0, 0]
    out = Kd.cpu().numpy()
    for b in range(3):
        Kref = go.biggp_matrix_literal("QuasiPeriodic", H[b], t, e1, e2, e3)
        Lref = np.linalg.cholesky(Kref)
        L = np.tril(out[b, :n, :n])
        assert np.abs(L - Lref).max() <= 1e-10 * np.abs(Lref).max()
        assert np.abs(L @ L.T - Kref).max() <= 1e-13 * np.abs(Kref).max()
        zref = np.linalg.solve(Lref, go.biggp_y(rv, rhk, bis))
        assert np.abs(out[b, n, :n] - zref).max() <= 1e-9 * np.abs(zref).max()
    # stand-alone solve + logdet on that factor, a different residual per sample
    rng = np.random.default_rng(3)
    R = rng.standard_normal((3, n))
    z, ll2 = eng.solve_logdet(Kd, n, eng.tensor(R), info)
    for b in range(3):
        Kref = go.biggp_matrix_literal("QuasiPeriodic", H[b], t, e1, e2, e3)
        assert rel_err(float(ll2[b]), go.loglike_from_matrix(Kref, R[b])) <= LL_RTOL


@pytest.mark.parametrize("N,G", [(30, 2), (70, 2), (70, 5), (200, 3), (500, 4), (500, 12)])
def test_cooperative_groups_match_one_cta_per_matrix(mf, N, G):
    """Several CTAs per matrix (the path taken when the batch is smaller than the GPU) give the
    same factor, bit for bit, as one CTA per matrix; failures stay per sample."""
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=5)
    gp = mf.BIGgp.BIGgp(mf.kernels.SquaredExponential, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    n = 3 * N
    good1 = [3.0, 0.5, 1.0, 0.5, 0.8, 0.9, 0.4]
    good2 = [20.0, 0.03, 1.0, 0.5, 0.8, 0.9, 0.4]
    bad = [5000.0, 0.0, 1e8, 1e8, 1e8, 1e8, 1e8]
    A = np.array([good1, bad, good2, good1, bad])
    prob = gp._problem(N)
    resid = eng.tensor(go.biggp_y(rv, rhk, bis)[None, :])
    out = {}
    try:
        for g in (1, G):
            eng.set_option("potrf_group", g)
            Kd = eng.build_cov(prob, eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(A), None, full=False)
            ll, info = eng.potrf_loglike(Kd, n, resid)
            out[g] = (Kd.cpu().numpy(), ll.cpu().numpy(), info.cpu().numpy())
    finally:
        eng.set_option("potrf_group", 0)
    K1, ll1, info1 = out[1]
    KG, llG, infoG = out[G]
    assert np.array_equal(info1, infoG)
    assert (info1 != 0).tolist() == [False, True, False, False, True]
    y = go.biggp_y(rv, rhk, bis)
    for b in (0, 2, 3):
        L1 = np.tril(K1[b, :n, :n])
        LG = np.tril(KG[b, :n, :n])
        assert np.array_equal(L1, LG)
        assert np.array_equal(K1[b, n, :n], KG[b, n, :n])
        assert float(llG[b]) == float(ll1[b])          # same read-back order whoever wrote the rows
        want = go.biggp_loglike_literal("SquaredExponential", A[b], t, y, e1, e2, e3)
        assert rel_err(float(llG[b]), want) <= LL_RTOL
    assert np.all(np.isinf(llG[[1, 4]])) and np.all(llG[[1, 4]] < 0)


@pytest.mark.parametrize("N", [1, 7, 21, 43, 64, 100, 150, 213])
def test_small_tile_variant_against_oracle_and_large_tile_variant(mf, N):
    """Matrices of order <= 1100 in batches larger than the GPU run on 64-row tiles, three CTAs per
    SM: same factor as the 128-row variant (the per-element operation order does not depend on the
    tile height), log-likelihoods within 1e-9 of the oracle, failures per sample."""
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=N)
    gp = mf.BIGgp.BIGgp(mf.kernels.SquaredExponential, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    n = 3 * N
    B = 460                                         # more than 3 x 148 CTAs: the persistent loop runs
    rng = np.random.default_rng(N)
    A = np.column_stack([rng.uniform(2, 30, B), rng.uniform(0.01, 0.5, B)] + [rng.uniform(0.2, 1.5, B)] * 5)
    A[5] = A[-2] = [5000.0, 0.0, 1e8, 1e8, 1e8, 1e8, 1e8]          # not positive definite in FP64
    prob = gp._problem(N)
    resid = eng.tensor(go.biggp_y(rv, rhk, bis)[None, :])
    out = {}
    try:
        eng.set_option("potrf_group", 1)
        for nmi in ("2", "1"):
            eng.set_option("potrf_tile", int(nmi))
            Kd = eng.build_cov(prob, eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(A), None, full=False)
            ll, info = eng.potrf_loglike(Kd, n, resid)
            out[nmi] = (Kd.cpu().numpy(), ll.cpu().numpy(), info.cpu().numpy())
    finally:
        eng.set_option("potrf_group", 0)
        eng.set_option("potrf_tile", 0)
    (K2, ll2, i2), (K1, ll1, i1) = out["2"], out["1"]
    assert np.array_equal(i1, i2)
    bad = i1 != 0
    if N >= 7:
        assert bad[5] and bad[-2]
    ok = np.nonzero(~bad)[0]
    assert ok.size >= B - 8
    for b in ok[:: max(1, ok.size // 40)]:
        assert np.array_equal(np.tril(K1[b, :n, :n]), np.tril(K2[b, :n, :n]))
        assert np.array_equal(K1[b, n, :n], K2[b, n, :n])
    assert np.array_equal(ll1[ok], ll2[ok]) and np.all(np.isinf(ll1[bad]))
    y = go.biggp_y(rv, rhk, bis)
    for b in ok[[0, 1, ok.size // 2, -1]]:
        want = go.biggp_loglike_literal("SquaredExponential", A[b], t, y, e1, e2, e3)
        assert rel_err(float(ll1[b]), want) <= LL_RTOL


def test_batch_mixes_failures_and_successes(mf, golden):
    """Non-positive-definite samples give -inf for themselves only."""
    d = golden.data("spot40")
    gp = mf.BIGgp.BIGgp(mf.kernels.SquaredExponential, [None, None, None], t=d["t"], rv=d["rv"],
                        rverr=d["rverr"], rhk=d["rhk"], sig_rhk=d["sig_rhk"], bis=d["bis"],
                        sig_bis=d["sig_bis"])
    # amplitudes 1e8 with a 5000-day scale: rounding noise of the pivots swamps yerr^2
    bad1 = [5000.0, 0.0, 1e8, 1e8, 1e8, 1e8, 1e8]
    bad2 = [800.0, 1e-6, 1e8, 5e7, 1e8, 1e8, 1e8]
    good1 = [3.0, 0.5, 1.0, 0.5, 0.8, 0.9, 0.4]
    good2 = [20.0, 0.03, 1.0, 0.5, 0.8, 0.9, 0.4]
    A = np.array([good1, bad1, good2, bad2, bad1, good1])
    y = go.biggp_y(d["rv"], d["rhk"], d["bis"])
    got = gp.log_likelihood_batch(A)
    want = np.array([go.biggp_loglike_literal("SquaredExponential", a, d["t"], y, d["rverr"],
                                              d["sig_rhk"], d["sig_bis"]) for a in A])
    assert np.isinf(want).sum() == 3
    assert np.array_equal(np.isinf(got), np.isinf(want))
    ok = ~np.isinf(want)
    assert np.all(np.abs(got[ok] - want[ok]) <= LL_RTOL * np.abs(want[ok]))
    assert gp.log_likelihood(bad1, []) == -np.inf


def test_plusjitt_and_smallgp_batches(mf, golden):
    d = golden.data("synth64")
    y = go.biggp_y(d["rv"], d["rhk"], d["bis"])
    H, J = go.synthetic_biggp_hypers(7, seed=4), go.synthetic_jitter(7, seed=4)
    gj = mf.BIGgpPlusJitt.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t=d["t"], rv=d["rv"],
                                rverr=d["rverr"], rhk=d["rhk"], sig_rhk=d["sig_rhk"], bis=d["bis"],
                                sig_bis=d["sig_bis"])
    got = gj.log_likelihood_batch(H, None, J)
    for i in range(7):
        want = go.biggp_loglike_literal("QuasiPeriodic", H[i], d["t"], y, d["rverr"], d["sig_rhk"],
                                        d["sig_bis"], jitter=J[i])
        assert rel_err(got[i], want) <= LL_RTOL
    # SMALLgp, M = 3 with an extra kernel, per-sample mean parameters
    gs = mf.SMALLgp.SMALLgp(mf.kernels.QuasiPeriodic, mf.kernels.SquaredExponential,
                            [mf.means.Constant, None, mf.means.Linear], d["t"],
                            d["rv"], d["rverr"], d["bis"], d["sig_bis"], d["rhk"], d["sig_rhk"])
    rng = np.random.default_rng(9)
    B = 6
    A = np.column_stack([rng.uniform(10, 40, B), rng.uniform(0.8, 1.6, B), rng.uniform(15, 35, B),
                         rng.uniform(0.01, 0.1, B)] +
                        [rng.uniform(0.2, 1.5, B), rng.uniform(0.0, 0.3, B), rng.uniform(0.0, 0.01, B)] * 3)
    C = np.column_stack([rng.uniform(2, 6, B), rng.uniform(0.01, 0.05, B)] + [rng.uniform(0.1, 0.6, B)] * 3)
    Bm = rng.normal(0, 0.2, (B, 3))
    got = gs.log_likelihood_batch(A, Bm, C)
    yerr2 = np.concatenate([d["rverr"], d["sig_bis"], d["sig_rhk"]]) ** 2
    ycat = np.concatenate([d["rv"], d["bis"], d["rhk"]])
    for i in range(B):
        gs.mean_pars = Bm[i]
        want = go.smallgp_loglike_literal("QuasiPeriodic", "SquaredExponential", A[i], C[i], d["t"],
                                          ycat - gs.mean(), yerr2, 3)
        assert rel_err(got[i], want) <= LL_RTOL, (i, got[i], want)


def test_full_size_n1500_against_oracle(mf):
    """BASELINE config 3 shape (N = 500, n = 1500), a handful of samples the CPU
    oracle can finish in seconds: covariance entries and log-likelihoods."""
    N = 500
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=0)
    H = go.synthetic_biggp_hypers(4, seed=0)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    y = go.biggp_y(rv, rhk, bis)
    got = gp.log_likelihood_batch(H)
    conds = []
    for i in range(4):
        Kref = go.biggp_matrix_literal("QuasiPeriodic", H[i], t, e1, e2, e3)
        want = go.loglike_from_matrix(Kref, y)
        assert rel_err(got[i], want) <= LL_RTOL, (i, got[i], want)
        if i == 0:
            assert_matrix_parity(gp.compute_matrix(H[0]), Kref)
            conds.append(np.linalg.cond(Kref))
    print("n=1500 parity ok; cond(K[0]) = %.3e" % conds[0])


def test_full_size_properties(mf):
    """Size-independent checks at the benchmark shape with more samples than
    resident CTAs (2 x 148): bit-reproducibility, independence from the batch slot,
    and the exact scaling law  ll(s K, s r) = ll(K, r) - n log s."""
    import torch
    N, B = 500, 320
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=1)
    H = go.synthetic_biggp_hypers(8, seed=1)
    A = np.tile(H, (B // 8, 1))
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    one = gp.log_likelihood_batch(A)
    two = gp.log_likelihood_batch(A)
    assert np.array_equal(one, two)
    assert np.array_equal(one.reshape(-1, 8), np.tile(one[:8], (B // 8, 1)))
    assert np.all(np.isfinite(one))
    s = 4.0                                  # power of two: every product scales exactly
    gp2 = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, s * rv, s * e1, s * rhk,
                         s * e2, s * bis, s * e3)
    Hs = H.copy()
    Hs[:, 4:] *= s
    scaled = gp2.log_likelihood_batch(Hs)
    n = 3 * N
    assert np.all(np.abs(scaled - (one[:8] - n * np.log(s))) <= 1e-11 * np.abs(one[:8]))
    # torch in -> device tensor out, same numbers
    dev = gp.log_likelihood_batch(torch.as_tensor(H, device="cuda"))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), one[:8])


def test_chunked_workspace_gives_identical_results(mf):
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    N = 40
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=2)
    H = go.synthetic_biggp_hypers(23, seed=2)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    whole = gp.log_likelihood_batch(H)
    eng.max_workspace_bytes = 5 * eng.matrix_stride(3 * N) * 8      # 5 matrices per chunk
    eng.drop_workspaces()
    try:
        chunked = gp.log_likelihood_batch(H)
    finally:
        eng.max_workspace_bytes = None
        eng.drop_workspaces()
    assert np.array_equal(whole, chunked)


def test_abi_argument_errors(mf):
    from miniframe_b200 import _lib
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    prob = _lib.Problem(framework=7, kernel=1, extra_kernel=-1, n_series=3, n_epochs=4,
                        hyper_len=9, extra_len=0, reserved=0, nugget=0.0)
    import torch
    buf = torch.zeros(4096, dtype=torch.float64, device="cuda")
    st = eng.lib.mf_build_cov(eng.handle, ctypes.byref(prob), 1, eng.ptr(buf), eng.ptr(buf), eng.ptr(buf),
                              None, 1, eng.ptr(buf), 16, 13 * 16, eng.stream())
    assert st == 1 and b"framework" in eng.lib.mf_last_error(eng.handle)
    st = eng.lib.mf_potrf_loglike(eng.handle, 1, 12, eng.ptr(buf), 12, 13 * 12, None, 0, eng.ptr(buf),
                                  eng.ptr(buf), eng.stream())
    assert st == 1
    # options: unknown names and values out of range are refused, the shipped library has no experiments
    h = eng.handle
    assert eng.lib.mf_set_option(h, b"no_such_option", 1) == 1
    assert eng.lib.mf_set_option(h, b"potrf_lookahead", 3) == 1 and eng.lib.mf_set_option(h, b"potrf_tile", 5) == 1
    assert eng.lib.mf_set_option(h, b"experiment", 8) != 0
    assert eng.lib.mf_set_option(h, b"potrf_lookahead", 0) == 0 and eng.lib.mf_set_option(h, b"l2_persist_mb", 0) == 0


def test_batched_keplerian_and_sine_means(mf, golden):
    """Per-sample mean parameters evaluated on the device for the whole batch (Keplerian: the
    reference's 100 fixed-point steps), against the oracle fed with the reference-style residual."""
    d = golden.data("synth64")
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [mf.means.Keplerian, mf.means.Sine, mf.means.Cubic],
                        t=d["t"], rv=d["rv"], rverr=d["rverr"], rhk=d["rhk"], sig_rhk=d["sig_rhk"],
                        bis=d["bis"], sig_bis=d["sig_bis"])
    rng = np.random.default_rng(12)
    B = 9
    H = go.synthetic_biggp_hypers(B, seed=6)
    Bm = np.column_stack([rng.uniform(5, 30, B), rng.uniform(0.1, 2, B), rng.uniform(0, 0.5, B),
                          rng.uniform(0, 6, B), rng.uniform(0, 6, B), rng.normal(0, 0.1, B),
                          rng.uniform(0.1, 1, B), rng.uniform(5, 30, B), rng.uniform(0, 6, B), rng.normal(0, 0.1, B),
                          rng.normal(0, 1e-6, B), rng.normal(0, 1e-4, B), rng.normal(0, 1e-2, B), rng.normal(0, 0.1, B)])
    assert gp.mean_pars_size == Bm.shape[1] == 14
    got = gp.log_likelihood_batch(H, Bm)
    for i in range(B):
        gp.mean_pars = Bm[i]
        resid = gp.y - gp.mean()                      # host numpy, as the reference does (BIGgp.py:303)
        K = go.biggp_matrix_literal("QuasiPeriodic", H[i], d["t"], d["rverr"], d["sig_rhk"], d["sig_bis"])
        assert rel_err(got[i], go.loglike_from_matrix(K, resid)) <= LL_RTOL
    assert rel_err(gp.log_likelihood(H[3], Bm[3]), float(got[3])) <= 1e-12


def test_predict_gp_against_reference_golden(mf):
    """predict_gp of BIGgp, BIGgpPlusJitt and SMALLgp against outputs of the unmodified
    reference (tests/golden/make_golden_predict.py): non-square, square (wn quirk) and on-data grids."""
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(here, "golden_predict_v1.npz"))
    meta = json.load(open(os.path.join(here, "golden_predict_v1.json")))
    means = [mf.means.Constant, mf.means.Linear, mf.means.Constant]
    for case in meta["cases"]:
        jit = case.get("jitter")
        if case.get("framework") == "base":
            # predict_G, predict_Gdot, predict_rv, predict_rhk, predict_bis (BIGgp.py:381-549)
            gp = mf.BIGgp.BIGgp(getattr(mf.kernels, case["kernel"]), [None, None, None], d["t"], d["rv"],
                                d["rverr"], d["rhk"], d["sig_rhk"], d["bis"], d["sig_bis"])
            grid, a, what = d[case["grid"]], case["a"], case["what"]
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                out = {"G": lambda: gp.predict_G(grid, d["rv"], a), "Gdot": lambda: gp.predict_Gdot(grid, d["bis"], a),
                       "rv": lambda: gp.predict_rv(grid, a), "rhk": lambda: gp.predict_rhk(grid, a),
                       "bis": lambda: gp.predict_bis(grid, a)}[what]()
            assert len(out) == case["nout"]
            for k, got in enumerate(out):
                want = d["%s_%d" % (case["name"], k)]
                assert got.shape == want.shape
                okk = np.isfinite(want)
                sc = max(np.abs(want[okk]).max(), 1e-300) if okk.any() else 1.0
                if got.ndim == 1 and k == 2:          # std: sqrt of a cancelling difference
                    big = okk & (want > 1e-4 * sc)
                    assert np.allclose(got[big], want[big], rtol=1e-6, atol=0), case["name"]
                else:
                    assert np.abs(got[okk] - want[okk]).max() <= 1e-9 * sc, (case["name"], k)
            continue
        if case.get("framework") == "SMALLgp":
            ek = None if case["extrakernel"] is None else getattr(mf.kernels, case["extrakernel"])
            gp = mf.SMALLgp.SMALLgp(getattr(mf.kernels, case["kernel"]), ek,
                                    [mf.means.Constant, None, mf.means.Linear], d["t"], d["rv"], d["rverr"],
                                    d["bis"], d["sig_bis"], d["rhk"], d["sig_rhk"])
            mean, std, cov = gp.predict_gp(d[case["grid"]], case["a"], case["b"], case["c"], model=case["model"])
        else:
            cls = mf.BIGgpPlusJitt.BIGgp if jit else mf.BIGgp.BIGgp
            gp = cls(getattr(mf.kernels, case["kernel"]), list(means), d["t"], d["rv"], d["rverr"],
                     d["rhk"], d["sig_rhk"], d["bis"], d["sig_bis"])
            args = (d[case["grid"]], case["a"], case["b"]) + ((jit,) if jit else ())
            mean, std, cov = gp.predict_gp(*args, model=case["model"])
        name = case["name"]
        wm, ws, wc = d[name + "_mean"], d[name + "_std"], d[name + "_cov"]
        # measured: mean 2e-13, covariance 7e-13 of max|cov|, std 1e-9 relative (K** - K* K^-1 K*^T
        # cancels; K carries only wn^2 on its diagonal, no observational errors)
        scale = max(np.abs(wc).max(), 1e-300)
        assert mean.shape == wm.shape and cov.shape == wc.shape
        assert np.abs(mean - wm).max() <= 1e-10 * max(1.0, np.abs(wm).max()), name
        assert np.abs(cov - wc).max() <= 1e-10 * scale, name
        ok = np.isfinite(ws) & (ws > 1e-4 * np.sqrt(scale))
        assert np.allclose(std[ok], ws[ok], rtol=1e-7, atol=0), name


def test_samples_have_the_covariance_they_are_drawn_from(mf):
    """sample / sample_from_G / sample_from_Gdot (reference BIGgp.py:320-378): the reference draws
    with scipy + numpy's global generator, so only the distribution can be compared."""
    N = 14
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=2)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    a = [20.0, 1.1, 23.0, 0.3, 1.3, 0.4, 0.9, 0.7, 0.5]
    draws = 20000
    for what, K in (("sample", gp.compute_matrix(a)), ("sample_from_G", None), ("sample_from_Gdot", None)):
        X = getattr(gp, what)(a, size=draws, seed=5)
        if what == "sample_from_G":
            K = mf.kernels.QuasiPeriodic(*a[:4])(t[:, None] - t[None, :])
        if what == "sample_from_Gdot":
            dd = mf.kernels.QuasiPeriodic.__subclasses__()[2]
            K = dd(*a[:4])(t[:, None] - t[None, :])
        assert X.shape == (draws, K.shape[0])
        C = X.T @ X / draws
        scale = np.abs(K).max()
        assert np.abs(C - K).max() <= 6.0 * scale * np.sqrt(2.0 / draws)
        assert np.abs(X.mean(axis=0)).max() <= 5.0 * np.sqrt(scale / draws)
        one = getattr(gp, what)(a, seed=5)
        assert one.shape == (K.shape[0],) and np.array_equal(one, getattr(gp, what)(a, seed=5))
    gj = mf.BIGgpPlusJitt.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    with pytest.raises(TypeError):
        gj.sample(a)
    Xj = gj.sample(a, [0.5, 0.2, 0.3], size=draws, seed=1)
    Kj = gj.compute_matrix(a, [0.5, 0.2, 0.3])
    assert np.abs(Xj.T @ Xj / draws - Kj).max() <= 6.0 * np.abs(Kj).max() * np.sqrt(2.0 / draws)


@pytest.mark.parametrize("tile", [2, 1])
@pytest.mark.parametrize("N", [43, 64, 81, 100, 150, 213, 400])
def test_lookahead_kernel_matches_plain_kernel(mf, N, tile):
    """Batches that fill the GPU take the look-ahead instantiation of the factorisation (the
    diagonal block of every block column factorised by one warp behind the next tile; on 128-row
    tiles a fragment is parked in place meanwhile): same factor, z, log-likelihood and info, bit
    for bit, as the plain one-CTA-per-matrix kernel - failures included - on both tile heights, at
    sizes that walk the tile boundaries (no tile 1, a ragged tile 1, ragged later tiles), and
    within 1e-9 of the oracle."""
    from miniframe_b200._engine import Engine
    eng = Engine.get()
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=100 + N)
    gp = mf.BIGgp.BIGgp(mf.kernels.SquaredExponential, [None, None, None], t, rv, e1, rhk, e2, bis, e3)
    n = 3 * N
    B = 330 if tile == 2 else 470                   # more than 2 (3) x 148 CTAs: the persistent loop hands out
    rng = np.random.default_rng(N)
    A = np.column_stack([rng.uniform(2, 30, B), rng.uniform(0.01, 0.5, B)] + [rng.uniform(0.2, 1.5, B)] * 5)
    A[3] = A[200] = A[-1] = [5000.0, 0.0, 1e8, 1e8, 1e8, 1e8, 1e8]          # not positive definite in FP64
    prob = gp._problem(N)
    resid = eng.tensor(go.biggp_y(rv, rhk, bis)[None, :])
    out = {}
    try:
        eng.set_option("potrf_group", 1)
        eng.set_option("potrf_tile", tile)
        for la in (1, 2):
            eng.set_option("potrf_lookahead", la)
            Kd = eng.build_cov(prob, eng.tensor(t), eng.tensor(gp._diag_yerr()), eng.tensor(A), None, full=False)
            ll, info = eng.potrf_loglike(Kd, n, resid)
            out[la] = (Kd.cpu().numpy(), ll.cpu().numpy(), info.cpu().numpy())
    finally:
        eng.set_option("potrf_group", 0)
        eng.set_option("potrf_tile", 0)
        eng.set_option("potrf_lookahead", 0)
    (K1, ll1, i1), (K2, ll2, i2) = out[1], out[2]
    assert np.array_equal(i1, i2)
    bad = i1 != 0
    assert bad[3] and bad[200] and bad[-1]
    ok = np.nonzero(~bad)[0]
    assert ok.size >= B - 8
    for b in ok[:: max(1, ok.size // 40)]:
        assert np.array_equal(np.tril(K1[b, :n, :n]), np.tril(K2[b, :n, :n]))
        assert np.array_equal(K1[b, n, :n], K2[b, n, :n])
    assert np.array_equal(ll1[ok], ll2[ok]) and np.all(np.isinf(ll2[bad])) and np.all(ll2[bad] < 0)
    y = go.biggp_y(rv, rhk, bis)
    for b in ok[[0, ok.size // 2, -1]]:
        want = go.biggp_loglike_literal("SquaredExponential", A[b], t, y, e1, e2, e3)
        assert rel_err(float(ll2[b]), want) <= LL_RTOL


def test_many_waves_hand_out(mf):
    """More than four waves of small matrices: the persistent kernels hand the matrices out through
    the atomic counter with no per-CTA limit (with fewer waves a CTA takes at most its share); every
    sample is evaluated exactly once, reproducibly, and agrees with the oracle."""
    N, B = 21, 3000
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=77)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1, rhk=rhk,
                        sig_rhk=e2, bis=bis, sig_bis=e3)
    H = go.synthetic_biggp_hypers(B, seed=78)
    a = np.asarray(gp.log_likelihood_batch(H))
    b = np.asarray(gp.log_likelihood_batch(H))
    assert a.shape == (B,) and np.all(np.isfinite(a)) and np.array_equal(a, b)
    y = go.biggp_y(rv, rhk, bis)
    for i in (0, 1, 443, 444, 445, B // 2, B - 2, B - 1):
        want = go.biggp_loglike_literal("QuasiPeriodic", H[i], t, y, e1, e2, e3)
        assert rel_err(float(a[i]), want) <= LL_RTOL
    # the same samples in another order give the same values: no slot dependence
    perm = np.random.default_rng(5).permutation(B)
    c = np.asarray(gp.log_likelihood_batch(H[perm]))
    assert np.array_equal(c, a[perm])

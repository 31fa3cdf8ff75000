"""Host logic of the ensemble sampler (stretch move, box prior) on CPU tensors, and -- on the
GPU -- that the chain's log-probabilities are the reference's log-likelihoods."""
import math

import numpy as np
import pytest
import torch

from miniframe_b200.sampler import EnsembleSampler, box_log_prob


def test_stretch_move_samples_a_gaussian():
    mu = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    sig = torch.tensor([0.5, 2.0, 1.0], dtype=torch.float64)

    def logp(P):
        return -0.5 * (((P - mu) / sig) ** 2).sum(dim=1)

    s = EnsembleSampler(32, 3, logp, seed=7, device="cpu")
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal((32, 3))
    p, lnp, _ = s.run_mcmc(p0, 300)
    s.reset()
    s.seed = 8
    p, lnp, _ = s.run_mcmc(p, 1500)
    flat = s.flatchain
    assert flat.shape == (32 * 1500, 3)
    assert torch.allclose(flat.mean(dim=0), mu, atol=0.12)
    assert torch.allclose(flat.std(dim=0), sig, rtol=0.08)
    acc = s.acceptance_fraction
    assert 0.3 < float(acc.mean()) < 0.8
    assert torch.equal(s.lnprobability[:, -1], lnp) and torch.equal(s.chain[:, -1], p)
    assert torch.allclose(lnp, logp(p))


def test_same_seed_same_chain_and_walker_count_rules():
    def logp(P):
        return -0.5 * (P ** 2).sum(dim=1)

    p0 = np.random.default_rng(1).standard_normal((8, 2))
    a = EnsembleSampler(8, 2, logp, seed=3, device="cpu")
    b = EnsembleSampler(8, 2, logp, seed=3, device="cpu")
    assert torch.equal(a.run_mcmc(p0, 20)[0], b.run_mcmc(p0, 20)[0])
    with pytest.raises(ValueError):
        EnsembleSampler(7, 2, logp)
    with pytest.raises(ValueError):
        EnsembleSampler(2, 2, logp)


def test_box_prior_rejects_without_evaluating():
    calls = []

    class FakeGP(object):
        def log_likelihood_batch(self, A):
            calls.append(A.clone())
            return -A.sum(dim=1)

    lp = box_log_prob(FakeGP(), [-1.0, -1.0], [1.0, math.log(10.0)])
    P = torch.tensor([[0.0, 0.0], [2.0, 0.0], [0.5, 2.2], [0.0, 2.4], [-1.0, -1.0]], dtype=torch.float64)
    out = lp(P)
    assert torch.isinf(out[[1, 3]]).all() and (out[[1, 3]] < 0).all()
    assert calls[0].shape == (3, 2)                       # only the rows inside the box were evaluated
    assert torch.allclose(out[[0, 2, 4]], -torch.exp(P[[0, 2, 4]]).sum(dim=1))
    # a whole half-ensemble outside the box: no likelihood call at all
    n = len(calls)
    assert torch.isinf(lp(torch.full((4, 2), 5.0, dtype=torch.float64))).all() and len(calls) == n


@pytest.mark.gpu
def test_chain_log_probabilities_are_reference_likelihoods():
    import miniframe_b200 as mf
    from oracle import gp_oracle as go
    N = 40
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(N, seed=3)
    gp = mf.BIGgp.BIGgp(mf.kernels.QuasiPeriodic, [None, None, None], t=t, rv=rv, rverr=e1,
                        rhk=rhk, sig_rhk=e2, bis=bis, sig_bis=e3)
    # the box of examples/BIGgp_mcmc.py:48-61, narrowed at the lower end to keep K well conditioned
    lo = np.log([5.0, 0.5, 15.0, 1e-3, 0.1, 0.1, 0.1, 0.1, 0.1])
    hi = np.log([100.0, 2.0, 35.0, 0.1, 10.0, 10.0, 1.0, 10.0, 10.0])
    rng = np.random.default_rng(5)
    p0 = rng.uniform(lo, hi, size=(18, 9))
    s = EnsembleSampler(18, 9, box_log_prob(gp, lo, hi), seed=11, device="cuda")
    p, lnp, _ = s.run_mcmc(p0, 25)
    assert s.chain.shape == (18, 25, 9) and s.chain.is_cuda
    P = p.cpu().numpy()
    assert np.all(P >= lo) and np.all(P <= hi)
    y = go.biggp_y(rv, rhk, bis)
    for w in range(18):
        want = go.biggp_loglike_literal("QuasiPeriodic", np.exp(P[w]), t, y, e1, e2, e3)
        assert abs(float(lnp[w]) - want) <= 1e-9 * abs(want)
    assert float(s.acceptance_fraction.mean()) > 0.05
    # the chain moved and every stored log-probability belongs to its stored position
    assert not torch.equal(s.chain[:, 0], s.chain[:, -1])
    again = box_log_prob(gp, lo, hi)(s.chain[:, 10])
    assert torch.allclose(again, s.lnprobability[:, 10], rtol=1e-12, atol=0)

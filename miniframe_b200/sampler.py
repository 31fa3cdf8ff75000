"""Device-resident affine-invariant ensemble sampler.

The reference drives its likelihood from ``emcee.EnsembleSampler`` one walker at a
time (examples/BIGgp_mcmc.py:87-96: ``logprob(p)`` -> ``gpObj.log_likelihood``).
Here the whole half-ensemble proposed by a stretch move (Goodman & Weare 2010,
the move emcee uses, a = 2) is evaluated by ONE batched likelihood call, and the
walker positions, proposals, acceptance tests and the chain stay on the device;
nothing crosses PCIe between steps.

    logprob = box_log_prob(gp, lo, hi)          # examples/BIGgp_mcmc.py:48-65, batched
    sampler = EnsembleSampler(nwalkers, ndim, logprob, seed=1)
    pos, lnp, _ = sampler.run_mcmc(p0, 1000)    # same call shape as emcee
    sampler.chain            # (nwalkers, nsteps, ndim)
    sampler.lnprobability    # (nwalkers, nsteps)
    sampler.acceptance_fraction

``log_prob_batch`` maps a (B, ndim) tensor to a (B,) tensor on the same device, so
the sampler itself is plain torch and also runs on CPU tensors (the move logic is
tested there); with ``box_log_prob`` the likelihoods come from the sm_100a kernels.
"""
import math

import numpy as np


def _torch():
    import torch
    return torch


def box_log_prob(gp, log_lo, log_hi, transform=None):
    """The reference example's ``logprob`` (examples/BIGgp_mcmc.py:48-65) for a batch.

    Walkers live in log-parameter space; outside the box ``[log_lo, log_hi]`` the
    log-probability is -inf, inside it is ``gp.log_likelihood(exp(p), [])`` (the
    example's log-prior is the constant 0.0).  Rows outside the box are not sent
    to the GPU at all.  ``transform``: optional callable mapping the rows inside the
    box to the argument tuple of ``gp.log_likelihood_batch`` - default
    ``lambda P: (exp(P),)``; e.g. ``lambda P: (exp(P[:, :9]), P[:, 9:])`` samples nine
    log kernel parameters and linear mean parameters (Keplerian orbit ...), and
    ``lambda P: (exp(P[:, :9]), None, exp(P[:, 9:]))`` adds BIGgpPlusJitt's jitters.
    """
    torch = _torch()

    def logprob(P):
        lo = torch.as_tensor(np.asarray(log_lo, dtype=float), device=P.device, dtype=P.dtype)
        hi = torch.as_tensor(np.asarray(log_hi, dtype=float), device=P.device, dtype=P.dtype)
        inside = ((P >= lo) & (P <= hi)).all(dim=1)
        out = torch.full((P.shape[0],), -math.inf, dtype=P.dtype, device=P.device)
        idx = torch.nonzero(inside).flatten()
        if idx.numel():
            Pin = P.index_select(0, idx)
            args = (torch.exp(Pin),) if transform is None else tuple(transform(Pin))
            ll = gp.log_likelihood_batch(*args)
            out.index_copy_(0, idx, ll.to(P.dtype))
        return out

    return logprob


class EnsembleSampler(object):
    """Stretch-move ensemble sampler with emcee's call shape, batched and device-resident."""

    def __init__(self, nwalkers, dim, log_prob_batch, a=2.0, seed=None, device=None):
        if nwalkers % 2 != 0 or nwalkers < 2 * dim:
            # emcee's own requirement (its "parallel" stretch move updates half the ensemble
            # from the other half and needs more walkers than dimensions in each)
            raise ValueError("nwalkers must be even and at least 2 * dim")
        self.k = int(nwalkers)
        self.dim = int(dim)
        self.log_prob_batch = log_prob_batch
        self.a = float(a)
        self.seed = seed
        self.device = device
        self.reset()

    def reset(self):
        self._chain = []
        self._lnprob = []
        self.naccepted = None
        self.iterations = 0
        self._gen = None

    # -- emcee-compatible views ------------------------------------------------
    @property
    def chain(self):
        torch = _torch()
        return torch.stack(self._chain, dim=1) if self._chain else None

    @property
    def lnprobability(self):
        torch = _torch()
        return torch.stack(self._lnprob, dim=1) if self._lnprob else None

    @property
    def flatchain(self):
        c = self.chain
        return c.reshape(-1, self.dim)

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(max(1, self.iterations))

    # -- the move ----------------------------------------------------------------
    def _generator(self, device):
        torch = _torch()
        if self._gen is None:
            self._gen = torch.Generator(device=device)
            if self.seed is None:
                self._gen.seed()
            else:
                self._gen.manual_seed(int(self.seed))
        return self._gen

    def run_mcmc(self, p0, nsteps, lnprob0=None):
        """Advance the ensemble ``nsteps`` steps from ``p0`` (nwalkers, dim).
        Returns (positions, log-probabilities, None) like emcee."""
        torch = _torch()
        dev = self.device
        if dev is None:
            dev = p0.device if hasattr(p0, "device") else ("cuda" if torch.cuda.is_available() else "cpu")
        p = torch.as_tensor(np.asarray(p0, dtype=float) if not hasattr(p0, "device") else p0,
                            dtype=torch.float64, device=dev).clone()
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        gen = self._generator(p.device)
        lnp = self.log_prob_batch(p) if lnprob0 is None else torch.as_tensor(lnprob0, dtype=torch.float64, device=dev).clone()
        if torch.isnan(lnp).any():
            raise ValueError("the initial log-probability is NaN")
        if self.naccepted is None:
            self.naccepted = torch.zeros(self.k, dtype=torch.float64, device=dev)
        half = self.k // 2
        halves = (torch.arange(0, half, device=dev), torch.arange(half, self.k, device=dev))
        for _ in range(int(nsteps)):
            for s in (0, 1):
                mine, other = halves[s], halves[1 - s]
                ns = mine.numel()
                # z ~ g(z) ~ 1/sqrt(z) on [1/a, a]:  z = ((a - 1) u + 1)^2 / a
                u = torch.rand(ns, generator=gen, dtype=torch.float64, device=dev)
                z = ((self.a - 1.0) * u + 1.0) ** 2 / self.a
                partner = other[torch.randint(other.numel(), (ns,), generator=gen, device=dev)]
                x, c = p[mine], p[partner]
                q = c + z[:, None] * (x - c)
                lq = self.log_prob_batch(q)
                lnpdiff = (self.dim - 1.0) * torch.log(z) + lq - lnp[mine]
                accept = torch.log(torch.rand(ns, generator=gen, dtype=torch.float64, device=dev)) < lnpdiff
                accept &= ~torch.isnan(lq)
                p[mine] = torch.where(accept[:, None], q, x)
                lnp[mine] = torch.where(accept, lq, lnp[mine])
                self.naccepted[mine] += accept.to(torch.float64)
            self._chain.append(p.clone())
            self._lnprob.append(lnp.clone())
            self.iterations += 1
        return p, lnp, None

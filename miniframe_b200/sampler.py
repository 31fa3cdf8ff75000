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
    example's log-prior is the constant 0.0).  On CPU tensors the rows outside the box are
    not evaluated at all; on device tensors they are evaluated at the nearest point of the box
    and discarded - a few wasted likelihoods instead of a host synchronisation (``nonzero``)
    per half-step, which is what lets a whole step be captured in a CUDA graph.
    ``transform``: optional callable mapping the rows to the argument tuple of
    ``gp.log_likelihood_batch`` - default ``lambda P: (exp(P),)``; e.g.
    ``lambda P: (exp(P[:, :9]), P[:, 9:])`` samples nine log kernel parameters and linear mean
    parameters (Keplerian orbit ...), and ``lambda P: (exp(P[:, :9]), None, exp(P[:, 9:]))``
    adds BIGgpPlusJitt's jitters.
    """
    torch = _torch()
    bounds = {}

    def logprob(P):
        key = (P.device, P.dtype)
        if key not in bounds:
            bounds[key] = (torch.as_tensor(np.asarray(log_lo, dtype=float), device=P.device, dtype=P.dtype),
                           torch.as_tensor(np.asarray(log_hi, dtype=float), device=P.device, dtype=P.dtype))
        lo, hi = bounds[key]
        inside = ((P >= lo) & (P <= hi)).all(dim=1)
        if P.is_cuda:
            Pin = torch.minimum(torch.maximum(P, lo), hi)
            args = (torch.exp(Pin),) if transform is None else tuple(transform(Pin))
            ll = gp.log_likelihood_batch(*args).to(P.dtype)
            return torch.where(inside, ll, torch.full_like(ll, -math.inf))
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
    """Stretch-move ensemble sampler with emcee's call shape, batched and device-resident.

    One step = two half-ensemble updates, each ONE batched likelihood call.  The step is a fixed
    sequence of device operations with no host synchronisation: the random numbers of a block of
    steps are drawn up front, a device counter selects the step's row, positions / log-probabilities
    / acceptance counts / the chain are updated in place.  On a CUDA device the step is captured
    once in a CUDA graph and replayed, so that small ensembles (the reference example's 18
    walkers) are bound by the factorisation kernel, not by launch overhead.  ``use_graph=False``
    runs the same operations eagerly (CPU tensors always do)."""

    BLOCK = 1024            # steps per block of pre-drawn random numbers / chain storage

    def __init__(self, nwalkers, dim, log_prob_batch, a=2.0, seed=None, device=None, use_graph=True):
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
        self.use_graph = bool(use_graph)
        self.reset()

    def reset(self):
        self._blocks = []       # finished (chain, lnprob) blocks
        self._state = None      # static tensors of the running block
        self._graph = None
        self.naccepted = None
        self.iterations = 0
        self._gen = None

    # -- emcee-compatible views ------------------------------------------------
    def _views(self):
        torch = _torch()
        parts = list(self._blocks)
        st = self._state
        if st is not None and st["filled"] > 0:
            parts.append((st["chain"][:, :st["filled"]], st["lnchain"][:, :st["filled"]]))
        if not parts:
            return None, None
        return torch.cat([c for c, _ in parts], dim=1), torch.cat([l for _, l in parts], dim=1)

    @property
    def chain(self):
        return self._views()[0]

    @property
    def lnprobability(self):
        return self._views()[1]

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

    def _one_step(self):
        """Both half-ensemble updates of one step, in place on the static tensors (Goodman & Weare
        2010; emcee's stretch move with a = 2)."""
        torch = _torch()
        st = self._state
        p, lnp, acc, step = st["p"], st["lnp"], st["acc"], st["step"]
        half = self.k // 2
        R = st["rand"].index_select(0, step)[0]                    # (2, 3, half): this step's uniforms
        for s in (0, 1):
            mine = slice(0, half) if s == 0 else slice(half, self.k)
            other = slice(half, self.k) if s == 0 else slice(0, half)
            # z ~ g(z) ~ 1/sqrt(z) on [1/a, a]:  z = ((a - 1) u + 1)^2 / a
            z = ((self.a - 1.0) * R[s, 0] + 1.0) ** 2 / self.a
            partner = torch.clamp((R[s, 1] * half).to(torch.long), max=half - 1)
            x, c = p[mine], p[other].index_select(0, partner)
            q = c + z[:, None] * (x - c)
            lq = self.log_prob_batch(q)
            lnpdiff = (self.dim - 1.0) * torch.log(z) + lq - lnp[mine]
            accept = (torch.log(R[s, 2]) < lnpdiff) & ~torch.isnan(lq)
            p[mine] = torch.where(accept[:, None], q, x)
            lnp[mine] = torch.where(accept, lq, lnp[mine])
            acc[mine] += accept.to(acc.dtype)
        st["chain"].index_copy_(1, step, p.unsqueeze(1))
        st["lnchain"].index_copy_(1, step, lnp.unsqueeze(1))
        step += 1

    def _new_block(self, dev, p, lnp):
        torch = _torch()
        half = self.k // 2
        gen = self._generator(dev)
        old = self._state
        if old is not None and old["filled"] > 0:
            self._blocks.append((old["chain"][:, :old["filled"]].clone(), old["lnchain"][:, :old["filled"]].clone()))
        if old is None:
            self._state = {
                "p": p.clone(), "lnp": lnp.clone(), "acc": self.naccepted,
                "step": torch.zeros(1, dtype=torch.long, device=dev),
                "rand": torch.empty((self.BLOCK, 2, 3, half), dtype=torch.float64, device=dev),
                "chain": torch.empty((self.k, self.BLOCK, self.dim), dtype=torch.float64, device=dev),
                "lnchain": torch.empty((self.k, self.BLOCK), dtype=torch.float64, device=dev),
                "filled": 0}
        st = self._state
        st["step"].zero_()
        st["filled"] = 0
        # log(u) must be finite: draw on (0, 1]
        st["rand"].copy_(1.0 - torch.rand(st["rand"].shape, generator=gen, dtype=torch.float64, device=dev))

    def _capture(self, dev):
        """Capture one step in a CUDA graph (after two eager warm-up steps on the capture stream,
        which create the stream's library handle and workspace; the state is restored after them)."""
        torch = _torch()
        st = self._state
        keep = {k: st[k].clone() for k in ("p", "lnp", "acc", "step")}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        try:
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._one_step()
                for k, v in keep.items():
                    st[k].copy_(v)
                # the graph's kernels hold pointers into the likelihood workspace of this stream:
                # keep that buffer alive for as long as the graph (the engine may outgrow it)
                try:
                    from ._engine import Engine
                    self._pinned = Engine.get().current_workspace()
                except Exception:
                    self._pinned = None
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self._one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._graph = graph
        except Exception as exc:                      # capture not possible: same operations, eagerly
            torch.cuda.synchronize(dev)
            for k, v in keep.items():
                st[k].copy_(v)
            self._graph = False
            self.graph_error = repr(exc)

    def run_mcmc(self, p0, nsteps, lnprob0=None):
        """Advance the ensemble ``nsteps`` steps from ``p0`` (nwalkers, dim).
        Returns (positions, log-probabilities, None) like emcee."""
        torch = _torch()
        dev = self.device
        if dev is None:
            dev = p0.device if hasattr(p0, "device") else ("cuda" if torch.cuda.is_available() else "cpu")
        dev = torch.device(dev)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        p = torch.as_tensor(np.asarray(p0, dtype=float) if not hasattr(p0, "device") else p0,
                            dtype=torch.float64, device=dev).clone()
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        lnp = self.log_prob_batch(p) if lnprob0 is None else torch.as_tensor(lnprob0, dtype=torch.float64, device=dev).clone()
        if torch.isnan(lnp).any():
            raise ValueError("the initial log-probability is NaN")
        if self.naccepted is None:
            self.naccepted = torch.zeros(self.k, dtype=torch.float64, device=dev)
        if self._state is None:
            self._new_block(dev, p, lnp)
        else:
            self._state["p"].copy_(p)
            self._state["lnp"].copy_(lnp)
        st = self._state
        graphed = dev.type == "cuda" and self.use_graph
        if graphed and self._graph is None:
            self._capture(dev)
        for _ in range(int(nsteps)):
            if st["filled"] == self.BLOCK:
                self._new_block(dev, st["p"], st["lnp"])
            if graphed and self._graph:
                self._graph.replay()
            else:
                self._one_step()
            st["filled"] += 1
            self.iterations += 1
        return st["p"].clone(), st["lnp"].clone(), None

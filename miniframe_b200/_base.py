"""Machinery shared by the three GP classes: mean-function bookkeeping (the
reference's mean_pars protocol), residuals for a batch of mean-parameter sets,
and the calls into the device engine."""
from copy import copy

import numpy as np

from . import _lib
from ._engine import Engine

flatten = lambda l: [item for sublist in l for item in sublist]

NUGGET_VALUE = 0.01     # reference BIGgp.py:276, SMALLgp.py:247


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class GPBase(object):
    """Not part of the public API."""

    # ---- mean functions (reference BIGgp.py:121-164, SMALLgp.py:94-136) ----
    def _init_means(self, means):
        self.means = means
        self._mean_pars = []
        for i, m in enumerate(self.means):
            if m is None:
                continue
            self.means[i] = m.initialize()
            self._mean_pars.append(self.means[i].pars)
        self._mean_pars = flatten(self._mean_pars)

    @property
    def mean_pars_size(self):
        self._mean_pars_size = 0
        for m in self.means:
            if m is not None:
                self._mean_pars_size += m._parsize
        return self._mean_pars_size

    @property
    def mean_pars(self):
        return self._mean_pars

    @mean_pars.setter
    def mean_pars(self, pars):
        pars = list(pars)
        assert len(pars) == self.mean_pars_size
        self._mean_pars = copy(pars)
        for i, m in enumerate(self.means):
            if m is None:
                continue
            for j in range(m._parsize):
                m.pars[j] = pars.pop(0)

    def mean(self):
        N = self.t.size
        m = np.zeros_like(self.tt)
        for i, meanfun in enumerate(self.means):
            if meanfun is None:
                continue
            m[i * N:(i + 1) * N] = meanfun(self.t)
        return m

    def _mean_vector(self, MeanModel, x):
        return MeanModel(self.t)

    # ---- device plumbing ---------------------------------------------------
    @property
    def engine(self):
        return Engine.get()

    def _kernel_id(self, kernel=None):
        kernel = self.kernel if kernel is None else kernel
        name = kernel.__name__
        if name not in _lib.KERNEL_IDS:
            # the reference's _kernel_pars returns None for any other class and the
            # caller then fails on `*None` (BIGgp.py:110-119)
            raise TypeError("kernel class %r is not SquaredExponential or QuasiPeriodic" % name)
        return _lib.KERNEL_IDS[name]

    def _dev(self, key, make):
        """Device copy of the host array ``make()``.  The reference rereads t, y and the errors on
        every call, so the copy is refreshed whenever the host values differ from the ones it was
        made from (reassigned or edited in place); comparing a few thousand doubles costs
        microseconds, an upload would cost a launch.  Non-finite data raise ValueError here, as
        scipy's check_finite does inside the reference's cho_factor / cho_solve (BIGgp.py:306-307)."""
        cache = self.__dict__.setdefault("_device_cache", {})
        host = np.ascontiguousarray(make(), dtype=float)
        hit = cache.get(key)
        dev_index = self.engine.device.index
        if hit is None or hit[2] != dev_index or hit[0].shape != host.shape or not np.array_equal(hit[0], host):
            self._check_finite(host)
            hit = (host.copy(), self.engine.tensor(host), dev_index)
            cache[key] = hit
        return hit[1]

    def _y_flat(self):
        raise NotImplementedError

    def _residual_rows(self, Bm, batch):
        """(1, n) when no mean has parameters, else (batch, n): y - mean(b) per row."""
        y = self._y_flat()
        if self.mean_pars_size == 0:
            if Bm is not None and np.size(Bm) != 0:
                raise AssertionError("mean parameters given but no mean function takes any")
            return (y - self.mean())[None, :]
        if Bm is None:
            raise ValueError("this model has %d mean parameters: pass one row per sample"
                             % self.mean_pars_size)
        Bm = np.asarray(Bm.detach().cpu() if _is_torch(Bm) else Bm, dtype=float)
        if Bm.ndim == 1:
            Bm = np.tile(Bm, (batch, 1))
        assert Bm.shape == (batch, self.mean_pars_size)
        keep = copy(self._mean_pars)
        rows = np.empty((batch, y.size))
        for i in range(batch):
            self.mean_pars = Bm[i]
            rows[i] = y - self.mean()
        self.mean_pars = keep
        return rows

    def _residual_rows_device(self, Bm, batch):
        """(batch, n) residuals on the device, every mean evaluated for the whole batch at once
        (miniframe_b200.means.batch_eval); None when some mean class has no batched form."""
        import torch
        from . import means as _means
        eng = self.engine
        Bm = eng.tensor(Bm)
        if Bm.ndim == 1:
            Bm = Bm[None, :].expand(batch, -1)
        assert tuple(Bm.shape) == (batch, self.mean_pars_size)
        t = self._dev("t", lambda: np.asarray(self.t, dtype=float))
        N = self.t.size
        y = self._dev("y_flat", lambda: np.asarray(self._y_flat(), dtype=float))
        rows = y[None, :].repeat(batch, 1)
        k = 0
        for i, m in enumerate(self.means):
            if m is None:
                continue
            vals = _means.batch_eval(m, t, Bm[:, k:k + m._parsize])
            if vals is None:
                return None
            rows[:, i * N:(i + 1) * N] -= vals
            k += m._parsize
        return rows

    @staticmethod
    def _check_finite(*arrays):
        # scipy's cho_factor(check_finite=True) raises ValueError on NaN/Inf and the
        # reference does not catch it (only LinAlgError -> -inf)
        for x in arrays:
            if x is not None and not np.all(np.isfinite(x)):
                raise ValueError("array must not contain infs or NaNs")

    def _batch_loglike(self, prob, diag_key, diag_make, A, extra, Bm):
        """A: (B, hyper_len) numpy or torch; returns same kind, shape (B,)."""
        eng = self.engine
        want_torch = _is_torch(A)
        hyper = eng.tensor(A)
        if hyper.ndim != 2 or hyper.shape[1] != prob.hyper_len:
            raise ValueError("expected hyper-parameters of shape (B, %d), got %s"
                             % (prob.hyper_len, tuple(hyper.shape)))
        B = hyper.shape[0]
        ex = None
        if prob.extra_len:
            ex = eng.tensor(extra)
            if ex.ndim != 2 or ex.shape != (B, prob.extra_len):
                raise ValueError("expected extra parameters of shape (%d, %d)" % (B, prob.extra_len))
        if self.mean_pars_size == 0:
            # no mean function takes parameters: the residual is the same constant row for every
            # call, keep it on the device (a sampler calls this thousands of times)
            if Bm is not None and np.size(Bm.detach().cpu() if _is_torch(Bm) else Bm) != 0:
                raise AssertionError("mean parameters given but no mean function takes any")
            resid = self._dev("resid_no_mean_pars", lambda: self._residual_rows(None, B))
        else:
            if Bm is None:
                raise ValueError("this model has %d mean parameters: pass one row per sample"
                                 % self.mean_pars_size)
            resid = self._residual_rows_device(Bm, B)
            if resid is None:                       # a mean class without a batched form
                resid = eng.tensor(self._residual_rows(Bm, B))
        t = self._dev("t", lambda: np.asarray(self.t, dtype=float))
        diag = self._dev(diag_key, diag_make)
        ll, info = eng.loglike(prob, t, diag, hyper, ex, resid)
        if want_torch:
            return ll
        return ll.cpu().numpy()

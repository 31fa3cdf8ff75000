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
            vals = _means.batch_eval(m, t, Bm[:, k:k + m._parsize], t_host=self.t)
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

    def _batch_loglike_grad(self, prob, diag_key, diag_make, A, extra, Bm):
        """Log-likelihood and its gradient with respect to the rows of ``A`` (and of ``extra``: the
        jitters); returns (ll (B,), grad (B, hyper_len + extra_len)) of the
        kind of ``A``.  The reference has no gradient; its consumer would be an optimiser on
        minus_log_likelihood (BIGgp.py:315-317).

            d ll / d theta = -1/2 sum W dK/dtheta,   W = K^-1 - alpha alpha^T,   alpha = K^-1 r

        The factor comes from the likelihood path itself (it stays in the workspace, in the
        epoch-interleaved order), alpha and K^-1 from the library's triangular routines on that
        factor, and the contraction with dK/dtheta - which is never formed - from the hand-written
        kernel mf_grad_contract.  Samples whose covariance is not positive definite give -inf and a
        NaN gradient."""
        import ctypes
        import torch
        if prob.framework == _lib.FRAMEWORK_SMALLGP or prob.nugget != 0.0:
            raise NotImplementedError("the gradient is built for BIGgp / BIGgpPlusJitt")
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
            resid = self._dev("resid_no_mean_pars", lambda: self._residual_rows(None, B))
        else:
            if Bm is None:
                raise ValueError("this model has %d mean parameters: pass one row per sample"
                                 % self.mean_pars_size)
            resid = self._residual_rows_device(Bm, B)
            if resid is None:
                resid = eng.tensor(self._residual_rows(Bm, B))
        t = self._dev("t", lambda: np.asarray(self.t, dtype=float))
        diag = self._dev(diag_key, diag_make)
        M, N = prob.n_series, prob.n_epochs
        n = M * N
        r = torch.arange(n, device=eng.device)
        perm = (r % M) * N + r // M                    # interleaved row r = M i + P  <-  P N + i
        npar = prob.hyper_len + prob.extra_len
        ll = torch.empty(B, dtype=torch.float64, device=eng.device)
        grad = torch.empty((B, npar), dtype=torch.float64, device=eng.device)
        info = torch.empty(B, dtype=torch.int32, device=eng.device)
        ld, stride = eng.ld(n), eng.matrix_stride(n)
        free, _ = torch.cuda.mem_get_info(eng.device)
        chunk = max(1, min(B, int(0.2 * free) // (stride * 8 + 3 * n * n * 8)))
        ctas = int(eng.lib.mf_grad_ctas(N))
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            sub = slice(b0, b0 + nb)
            rs = resid if resid.shape[0] == 1 else resid[sub]
            eng.loglike(prob, t, diag, hyper[sub], None if ex is None else ex[sub], rs,
                        out_ll=ll[sub], out_info=info[sub], max_chunk=nb)
            ws = eng.current_workspace()
            view = ws[:nb * stride * 8].view(torch.float64).view(nb, stride // ld, ld)
            L = torch.tril(view[:, :n, :n])
            rp = rs[:, perm].expand(nb, n) if rs.shape[0] == 1 else rs[:, perm]
            alpha = torch.cholesky_solve(rp.unsqueeze(2).contiguous(), L)
            W = torch.cholesky_inverse(L)
            W.baddbmm_(alpha, alpha.transpose(1, 2), alpha=-1.0)
            partial = torch.empty(nb * ctas * 24, dtype=torch.float64, device=eng.device)
            hy = hyper[sub].contiguous()
            exs = None if ex is None else ex[sub].contiguous()
            st = eng.lib.mf_grad_contract(eng.handle, ctypes.byref(prob), nb, eng.ptr(t), eng.ptr(hy), eng.ptr(exs),
                                          eng.ptr(W), n, n * n, eng.ptr(partial),
                                          ctypes.c_void_p(grad.data_ptr() + b0 * npar * 8), eng.stream())
            _lib.check(eng.lib, eng.handle, st, "mf_grad_contract")
            eng.launches += 2
            del L, W, alpha, partial
        grad[info != 0] = float("nan")
        if want_torch:
            return ll, grad
        return ll.cpu().numpy(), grad.cpu().numpy()

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

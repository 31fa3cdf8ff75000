"""Jones et al. (2017) framework on the B200 (reference: miniframe/SMALLgp.py).

M = len(means) series, each a combination of G, dG/dt and d2G/dt2 of one latent
GP, plus an optional extra kernel Z(t) on the diagonal blocks.  Same constructor
and ``compute_matrix`` / ``log_likelihood`` / ``minus_log_likelihood`` as the
reference; ``log_likelihood_batch`` evaluates many parameter sets in one pass.

Reference behaviours kept on purpose:
  Q3  the nugget (K <- 0.99 K + 0.01 diag K) is ON by default here (SMALLgp.py:211, 263);
  Q4  the third-derivative kernels use sin*sin where the fourth uses cos*sin;
  Q5  with a single series the covariance is only diag(yerr^2) (SMALLgp.py:227-243).
"""
import numpy as np

from . import _lib
from ._base import NUGGET_VALUE, GPBase, _is_torch


class SMALLgp(GPBase):
    """
    Parameters
    ----------
    kernel : class            X(t) kernel (QuasiPeriodic or SquaredExponential)
    extrakernel : class|None  Z(t) kernel
    means : list              one mean-function class or None per series
    t : array                 times
    *args                     data1, data1_error, data2, data2_error, ...
    """

    def __init__(self, kernel, extrakernel, means, t, *args):
        self.kernel = kernel
        self.extrakernel = extrakernel
        self.dKdt1, self.dKdt2, self.ddKdt2dt1, self.dddKdt2ddt1, \
            self.dddKddt2dt1, self.ddddKddt2ddt1 = self.kernel.__subclasses__()
        self._init_means(means)
        self.t = t
        self.number_models = len(means)
        self.tt = np.tile(t, self.number_models)
        self.args = args
        self.y = []
        self.yerr = []
        i = -1
        for i, j in enumerate(args):
            if i % 2 == 0:
                self.y.append(j)
            else:
                self.yerr.append(j ** 2)          # squared once, here (SMALLgp.py:47)
        self.y = np.array(self.y)
        self.yerr = np.concatenate(self.yerr)
        assert (i + 1) / 2 == self.number_models, 'Given data and number models dont match'

    # ---- parameter bookkeeping ------------------------------------------------
    def _kernel_pars(self, a):
        if self.kernel.__name__ == 'SquaredExponential':
            l, wn = a[:2]
            return [l, wn]
        elif self.kernel.__name__ == 'QuasiPeriodic':
            lp, le, p, wn = a[:4]
            return [lp, le, p, wn]

    def _extrakernel_pars(self, c):
        if self.extrakernel is None:
            return []
        elif self.extrakernel.__name__ == 'SquaredExponential':
            l, wn = c[:2]
            return [l, wn]
        elif self.extrakernel.__name__ == 'QuasiPeriodic':
            lp, le, p, wn = c[:4]
            return [lp, le, p, wn]

    def _scaling_pars(self, a, position):
        M = self.number_models
        if position == M:
            return a[-3 * M + 3 * (position - 1):]
        return a[-3 * M + 3 * (position - 1): -3 * M + 3 * position]

    def _scaling_extrapars(self, c, position):
        return c[-self.number_models + (position - 1)]

    def _y_flat(self):
        return np.concatenate(self.y)

    def _problem(self, n_epochs, nugget=True):
        M = self.number_models
        if M > _lib.MAX_SERIES:
            raise ValueError("at most %d series are supported" % _lib.MAX_SERIES)
        kid = self._kernel_id()
        p = _lib.Problem()
        p.framework = _lib.FRAMEWORK_SMALLGP
        p.kernel = kid
        p.n_series = M
        p.n_epochs = int(n_epochs)
        p.hyper_len = _lib.KERNEL_NPAR[kid] + 3 * M
        if self.extrakernel is None:
            p.extra_kernel = _lib.KERNEL_NONE
            p.extra_len = 0
        else:
            p.extra_kernel = self._kernel_id(self.extrakernel)
            p.extra_len = _lib.KERNEL_NPAR[p.extra_kernel] + M
        p.nugget = NUGGET_VALUE if nugget else 0.0
        return p

    def _pack(self, prob, A, C):
        """Rows [kernel pars | 3M scaling pars] and [extra pars | M amplitudes]: the
        reference indexes kernel pars from the front and scaling pars from the back."""
        M = self.number_models
        npar = _lib.KERNEL_NPAR[prob.kernel]
        cat = np.concatenate
        if _is_torch(A):
            import torch
            cat = torch.cat
        if A.shape[1] < npar + 3 * M:
            raise ValueError("need at least %d parameters per row in a" % (npar + 3 * M))
        hyper = cat([A[:, :npar], A[:, A.shape[1] - 3 * M:]], 1)
        extra = None
        if prob.extra_len:
            ne = _lib.KERNEL_NPAR[prob.extra_kernel]
            if C is None or C.shape[1] < ne + M:
                raise ValueError("need at least %d parameters per row in c" % (ne + M))
            ccat = np.concatenate           # by C's own array kind, whatever A is
            if _is_torch(C):
                import torch
                ccat = torch.cat
            extra = ccat([C[:, :ne], C[:, C.shape[1] - M:]], 1)
        return hyper, extra

    # ---- covariance -------------------------------------------------------------
    def compute_matrix(self, a, c, yerr=True, nugget=True):
        """The (MN, MN) covariance (reference SMALLgp.py:211-249)."""
        eng = self.engine
        M, N = self.number_models, self.t.size
        if yerr:
            diag = np.asarray(self.yerr, dtype=float)
        else:
            # the reference adds np.diag(1e-12 * identity(N)), an N-vector, to the
            # (MN, MN) matrix: that only broadcasts for a single series
            if M != 1:
                raise ValueError("operands could not be broadcast together with shapes "
                                 "(%d,%d) (%d,%d) " % (M * N, M * N, N, N))
            diag = np.full(N, 1e-12)
        prob = self._problem(N, nugget=nugget)
        A = np.asarray(a, dtype=float).ravel()[None, :]
        C = np.asarray(c, dtype=float).ravel()[None, :] if prob.extra_len else None
        hyper, extra = self._pack(prob, A, C)
        K = eng.build_cov(prob, eng.tensor(np.asarray(self.t, dtype=float)), eng.tensor(diag),
                          eng.tensor(hyper), None if extra is None else eng.tensor(extra), full=True)
        n = M * N
        return K[0, :n, :n].cpu().numpy()

    # ---- likelihood ---------------------------------------------------------------
    def log_likelihood(self, a, b=[], c=[], nugget=True):
        """Marginal log-likelihood (reference SMALLgp.py:252-277)."""
        self.mean_pars = b
        prob = self._problem(self.t.size)
        A = np.asarray(a, dtype=float).ravel()[None, :]
        C = np.asarray(c, dtype=float).ravel()[None, :] if prob.extra_len else None
        hyper, extra = self._pack(prob, A, C)
        self._check_finite(hyper, extra)
        ll = self._batch_loglike(prob, "diag_yerr2", lambda: np.asarray(self.yerr, dtype=float),
                                 hyper, extra,
                                 np.asarray(self._mean_pars, dtype=float)[None, :]
                                 if self.mean_pars_size else None)
        if np.isnan(ll[0]):
            raise ValueError("array must not contain infs or NaNs")
        return float(ll[0])

    def minus_log_likelihood(self, a, b=[], c=[], nugget=True):
        return -self.log_likelihood(a, b, c, nugget=True)

    # ---- prediction -----------------------------------------------
    def predict_gp(self, time, a, b=[], c=[], model=1):
        """Conditional predictive distribution of series ``model`` (1-based) given its own data
        (reference SMALLgp.py:317-368), on the device like BIGgp.predict_gp.  The reference's
        K* is NOT the kii formula evaluated at time - t: its a1 a3 term has the opposite sign
        (+2 a1 a3 ddK instead of kii's -2 a1 a3 ddK, SMALLgp.py:353-358 vs 165-170) and the
        extra kernel is left out (commented, 343-349).  Three covariance builds on the
        concatenated times [time, t] give the pieces: the full kii (K**, K), kii without the
        extra kernel (K* base) and ddK alone (the 4 a1 a3 correction).  No mean is added to
        y_mean (the reference does not), there are no observational errors in K, and a square
        K* (len(time) == len(t)) gets wn^2 (a1 + a2 + a3)^2 on its diagonal."""
        import torch
        M, N = self.number_models, self.t.size
        if M < 2:
            raise NotImplementedError("predict_gp needs at least two series on this path "
                                      "(the single-series covariance build keeps quirk Q5)")
        if not 1 <= int(model) <= M:
            raise IndexError("model must be between 1 and %d" % M)
        p = int(model) - 1
        eng = self.engine
        time = np.asarray(time, dtype=float)
        T = time.size
        self.mean_pars = b
        r = np.array_split(np.concatenate(self.y, axis=0) - self.mean(), M)[p]
        x = np.concatenate([time, np.asarray(self.t, dtype=float)])
        m = x.size
        A = np.asarray(a, dtype=float).ravel()[None, :]
        prob = self._problem(m, nugget=False)
        C = np.asarray(c, dtype=float).ravel()[None, :] if prob.extra_len else None
        hyper, extra = self._pack(prob, A, C)
        self._check_finite(hyper, extra)
        xd, zd = eng.tensor(x), eng.tensor(np.zeros(M * m))

        def block(pr, hy, ex):
            K = eng.build_cov(pr, xd, zd, eng.tensor(hy), None if ex is None else eng.tensor(ex), full=True)
            return K[0, p * m:(p + 1) * m, p * m:(p + 1) * m]

        S1 = block(prob, hyper, extra)
        Kss = S1[:T, :T].clone()
        Kn = S1[T:, T:].clone()
        bare = self._problem(m, nugget=False)
        bare.extra_kernel, bare.extra_len = _lib.KERNEL_NONE, 0
        npar = _lib.KERNEL_NPAR[prob.kernel]
        a1, a2, a3 = (float(v) for v in hyper[0, npar + 3 * p:npar + 3 * p + 3])
        Ks = (S1 if not prob.extra_len else block(bare, hyper, None))[:T, T:].clone()
        only_dd = np.array(hyper, dtype=float)
        only_dd[0, npar:] = 0.0
        only_dd[0, npar + 3 * p + 1] = 1.0                       # a2 = 1: the block is ddK
        Ks += 4.0 * a1 * a3 * block(bare, only_dd, None)[:T, T:]
        wn = float(hyper[0, npar - 1])
        if T == N:
            Ks += wn * wn * (a1 + a2 + a3) ** 2 * torch.eye(T, dtype=torch.float64, device=Ks.device)
        elif min(T, N) == 1:
            # one prediction time (or one observation): the reference's wn^2 diag(diag(ones_like(r)))
            # is 1 x 1 and broadcasts over every entry of K* (kernels.py:296-297)
            Ks += wn * wn * (a1 + a2 + a3) ** 2
        ld, stride = eng.ld(N), eng.matrix_stride(N)
        Kw = torch.zeros((1, stride // ld, ld), dtype=torch.float64, device=Ks.device)
        Kw[0, :N, :N] = Kn
        ll, info = eng.potrf_loglike(Kw, N, eng.tensor(r[None, :]))
        if int(info[0]) != 0:
            raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite"
                                        % int(info[0]))
        z = Kw[0, N, :N]
        V, _ = eng.solve_logdet(Kw, N, Ks.contiguous(), None, shared=True)
        y_mean = V @ z
        y_cov = Kss - V @ V.T
        y_std = torch.sqrt(torch.diagonal(y_cov))
        return y_mean.cpu().numpy(), y_std.cpu().numpy(), y_cov.cpu().numpy()

    def log_likelihood_batch(self, A, Bm=None, C=None):
        """Rows of ``A`` (B, npar + 3M), extra-kernel rows ``C`` (B, ne + M) when an
        extra kernel is set, mean-parameter rows ``Bm`` when the means take any."""
        prob = self._problem(self.t.size)
        if not _is_torch(A):
            A = np.asarray(A, dtype=float)
        if C is not None and not _is_torch(C):
            C = np.asarray(C, dtype=float)
        hyper, extra = self._pack(prob, A, C)
        return self._batch_loglike(prob, "diag_yerr2", lambda: np.asarray(self.yerr, dtype=float),
                                   hyper, extra, Bm)

"""Rajpaul et al. (2015) framework on the B200 (reference: miniframe/BIGgp.py).

Same constructor, same ``compute_matrix`` / ``log_likelihood`` /
``minus_log_likelihood`` signatures and results as the reference class, plus
``log_likelihood_batch`` for many hyper-parameter sets at once (what an MCMC
ensemble asks for, reference examples/BIGgp_mcmc.py:47-64, 87-96).  The block
covariance, the Cholesky factorisation, the solve and the log-determinant all
run in the sm_100a kernels behind include/miniframe_b200.h.

Reference behaviours kept on purpose:
  Q1  white noise wn^2 sits on the diagonal of every block, off-diagonal ones too;
  Q2  y = [rv, bis, rhk] (BIGgp.py:52) while the covariance blocks are ordered
      [rv, rhk, bis] (BIGgp.py:256-265);
  Q3  log_likelihood ignores its ``nugget`` argument (BIGgp.py:300).
"""
import numpy as np

from . import _lib
from ._base import NUGGET_VALUE, GPBase, _is_torch
from .kernels import SquaredExponential


class BIGgp(GPBase):
    """GP over RV, log R'hk and BIS sharing one latent function and its derivative.

    Parameters
    ----------
    kernel : class
        ``kernels.QuasiPeriodic`` or ``kernels.SquaredExponential`` (the class itself)
    means : list
        three mean-function classes or None
    t, rv, rverr, rhk, sig_rhk, bis, sig_bis : arrays
    """
    _framework = _lib.FRAMEWORK_BIGGP

    def __init__(self, kernel, means, t, rv, rverr, rhk, sig_rhk, bis, sig_bis):
        self.kernel = kernel
        self._init_means(means)
        self.dKdt1, self.dKdt2, self.ddKdt2dt1, self.dddKdt2ddt1, \
            self.dddKddt2dt1, self.ddddKddt2ddt1 = self.kernel.__subclasses__()
        self.t = t
        self.rv = rv
        self.rverr = rverr
        self.rhk = rhk
        self.sig_rhk = sig_rhk
        self.bis = bis
        self.sig_bis = sig_bis
        self.L = 2
        self.y = np.concatenate([rv, bis, rhk])
        self.yerr = np.concatenate([rverr, sig_bis, sig_rhk])
        self.tt = np.tile(t, self.L + 1)
        span = np.ptp(np.asarray(self.t))        # ndarray.ptp is gone in NumPy 2
        self.tplot = np.array([self.t + 1.5 * span * i for i in range(self.L + 1)]).flatten()

    @classmethod
    def from_rdb(cls, filename, kernel=SquaredExponential, means=None, **kwargs):
        """Build from an .rdb file with columns t, rv, rverr, bis, rhk, sig_rhk
        (reference BIGgp.py:60-100, whose final call drops ``means`` and cannot
        run; here ``means`` defaults to three None)."""
        skip = kwargs.pop('skiprows', 2)
        kwargs.pop('unpack', None)
        t, rv, rverr, bis, rhk, sig_rhk = np.loadtxt(filename, skiprows=skip, unpack=True, **kwargs)
        t = np.linspace(0, 1, t.size)
        biserr = 2 * rverr
        rv, rverr = scale(rv, rverr)
        bis, sig_bis = scale(bis, biserr)
        rhk, sig_rhk = scale(rhk, sig_rhk)
        means = [None, None, None] if means is None else means
        return cls(kernel, means, t, rv, rverr, rhk, sig_rhk, bis, sig_bis)

    # ---- parameter bookkeeping ------------------------------------------------
    def _kernel_pars(self, a):
        if self.kernel.__name__ == 'SquaredExponential':
            l, wn, _, _, _, _, _ = a
            return [l, wn]
        elif self.kernel.__name__ == 'QuasiPeriodic':
            lp, le, p, wn, _, _, _, _, _ = a
            return [lp, le, p, wn]

    def _scaling_pars(self, a):
        """vc, vr, lc, bc, br"""
        return a[-5:]

    def _y_flat(self):
        return self.y

    def _problem(self, n_epochs, nugget=False, jitter=False):
        kid = self._kernel_id()
        p = _lib.Problem()
        p.framework = _lib.FRAMEWORK_BIGGP_JITT if jitter else _lib.FRAMEWORK_BIGGP
        p.kernel = kid
        p.extra_kernel = _lib.KERNEL_NONE
        p.n_series = 3
        p.n_epochs = int(n_epochs)
        p.hyper_len = _lib.KERNEL_NPAR[kid] + 5
        p.extra_len = 3 if jitter else 0
        p.nugget = NUGGET_VALUE if nugget else 0.0
        return p

    def _hyper_row(self, a):
        a = np.asarray(a, dtype=float).ravel()
        if self._kernel_pars(a) is None:
            raise TypeError("kernel class %r is not supported" % self.kernel.__name__)
        return a[None, :]

    def _diag_yerr(self):
        # block order [rv, rhk, bis]: BIGgp.py:256-258
        return np.concatenate([np.asarray(self.rverr, dtype=float) ** 2 * np.ones(self.t.size),
                               np.asarray(self.sig_rhk, dtype=float) ** 2 * np.ones(self.t.size),
                               np.asarray(self.sig_bis, dtype=float) ** 2 * np.ones(self.t.size)])

    def _diag(self, yerr):
        if yerr:
            return "diag_yerr", self._diag_yerr
        return "diag_1e-12", lambda: np.full(3 * self.t.size, 1e-12)       # BIGgp.py:260-261

    # ---- covariance -------------------------------------------------------------
    def _matrix(self, a, x, diag, nugget=False, jitter=None):
        """Full (3 len(x)) square for times x on the device, back as numpy."""
        eng = self.engine
        x = np.asarray(x, dtype=float)
        prob = self._problem(x.size, nugget=nugget, jitter=jitter is not None)
        hyper = eng.tensor(self._hyper_row(a))
        extra = None if jitter is None else eng.tensor(np.asarray(jitter, dtype=float).ravel()[None, :3])
        K = eng.build_cov(prob, eng.tensor(x), eng.tensor(diag), hyper, extra, full=True)
        n = 3 * x.size
        return K[0, :n, :n].cpu().numpy()

    def _block(self, a, x, p, q):
        n = np.asarray(x).size
        K = self._matrix(a, x, np.zeros(3 * n))
        return np.ascontiguousarray(K[p * n:(p + 1) * n, q * n:(q + 1) * n])

    def k11(self, a, x):
        """Equation 18 of Rajpaul et al. (reference BIGgp.py:172-181)."""
        return self._block(a, x, 0, 0)

    def k22(self, a, x):
        """Equation 19 (BIGgp.py:184-190)."""
        return self._block(a, x, 1, 1)

    def k33(self, a, x):
        """Equation 20 (BIGgp.py:193-202)."""
        return self._block(a, x, 2, 2)

    def k12(self, a, x):
        """Equation 21 (BIGgp.py:205-212)."""
        return self._block(a, x, 0, 1)

    def k13(self, a, x):
        """Equation 22 (BIGgp.py:215-224)."""
        return self._block(a, x, 0, 2)

    def k23(self, a, x):
        """Equation 23 (BIGgp.py:227-234)."""
        return self._block(a, x, 1, 2)

    def compute_matrix(self, a, yerr=True, nugget=False):
        """The (3N, 3N) covariance of equation 24 (reference BIGgp.py:237-278)."""
        _, make = self._diag(yerr)
        return self._matrix(a, self.t, make(), nugget=nugget)

    # ---- likelihood ---------------------------------------------------------------
    def log_likelihood(self, a, b, nugget=True):
        """Marginal log-likelihood for kernel parameters ``a`` and mean
        parameters ``b`` (reference BIGgp.py:281-312).  Returns -inf when the
        covariance is not positive definite."""
        self.mean_pars = b
        hyper = self._hyper_row(a)
        self._check_finite(hyper)
        key, make = self._diag(True)
        ll = self._batch_loglike(self._problem(self.t.size), key, make, hyper, None,
                                 np.asarray(self._mean_pars, dtype=float)[None, :]
                                 if self.mean_pars_size else None)
        if np.isnan(ll[0]):
            raise ValueError("array must not contain infs or NaNs")
        return float(ll[0])

    def minus_log_likelihood(self, a, b, nugget=True):
        return -self.log_likelihood(a, b, nugget=True)

    # ---- gradient (not in the reference) -------------------------
    def log_likelihood_grad_batch(self, A, Bm=None):
        """(ll (B,), d ll / d A (B, 9 | 7)) for the rows of ``A``: what an optimiser or HMC on
        ``minus_log_likelihood`` (reference BIGgp.py:315-317) needs.  See GPBase._batch_loglike_grad."""
        if not _is_torch(A):
            A = np.asarray(A, dtype=float)
        key, make = self._diag(True)
        return self._batch_loglike_grad(self._problem(self.t.size), key, make, A, None, Bm)

    def log_likelihood_grad(self, a, b):
        """(log_likelihood(a, b), its gradient with respect to ``a``)."""
        self.mean_pars = b
        hyper = self._hyper_row(a)
        self._check_finite(hyper)
        ll, g = self.log_likelihood_grad_batch(hyper, np.asarray(self._mean_pars, dtype=float)[None, :]
                                               if self.mean_pars_size else None)
        return float(ll[0]), g[0]

    # ---- sampling (reference BIGgp.py:320-378) --------------------------------------------
    def _draw(self, K, n, size, seed):
        """x = L xi with K = L L^T from the factorisation kernel and xi ~ N(0, I) drawn on the
        device.  The reference draws with scipy's multivariate_normal (eigen-decomposition, numpy's
        global generator): the same distribution, not the same numbers."""
        import torch
        eng = self.engine
        ll, info = eng.potrf_loglike(K, n, None)
        if int(info[0]) != 0:
            raise np.linalg.LinAlgError("the covariance is not positive definite in FP64 "
                                        "(pivot %d)" % int(info[0]))
        L = torch.tril(K[0, :n, :n])
        gen = torch.Generator(device=L.device)
        if seed is None:
            gen.seed()
        else:
            gen.manual_seed(int(seed))
        xi = torch.randn((n, 1 if size is None else int(size)), generator=gen, dtype=torch.float64, device=L.device)
        x = (L @ xi).T.cpu().numpy()
        return x[0] if size is None else x

    def _workspace_copy(self, S, n):
        import torch
        eng = self.engine
        ld, stride = eng.ld(n), eng.matrix_stride(n)
        Kw = torch.zeros((1, stride // ld, ld), dtype=torch.float64, device=S.device)
        Kw[0, :n, :n] = S
        return Kw

    def sample(self, a, size=None, seed=None):
        """Draw from N(0, K) with K = compute_matrix(a) (reference BIGgp.py:320-334)."""
        eng = self.engine
        _, make = self._diag(True)
        n = 3 * self.t.size
        K = eng.build_cov(self._problem(self.t.size), eng.tensor(np.asarray(self.t, dtype=float)),
                          eng.tensor(make()), eng.tensor(self._hyper_row(a)), None, full=False)
        return self._draw(K, n, size, seed)

    def _sample_base(self, a, scaling, blk, size, seed):
        eng = self.engine
        N = self.t.size
        hyper = np.array(self._hyper_row(a), dtype=float)
        npar = _lib.KERNEL_NPAR[self._kernel_id()]
        hyper[0, npar:npar + 5] = scaling
        Kbig = eng.build_cov(self._problem(N), eng.tensor(np.asarray(self.t, dtype=float)),
                             eng.tensor(np.zeros(3 * N)), eng.tensor(hyper), None, full=True)
        S = Kbig[0, blk * N:(blk + 1) * N, blk * N:(blk + 1) * N]
        return self._draw(self._workspace_copy(S, N), N, size, seed)

    def sample_from_G(self, a, size=None, seed=None):
        """Draw from the latent process G(t) at the observation times (reference BIGgp.py:339-357)."""
        return self._sample_base(a, [0.0, 0.0, 1.0, 0.0, 0.0], 1, size, seed)

    def sample_from_Gdot(self, a, size=None, seed=None):
        """Draw from dG/dt at the observation times (reference BIGgp.py:360-378)."""
        return self._sample_base(a, [0.0, 1.0, 0.0, 0.0, 0.0], 0, size, seed)

    # ---- prediction -----------------------------------------------
    _MODELS = {"rv": (0, 0, 0), "rhk": (1, 2, 2), "bis": (2, 1, 1)}   # block, chunk of y, mean slot

    def _predict(self, time, a, b, model, extra_diag=0.0):
        """Conditional predictive distribution of ONE series given its own data
        (reference BIGgp.py:570-641), on the device:
          * one covariance build on the concatenated times [time, t] gives K** (k11(a, time)),
            K* (no white noise: the reference's non-square branch, kernels.py:299-304) and
            K (k11(a, t)) as sub-blocks of the series' diagonal block;
          * the factorisation kernel factors K with the residual as row n (z = L^-1 r);
          * V = L^-1 K*^T is a forward substitution with T right-hand sides on that one factor;
          * mean = V^T z + m(time), cov = K** - V^T V (a plain library GEMM, not a hot path).
        Quirks kept: y is ordered [rv, bis, rhk] but the blocks [rv, rhk, bis] (Q2), so 'rhk'
        reads chunk 2 and mean slot 2, 'bis' chunk 1 and slot 1; the covariance being
        conditioned on carries NO observational errors (only wn^2, and the jitter in PlusJitt);
        when len(time) == len(t) every kernel class adds wn^2 on the diagonal of K* too."""
        if model not in self._MODELS:
            raise UnboundLocalError("model must be 'rv', 'rhk' or 'bis'")   # the reference falls through
        blk, chunk, slot = self._MODELS[model]
        time = np.asarray(time, dtype=float)
        self.mean_pars = b
        if self.means[slot] is None:
            raise TypeError("'NoneType' object is not callable")           # BIGgp.py:609 with no mean
        r = np.array_split(self.y - self.mean(), 3)[chunk]
        mean_val = np.asarray(self.means[slot](time), dtype=float)
        hyper = self._hyper_row(a)
        self._check_finite(hyper)
        npar = _lib.KERNEL_NPAR[self._kernel_id()]
        vc, vr, lc, bc, br = (float(v) for v in hyper[0, npar:npar + 5])
        cw = {0: (vc + vr) ** 2, 1: lc * lc, 2: (bc + br) ** 2}[blk]
        y_mean, y_cov, info = self._conditional(time, hyper, blk, r, cw, extra_diag)
        if info != 0:
            raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite"
                                        % info)                            # scipy cho_factor, uncaught
        import torch
        y_mean = y_mean + self.engine.tensor(mean_val)
        y_std = torch.sqrt(torch.diagonal(y_cov))
        return y_mean.cpu().numpy(), y_std.cpu().numpy(), y_cov.cpu().numpy()

    def _conditional(self, time, hyper, blk, r, square_wn_coef, extra_diag=0.0):
        """mean = K* K^-1 r and cov = K** - K* K^-1 K*^T for the covariance of diagonal block
        ``blk`` under the hyper-parameter row ``hyper``, as device tensors, plus the
        factorisation status (0 = positive definite)."""
        import torch
        eng = self.engine
        T, N = time.size, self.t.size
        x = np.concatenate([time, np.asarray(self.t, dtype=float)])
        m = x.size
        prob = self._problem(m)
        Kbig = eng.build_cov(prob, eng.tensor(x), eng.tensor(np.zeros(3 * m)), eng.tensor(hyper), None, full=True)
        S = Kbig[0, blk * m:(blk + 1) * m, blk * m:(blk + 1) * m]
        Kss = S[:T, :T].clone()
        Ks = S[:T, T:].contiguous()                       # (T, N)
        if extra_diag:
            Kss += extra_diag * torch.eye(T, dtype=torch.float64, device=Kss.device)
        npar = _lib.KERNEL_NPAR[self._kernel_id()]
        wn = float(hyper[0, npar - 1])
        if T == N:
            # square tstar: every kernel-class call adds wn^2 I (kernels.py:296)
            Ks += square_wn_coef * wn * wn * torch.eye(T, dtype=torch.float64, device=Ks.device)
        elif min(T, N) == 1:
            # one prediction time (or one observation): wn^2 diag(diag(ones_like(r))) is 1 x 1 and
            # broadcasts, so every entry of K* gets the white noise (kernels.py:296-297)
            Ks += square_wn_coef * wn * wn
        ld, stride = eng.ld(N), eng.matrix_stride(N)
        Kw = torch.zeros((1, stride // ld, ld), dtype=torch.float64, device=Ks.device)
        Kw[0, :N, :N] = S[T:, T:]
        if extra_diag:
            Kw[0, :N, :N] += extra_diag * torch.eye(N, dtype=torch.float64, device=Ks.device)
        ll, info = eng.potrf_loglike(Kw, N, eng.tensor(np.asarray(r, dtype=float)[None, :]))
        if int(info[0]) != 0:
            return None, None, int(info[0])
        z = Kw[0, N, :N]
        V, _ = eng.solve_logdet(Kw, N, Ks, None, shared=True)               # (T, N): rows L^-1 K*_i
        return V @ z, Kss - V @ V.T, 0

    def _predict_base(self, time, y, a, scaling, blk):
        """predict_G / predict_Gdot: condition the bare kernel (or its second derivative) on
        ``y``.  That matrix is a diagonal block of the device build with unit scaling:
        k22 with lc = 1 is kernel(x), k11 with (vc, vr) = (0, 1) is ddKdt2dt1(x)."""
        time = np.asarray(time, dtype=float)
        hyper = np.array(self._hyper_row(a), dtype=float)
        self._check_finite(hyper)
        npar = _lib.KERNEL_NPAR[self._kernel_id()]
        hyper[0, npar:npar + 5] = scaling
        mean, cov, info = self._conditional(time, hyper, blk, np.asarray(y, dtype=float), 1.0)
        if info != 0:
            print('Not positive definite matrix')          # BIGgp.py:406-408
            return np.zeros_like(time), 0
        print('Positive definite matrix')
        return mean.cpu().numpy(), cov.cpu().numpy()

    def predict_G(self, time, y, a):
        """Prediction of the latent process G(t) from ``y`` (reference BIGgp.py:381-420)."""
        return self._predict_base(time, y, a, [0.0, 0.0, 1.0, 0.0, 0.0], 1)

    def predict_Gdot(self, time, y, a):
        """Prediction of dG/dt from ``y`` with the ddK/dt2dt1 kernel (reference BIGgp.py:423-463)."""
        return self._predict_base(time, y, a, [0.0, 1.0, 0.0, 0.0, 0.0], 0)

    def predict_rv(self, time, a):
        """mean, covariance, std for the RVs (reference BIGgp.py:466-492); the reference combines
        the two predictions LINEARLY in vc, vr (covariance included), and so does this."""
        mu, cov = self.predict_G(time, self.rv, a)
        mudot, covdot = self.predict_Gdot(time, self.rv, a)
        vc, vr, lc, bc, br = self._scaling_pars(a)
        mean = vc * mu + vr * mudot
        covariance = vc * cov + vr * covdot
        return mean, covariance, np.sqrt(np.diag(covariance))

    def predict_rhk(self, time, a):
        """mean, covariance, std for log R'hk (reference BIGgp.py:495-520)."""
        mu, cov = self.predict_G(time, self.rhk, a)
        vc, vr, lc, bc, br = self._scaling_pars(a)
        mean = lc * mu
        covariance = lc * cov
        return mean, covariance, np.sqrt(np.diag(covariance))

    def predict_bis(self, time, a):
        """mean, covariance, std for the BIS (reference BIGgp.py:523-549)."""
        mu, cov = self.predict_G(time, self.bis, a)
        mudot, covdot = self.predict_Gdot(time, self.bis, a)
        vc, vr, lc, bc, br = self._scaling_pars(a)
        mean = bc * mu + br * mudot
        covariance = bc * cov + br * covdot
        return mean, covariance, np.sqrt(np.diag(covariance))

    def predict_gp(self, time, a, b, model='rv'):
        """y_mean, y_std, y_cov of the chosen series at ``time`` (reference BIGgp.py:570-641)."""
        return self._predict(time, a, b, model)

    def log_likelihood_batch(self, A, Bm=None):
        """log_likelihood for every row of ``A`` (B, 9 | 7) in one device pass.
        ``Bm`` (B, mean_pars_size) holds the mean parameters per row when the
        model has any.  numpy in -> numpy out; torch in -> device tensor out.
        Rows whose covariance is not positive definite give -inf."""
        if not _is_torch(A):
            A = np.asarray(A, dtype=float)
        key, make = self._diag(True)
        return self._batch_loglike(self._problem(self.t.size), key, make, A, None, Bm)


def scale(x, xerr, m=None, s=None):
    """Remove the mean and divide by the standard deviation (BIGgp.py:645-651)."""
    m = x.mean() if m is None else m
    s = x.std() if s is None else s
    return (x - m) / s, xerr / s

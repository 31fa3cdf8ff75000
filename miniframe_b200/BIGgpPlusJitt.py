"""Rajpaul framework with one jitter term per series (reference:
miniframe/BIGgpPlusJitt.py, whose class is also called ``BIGgp`` and is not
imported by the package __init__: use ``miniframe_b200.BIGgpPlusJitt.BIGgp``).

``compute_matrix(a, j, ...)`` adds j[0]^2, j[1]^2, j[2]^2 to the diagonals of
K11 (rv), K22 (rhk), K33 (bis) (BIGgpPlusJitt.py:263-269) and
``log_likelihood(a, b, c)`` takes the jitters as ``c`` (BIGgpPlusJitt.py:285-318).
"""
import numpy as np

from . import BIGgp as _plain
from ._base import _is_torch


class BIGgp(_plain.BIGgp):

    def compute_matrix(self, a, j, yerr=True, nugget=False):
        """(3N, 3N) covariance with per-series jitter ``j`` (BIGgpPlusJitt.py:237-282)."""
        _, make = self._diag(yerr)
        return self._matrix(a, self.t, make(), nugget=nugget, jitter=np.asarray(j, dtype=float))

    def log_likelihood(self, a, b, c, nugget=True):
        """Marginal log-likelihood; ``c`` = the three jitters (BIGgpPlusJitt.py:285-318)."""
        self.mean_pars = b
        hyper = self._hyper_row(a)
        jit = np.asarray(c, dtype=float).ravel()[None, :3]
        self._check_finite(hyper, jit)
        key, make = self._diag(True)
        ll = self._batch_loglike(self._problem(self.t.size, jitter=True), key, make, hyper, jit,
                                 np.asarray(self._mean_pars, dtype=float)[None, :]
                                 if self.mean_pars_size else None)
        if np.isnan(ll[0]):
            raise ValueError("array must not contain infs or NaNs")
        return float(ll[0])

    def minus_log_likelihood(self, a, b, c=None, nugget=True):
        """The reference forwards only (a, b) (BIGgpPlusJitt.py:321-323) and so
        raises TypeError; the same happens here unless ``c`` is given."""
        if c is None:
            raise TypeError("log_likelihood() missing 1 required positional argument: 'c'")
        return -self.log_likelihood(a, b, c, nugget=True)

    def log_likelihood_grad_batch(self, A, Bm=None, C=None):
        """(ll (B,), gradient (B, 9 | 7 + 3)): columns follow ``A`` and then the jitters ``C``."""
        if C is None:
            raise TypeError("log_likelihood_grad_batch() needs the jitters C of shape (B, 3)")
        if not _is_torch(A):
            A = np.asarray(A, dtype=float)
        if not _is_torch(C):
            C = np.asarray(C, dtype=float)
        key, make = self._diag(True)
        return self._batch_loglike_grad(self._problem(self.t.size, jitter=True), key, make, A, C, Bm)

    def log_likelihood_grad(self, a, b, c):
        """(log_likelihood(a, b, c), its gradient with respect to [a, c])."""
        self.mean_pars = b
        hyper = self._hyper_row(a)
        jit = np.asarray(c, dtype=float).ravel()[None, :3]
        self._check_finite(hyper, jit)
        ll, g = self.log_likelihood_grad_batch(hyper, np.asarray(self._mean_pars, dtype=float)[None, :]
                                               if self.mean_pars_size else None, jit)
        return float(ll[0]), g[0]

    def sample(self, a, j=None, size=None, seed=None):
        """Draw from N(0, compute_matrix(a, j)).  The reference's sample(a) forwards only ``a`` to
        compute_matrix(a, j) and so raises TypeError (BIGgpPlusJitt.py:326-341); the same happens
        here unless the jitters are given."""
        if j is None:
            raise TypeError("compute_matrix() missing 1 required positional argument: 'j'")
        eng = self.engine
        _, make = self._diag(True)
        n = 3 * self.t.size
        K = eng.build_cov(self._problem(self.t.size, jitter=True), eng.tensor(np.asarray(self.t, dtype=float)),
                          eng.tensor(make()), eng.tensor(self._hyper_row(a)),
                          eng.tensor(np.asarray(j, dtype=float).ravel()[None, :3]), full=False)
        return self._draw(K, n, size, seed)

    def predict_gp(self, time, a, b, c, model='rv'):
        """As BIGgp.predict_gp with the series' jitter^2 added to the covariance being conditioned
        on and to K** (BIGgpPlusJitt.py:576-653).  The jitters are read in the order of y,
        [rv, bis, rhk]: 'rhk' takes c[2] and 'bis' c[1]."""
        c = np.asarray(c, dtype=float).ravel()
        idx = {"rv": 0, "rhk": 2, "bis": 1}.get(model, 0)
        return self._predict(time, a, b, model, extra_diag=float(c[idx]) ** 2)

    def log_likelihood_batch(self, A, Bm=None, C=None):
        """Rows of ``A`` (B, 9 | 7) with jitters ``C`` (B, 3)."""
        if C is None:
            raise TypeError("log_likelihood_batch() needs the jitters C of shape (B, 3)")
        if not _is_torch(A):
            A = np.asarray(A, dtype=float)
        if not _is_torch(C):
            C = np.asarray(C, dtype=float)
        key, make = self._diag(True)
        return self._batch_loglike(self._problem(self.t.size, jitter=True), key, make, A, C, Bm)

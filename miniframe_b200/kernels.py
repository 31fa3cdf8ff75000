"""Kernel classes of the drop-in API (reference: miniframe/kernels.py).

A GP object is built from a kernel CLASS (``kernels.QuasiPeriodic`` or
``kernels.SquaredExponential``); the frameworks select the device code path by
the class ``__name__`` (reference BIGgp.py:112-117) and keep the six derivative
classes they find through ``kernel.__subclasses__()`` in definition order
(reference BIGgp.py:42-43).  Instances stay callable on numpy lags ``r = t - t'``
exactly like the reference's (kernels.py:14-16), white-noise quirk included:
every class, derivatives too, adds ``wn**2`` on the diagonal when ``r`` is a square
2-D array and falls back to the bare function when numpy cannot broadcast that
term (kernels.py:296-304).

The log-likelihood hot path does NOT evaluate these Python callables: the same
closed forms live in csrc/mf_cov.cuh and run on the GPU.
"""
import numpy as np

pi = np.pi


class kernel(object):
    """Base class: stores the positional arguments in ``pars``."""
    is_kernel = True

    def __init__(self, *args):
        self.pars = np.array(args)

    def __call__(self, r):
        raise NotImplementedError

    def __add__(self, b):
        if not hasattr(b, "is_kernel"):
            return Sum(Constant(c=float(b)), self)
        return Sum(self, b)

    def __radd__(self, b):
        return self.__add__(b)

    def __mul__(self, b):
        if not hasattr(b, "is_kernel"):
            return Product(Constant(c=float(b)), self)
        return Product(self, b)

    def __rmul__(self, b):
        return self.__mul__(b)

    def __repr__(self):
        return "{0}({1})".format(self.__class__.__name__, ", ".join(map(str, self.pars)))


class _operator(kernel):
    def __init__(self, k1, k2):
        self.k1 = k1
        self.k2 = k2

    @property
    def pars(self):
        return np.append(self.k1.pars, self.k2.pars)


class Sum(_operator):
    def __repr__(self):
        return "{0} + {1}".format(self.k1, self.k2)

    def __call__(self, r):
        return self.k1(r) + self.k2(r)


class Product(_operator):
    def __repr__(self):
        return "{0} * {1}".format(self.k1, self.k2)

    def __call__(self, r):
        return self.k1(r) * self.k2(r)


class Constant(kernel):
    def __init__(self, c):
        super(Constant, self).__init__(c)
        self.c = c

    def __call__(self, r):
        return self.c * np.ones_like(r)


def _with_white_noise(value, r, wn):
    """value + wn^2 I when numpy can form it for this ``r`` (reference's
    try / except ValueError around every __call__)."""
    try:
        return value + wn ** 2 * np.diag(np.diag(np.ones_like(r)))
    except ValueError:
        return value


# --------------------------------------------------------------------------
# squared exponential (reference kernels.py:80-259)
# --------------------------------------------------------------------------
class SquaredExponential(kernel):
    """exp(-r^2 / (2 ell^2)); parameters ``ell`` (length scale), ``wn`` (white noise)."""

    def __init__(self, ell, wn):
        kernel.__init__(self, ell, wn)
        self.ell = ell
        self.wn = wn

    def _gauss(self, r):
        return np.exp(-0.5 * r ** 2 / self.ell ** 2)

    def _value(self, r):
        return self._gauss(r)

    def __call__(self, r):
        return _with_white_noise(self._value(r), r, self.wn)


class dSE_dt1(SquaredExponential):
    """d/dt1."""
    def _value(self, r):
        return -r / self.ell ** 2 * self._gauss(r)


class dSE_dt2(SquaredExponential):
    """d/dt2."""
    def _value(self, r):
        return r / self.ell ** 2 * self._gauss(r)


class ddSE_dt2dt1(SquaredExponential):
    """d2/dt2dt1."""
    def _value(self, r):
        f2 = self.ell ** 2
        return (1.0 / f2 - r ** 2 / f2 ** 2) * self._gauss(r)


class dddSE_dt2ddt1(SquaredExponential):
    """d3/dt2 dt1^2."""
    def _value(self, r):
        return (r ** 3 / self.ell ** 6 - 3.0 * r / self.ell ** 4) * self._gauss(r)


class dddSE_ddt2dt1(SquaredExponential):
    """d3/dt2^2 dt1."""
    def _value(self, r):
        return (-r ** 3 / self.ell ** 6 + 3.0 * r / self.ell ** 4) * self._gauss(r)


class ddddSE_ddt2ddt1(SquaredExponential):
    """d4/dt2^2 dt1^2."""
    def _value(self, r):
        return (r ** 4 / self.ell ** 8 - 6.0 * r ** 2 / self.ell ** 6
                + 3.0 / self.ell ** 4) * self._gauss(r)


# --------------------------------------------------------------------------
# quasi-periodic (reference kernels.py:263-617)
# --------------------------------------------------------------------------
class QuasiPeriodic(kernel):
    """exp(-2 sin^2(pi r / P) / ell_p^2 - r^2 / (2 ell_e^2)); parameters
    ``ell_e`` (evolutionary time scale), ``ell_p`` (periodic length scale),
    ``period``, ``wn`` (white noise)."""

    def __init__(self, ell_e, ell_p, period, wn):
        kernel.__init__(self, ell_e, ell_p, period, wn)
        self.ell_e = ell_e
        self.ell_p = ell_p
        self.period = period
        self.wn = wn

    def _parts(self, r):
        """sin, cos of (pi r)/P, the envelope E, and A with dK/dt2 = A E."""
        f2, f3, P = self.ell_p ** 2, self.ell_e ** 2, self.period
        s = np.sin(pi * r / P)
        c = np.cos(pi * r / P)
        E = np.exp(-(2.0 * s * s / f2) - 0.5 * r * r / f3)
        A = (4 * pi * s * c) / (f2 * P) + r / f3
        return f2, f3, P, s, c, E, A

    def _value(self, r):
        return self._parts(r)[5]

    def __call__(self, r):
        return _with_white_noise(self._value(r), r, self.wn)


class dQP_dt1(QuasiPeriodic):
    """d/dt1 (Rajpaul et al. 2015, eq. A8)."""
    def _value(self, r):
        _, _, _, _, _, E, A = self._parts(r)
        return -A * E


class dQP_dt2(QuasiPeriodic):
    """d/dt2 (eq. A9)."""
    def _value(self, r):
        _, _, _, _, _, E, A = self._parts(r)
        return A * E


class ddQP_dt2dt1(QuasiPeriodic):
    """d2/dt2dt1 (eq. A10)."""
    def _value(self, r):
        f2, f3, P, s, c, E, A = self._parts(r)
        return ((-A) * A + 1.0 / f3 + 4 * pi * pi * c * c / (f2 * P * P)
                - 4 * pi * pi * s * s / (f2 * P * P)) * E


def _qp_third(k, r):
    # the reference builds j2 from sin*sin (kernels.py:449); kept as shipped
    f2, f3, P, s, c, E, _ = k._parts(r)
    q0 = 1.0 / f3 + 4 * pi * pi * c * c / (f2 * P ** 2) - 4 * pi * pi * s * s / (f2 * P ** 2)
    j2 = r / f3 + 4 * pi * s * s / (f2 * P)
    return (-3.0 * q0 * j2 + j2 ** 3 - 16 * pi ** 3 * c * s / (f2 * P ** 3)) * E


class dddQP_dt2ddt1(QuasiPeriodic):
    """d3/dt2 dt1^2."""
    def _value(self, r):
        return _qp_third(self, r)


class dddQP_ddt2dt1(QuasiPeriodic):
    """d3/dt2^2 dt1."""
    def _value(self, r):
        return -_qp_third(self, r)


class ddddQP_ddt2ddt1(QuasiPeriodic):
    """d4/dt2^2 dt1^2."""
    def _value(self, r):
        f2, f3, P, s, c, E, A = self._parts(r)
        q0 = 1.0 / f3 + 4 * pi * pi * c * c / (f2 * P ** 2) - 4 * pi * pi * s * s / (f2 * P ** 2)
        return (A ** 4 - 6.0 * q0 * A ** 2 + 3.0 * q0 ** 2
                - 64 * pi ** 3 * c * s * A / (f2 * P ** 3)
                - 16 * pi ** 4 * s * s / (f2 * P ** 4)
                + 16 * pi ** 4 * c * c / (f2 * P ** 4)) * E

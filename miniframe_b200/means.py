"""Mean functions of the drop-in API (reference: miniframe/means.py).

Protocol used by the GP classes (reference means.py:18-36, BIGgp.py:36-41,
137-159): a class with ``_parsize``, ``initialize()`` returning an instance with
all parameters 0, a mutable ``pars`` list, and ``__call__(t)``.  Means only
produce the length-n residual r = y - m(t; b): O(N) host work next to the
O(n^3) device work, so they stay in numpy.
"""
from functools import wraps

import numpy as np

__all__ = ['Constant', 'Linear', 'Parabola', 'Cubic', 'Keplerian']


def array_input(f):
    """Hand ``__call__`` an at-least-1-D array."""
    @wraps(f)
    def wrapped(self, t):
        return f(self, np.atleast_1d(t))
    return wrapped


class MeanModel(object):
    _parsize = 0

    def __init__(self, *pars):
        self.pars = list(pars)

    def __repr__(self):
        return "{0}({1})".format(self.__class__.__name__, ", ".join(map(str, self.pars)))

    @classmethod
    def initialize(cls):
        return cls(*([0.] * cls._parsize))

    def __add__(self, b):
        return Sum(self, b)

    def __radd__(self, b):
        return self.__add__(b)


class Sum(MeanModel):
    def __init__(self, m1, m2):
        self.m1, self.m2 = m1, m2

    @property
    def _parsize(self):
        return self.m1._parsize + self.m2._parsize

    @property
    def pars(self):
        return self.m1.pars + self.m2.pars

    def initialize(self):
        return

    def __repr__(self):
        return "{0} + {1}".format(self.m1, self.m2)

    @array_input
    def __call__(self, t):
        return self.m1(t) + self.m2(t)


class Constant(MeanModel):
    """m(t) = c"""
    _parsize = 1

    def __init__(self, c):
        MeanModel.__init__(self, c)

    @array_input
    def __call__(self, t):
        return np.full(t.shape, self.pars[0])


class Linear(MeanModel):
    """m(t) = slope * (t - mean(t)) + intercept (the reference centres on t.mean())"""
    _parsize = 2

    def __init__(self, slope, intercept):
        MeanModel.__init__(self, slope, intercept)

    @array_input
    def __call__(self, t):
        return self.pars[0] * (t - t.mean()) + self.pars[1]


class Parabola(MeanModel):
    """m(t) = quad * t^2 + slope * t + intercept"""
    _parsize = 3

    def __init__(self, quad, slope, intercept):
        MeanModel.__init__(self, quad, slope, intercept)

    @array_input
    def __call__(self, t):
        return np.polyval(self.pars, t)


class Cubic(MeanModel):
    """m(t) = cub * t^3 + quad * t^2 + slope * t + intercept"""
    _parsize = 4

    def __init__(self, cub, quad, slope, intercept):
        MeanModel.__init__(self, cub, quad, slope, intercept)

    @array_input
    def __call__(self, t):
        return np.polyval(self.pars, t)


class Sine(MeanModel):
    """m(t) = amp * sin(2 pi t / P + phase) + displacement"""
    _parsize = 4

    def __init__(self, amp, P, phi, D):
        MeanModel.__init__(self, amp, P, phi, D)

    @array_input
    def __call__(self, t):
        amp, P, phi, D = self.pars
        return amp * np.sin((2 * np.pi * t / P) + phi) + D


def _eccentric_anomaly(M, e, iterations):
    """Fixed-count iteration of the reference (means.py:160-168, 204-212): the
    starting guess is the third-order series, each step is E += (M - M_k) /
    (1 - e cos E) with M_k lagging one step behind."""
    E0 = M + e * np.sin(M) + 0.5 * (e ** 2) * np.sin(2 * M)
    M0 = E0 - e * np.sin(E0)
    for _ in range(iterations):
        aux = M - M0
        E1 = E0 + aux / (1 - e * np.cos(E0))
        M1 = E0 - e * np.sin(E0)
        E0, M0 = E1, M1
    return E0


def _kepler_rv(t, P, K, e, w, T0, iterations=100):
    M = 2 * np.pi * (t - T0) / P
    E = _eccentric_anomaly(M, e, iterations)
    nu = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(E / 2))
    return K * (e * np.cos(w) + np.cos(w + nu))


class oldKeplerian(MeanModel):
    """Keplerian parametrised by the time of periastron T0: (P, K, e, w, T0)."""
    _parsize = 5

    def __init__(self, P, K, e, w, T0):
        MeanModel.__init__(self, P, K, e, w, T0)

    @array_input
    def __call__(self, t):
        P, K, e, w, T0 = self.pars
        return _kepler_rv(t, P, K, e, w, T0)


class Keplerian(MeanModel):
    """Keplerian parametrised by the orbital phase: (P, K, e, w, phi, offset);
    T0 = t[0] - P phi / (2 pi)."""
    _parsize = 6

    def __init__(self, P, K, e, w, phi, offset):
        MeanModel.__init__(self, P, K, e, w, phi, offset)

    @array_input
    def __call__(self, t):
        P, K, e, w, phi, offset = self.pars
        T0 = t[0] - (P * phi) / (2. * np.pi)
        return _kepler_rv(t, P, K, e, w, T0) + offset


class TOI175keplerians(MeanModel):
    """Three Keplerians (b, c, d) plus a constant; planet b iterates 500 times,
    c and d 100 times, as in the reference (means.py:237, 255, 273)."""
    _parsize = 16

    def __init__(self, Pb, Kb, eb, wb, phib, Pc, Kc, ec, wc, phic, Pd, Kd, ed, wd, phid, C):
        MeanModel.__init__(self, Pb, Kb, eb, wb, phib, Pc, Kc, ec, wc, phic,
                           Pd, Kd, ed, wd, phid, C)

    @array_input
    def __call__(self, t):
        p = self.pars
        total = np.full(t.shape, p[-1])
        rv = 0.0
        for k, iters in ((0, 500), (1, 100), (2, 100)):
            P, K, e, w, phi = p[5 * k:5 * k + 5]
            T0 = t[0] - (P * phi) / (2. * np.pi)
            rv = rv + _kepler_rv(t, P, K, e, w, T0, iterations=iters)
        return rv + total


# ---- batched evaluation (one row of parameters per sample) --------------------------------
# The GP classes evaluate a whole ensemble of mean-parameter sets at once.  For the classes of
# this module that is a handful of element-wise torch operations on a (B, N) tensor - on the
# device when the parameters live there - written in the same operation order as the scalar
# numpy ``__call__`` above, so both agree to the last couple of ulp of sin/cos.  Returns None for
# any other class (user-defined means, Sum): the caller then loops over ``__call__`` on the host.
def _kepler_rv_batch(torch, t, P, K, e, w, T0, iterations):
    if t.is_cuda:
        # one kernel for the whole batch and all iterations (csrc/mf_means.cu)
        import ctypes
        from ._engine import Engine
        eng = Engine.get()
        B, N = P.shape[0], t.shape[-1]
        pars = torch.cat([x.expand(B, 1) for x in (P, K, e, w, T0)], dim=1).contiguous()
        out = torch.empty((B, N), dtype=torch.float64, device=t.device)
        st = eng.lib.mf_kepler_rv(eng.handle, B, N, eng.ptr(t.reshape(-1).contiguous()), eng.ptr(pars),
                                  int(iterations), eng.ptr(out), N, 0, eng.stream())
        if st != 0:
            raise RuntimeError("mf_kepler_rv failed: %s" % eng.lib.mf_last_error(eng.handle).decode())
        eng.launches += 1
        return out
    M = 2 * np.pi * (t - T0) / P
    E0 = M + e * torch.sin(M) + 0.5 * (e ** 2) * torch.sin(2 * M)
    M0 = E0 - e * torch.sin(E0)
    for _ in range(iterations):
        aux = M - M0
        E1 = E0 + aux / (1 - e * torch.cos(E0))
        M1 = E0 - e * torch.sin(E0)
        E0, M0 = E1, M1
    nu = 2 * torch.atan(torch.sqrt((1 + e) / (1 - e)) * torch.tan(E0 / 2))
    return K * (e * torch.cos(w) + torch.cos(w + nu))


def batch_eval(mean, t, pars, t_host=None):
    """``mean``: an instance of a class of this module; ``t``: (N,) float64 tensor;
    ``pars``: (B, mean._parsize) tensor on the same device; ``t_host``: the same times as a host
    array when the caller has them (saves a device round trip).  Returns the (B, N) tensor of
    m(t; pars[b]) or None when the class has no batched form."""
    import torch
    cls = type(mean)
    c = [pars[:, k:k + 1] for k in range(pars.shape[1])]
    tt = t[None, :]
    if cls is Constant:
        return c[0].expand(-1, t.numel()).clone()
    if cls is Linear:
        # numpy's mean, as the scalar path
        tmean = float(np.asarray(t.detach().cpu() if t_host is None else t_host, dtype=float).mean())
        return c[0] * (tt - tmean) + c[1]
    if cls in (Parabola, Cubic):
        y = torch.zeros((pars.shape[0], t.numel()), dtype=t.dtype, device=t.device)
        for k in range(pars.shape[1]):                           # np.polyval's Horner order
            y = y * tt + c[k]
        return y
    if cls is Sine:
        return c[0] * torch.sin((2 * np.pi * tt / c[1]) + c[2]) + c[3]
    if cls is oldKeplerian:
        return _kepler_rv_batch(torch, tt, c[0], c[1], c[2], c[3], c[4], 100)
    if cls is Keplerian:
        T0 = t[0] - (c[0] * c[4]) / (2. * np.pi)
        return _kepler_rv_batch(torch, tt, c[0], c[1], c[2], c[3], T0, 100) + c[5]
    if cls is TOI175keplerians:
        total = c[15].expand(-1, t.numel())
        rv = 0.0
        for k, iters in ((0, 500), (1, 100), (2, 100)):
            P, K, e, w, phi = c[5 * k:5 * k + 5]
            T0 = t[0] - (P * phi) / (2. * np.pi)
            rv = rv + _kepler_rv_batch(torch, tt, P, K, e, w, T0, iters)
        return rv + total
    return None

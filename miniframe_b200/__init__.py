"""miniframe_b200: the Gaussian-process log-likelihood hot path of mini-frame
(Rajpaul et al. 2015 / Jones et al. 2017 block covariances, FP64 Cholesky, solve
and log-determinant) rebuilt for the NVIDIA B200.

Drop-in surface, mirroring ``miniframe`` (reference miniframe/__init__.py:19-25):

    from miniframe_b200 import kernels, means
    from miniframe_b200.BIGgp import BIGgp
    from miniframe_b200.BIGgpPlusJitt import BIGgp as BIGgpPlusJitt
    from miniframe_b200.SMALLgp import SMALLgp

Everything numerical on the likelihood path runs in hand-written sm_100a CUDA
behind the C ABI of include/miniframe_b200.h; there is no CPU fallback.
"""
from miniframe_b200 import BIGgp       # noqa: F401

from miniframe_b200 import kernels     # noqa: F401

from miniframe_b200 import means       # noqa: F401

from miniframe_b200 import SMALLgp     # noqa: F401

from miniframe_b200 import BIGgpPlusJitt   # noqa: F401

__version__ = "0.1.0"

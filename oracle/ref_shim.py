"""Import the UNMODIFIED reference: from /root/reference in the build container, else
from the copy ``oracle/make_ref.py`` staged under ``oracle/_ref/`` (git-ignored, travels
to the GPU box).

Test infrastructure.  Used by ``tests/golden/make_golden*.py``, by the CPU tests that
are skipped when no reference tree is present, and by the CPU legs of ``bench.py``.

Two harness-side shims, no edits to the reference:
  1. ``matplotlib`` / ``matplotlib.pyplot`` are not installed, and BIGgp.py:4 /
     SMALLgp.py:4 import them at module scope -> stub modules in sys.modules.
  2. ``BIGgp.__init__`` calls ``self.t.ptp()`` (BIGgp.py:57), removed from
     ndarray in NumPy 2 -> pass ``t`` as an ndarray subclass that has ``ptp``.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

def _find_root():
    here = os.path.dirname(os.path.abspath(__file__))
    for root in (os.environ.get("MINIFRAME_REFERENCE"), "/root/reference", os.path.join(here, "_ref")):
        if root and os.path.isdir(os.path.join(root, "miniframe")):
            return root
    return "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "miniframe"))


class TimeArray(np.ndarray):
    """ndarray with the ``ptp`` method NumPy 2 dropped."""

    def ptp(self, *args, **kwargs):
        return np.ptp(np.asarray(self), *args, **kwargs)


def as_time(t):
    return np.ascontiguousarray(t, dtype=float).view(TimeArray)


_cache = {}


def load():
    """Returns a namespace with kernels, means, BIGgp, BIGgpPlusJitt, SMALLgp
    (the reference's own modules)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    miniframe = importlib.import_module("miniframe")
    # BIGgpPlusJitt is not imported by miniframe/__init__.py:19-25
    jitt = importlib.import_module("miniframe.BIGgpPlusJitt")
    ns = types.SimpleNamespace(
        kernels=miniframe.kernels,
        means=miniframe.means,
        BIGgp=miniframe.BIGgp.BIGgp,
        BIGgpPlusJitt=jitt.BIGgp,
        SMALLgp=miniframe.SMALLgp.SMALLgp,
        datasets=os.path.join(REF_ROOT, "miniframe", "datasets"),
    )
    _cache["ns"] = ns
    return ns

#!/usr/bin/env python3
"""Generate tests/golden/golden_rdb_v1.{rdb,npz}: what the UNMODIFIED reference's
``BIGgp.from_rdb`` (reference BIGgp.py:60-100) and ``scale`` (645-651) compute from a small
.rdb file.  The reference's ``from_rdb`` ends in ``cls(kernel, t, rv, rverr, bis, sig_bis, rhk,
sig_rhk)`` - one argument short for ``BIGgp.__init__`` - so it is run with a recording ``cls``:
every line of its preprocessing executes as shipped and the fixture holds the arguments it hands
to the constructor.

Build container only:  python tests/golden/make_golden_rdb.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


class Recorder(object):
    def __init__(self, *args):
        self.args = args


def main():
    ref = ref_shim.load()
    rng = np.random.default_rng(7)
    n = 23
    t = np.sort(rng.uniform(2455000.0, 2455100.0, n))
    cols = [t, 3.0 + rng.normal(0, 2.5, n), rng.uniform(0.4, 1.1, n), rng.normal(-0.01, 0.02, n),
            -4.9 + rng.normal(0, 0.03, n), rng.uniform(0.002, 0.01, n)]
    path = os.path.join(HERE, "golden_rdb_v1.rdb")
    with open(path, "w") as f:
        f.write("jdb\tvrad\tsvrad\tbis_span\trhk\tsig_rhk\n---\t----\t-----\t--------\t---\t-------\n")
        for row in zip(*cols):
            f.write("\t".join("%.10f" % v for v in row) + "\n")
    import miniframe.BIGgp as ref_module
    with contextlib.redirect_stdout(io.StringIO()):
        rec = ref.BIGgp.from_rdb.__func__(Recorder, path, kernel=ref.kernels.QuasiPeriodic)
    kernel, t1, rv, rverr, bis, sig_bis, rhk, sig_rhk = rec.args
    assert kernel is ref.kernels.QuasiPeriodic
    x = np.array([1.0, 4.0, -2.5, 8.0]); xe = np.array([0.1, 0.2, 0.3, 0.4])
    sx, sxe = ref_module.scale(x, xe)
    np.savez(os.path.join(HERE, "golden_rdb_v1.npz"), t=t1, rv=rv, rverr=rverr, bis=bis, sig_bis=sig_bis,
             rhk=rhk, sig_rhk=sig_rhk, scale_x=x, scale_xerr=xe, scale_out_x=sx, scale_out_xerr=sxe)
    print("wrote golden_rdb_v1.rdb (%d rows) and golden_rdb_v1.npz" % n)


if __name__ == "__main__":
    main()

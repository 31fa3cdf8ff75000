#!/usr/bin/env python3
"""Generate tests/golden/golden_predict_v1.npz: outputs of the UNMODIFIED reference's
BIGgp.predict_gp (reference BIGgp.py:570-641), imported via oracle/ref_shim.py.

Build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_predict.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.gp_oracle import synthetic_series  # noqa: E402


def main():
    ref = ref_shim.load()
    K_, M_ = ref.kernels, ref.means
    arrays, cases = {}, []
    N = 48
    t, (rv, rhk, bis), (e1, e2, e3) = synthetic_series(N, 21)
    arrays["t"], arrays["rv"], arrays["rhk"], arrays["bis"] = t, rv, rhk, bis
    arrays["rverr"], arrays["sig_rhk"], arrays["sig_bis"] = e1, e2, e3
    grids = {"grid37": np.linspace(t.min() - 1.0, t.max() + 2.0, 37),
             "gridN": np.linspace(t.min() + 0.13, t.max() - 0.07, N),      # square tstar: the wn quirk
             "grid_on_data": np.concatenate([t[::6], [t[3] + 0.5]]),       # times that coincide with epochs
             "grid1": np.array([t[5] + 0.31])}                             # ONE time: 1 x 1 wn^2 broadcasts over K*
    for k, v in grids.items():
        arrays[k] = v
    hypers = {"QuasiPeriodic": [20.0, 1.1, 23.0, 0.08, 1.3, 0.4, 0.9, 0.7, 0.5],
              "SquaredExponential": [3.0, 0.11, 1.3, 0.4, 0.9, 0.7, 0.5]}
    b = [0.3, -0.2, 0.01, 1.5]          # Constant (rv), Linear (rhk slot), Constant (bis slot)
    for kname, a in hypers.items():
        for gname in grids:
            for model in ("rv", "rhk", "bis"):
                gp = ref.BIGgp(getattr(K_, kname), [M_.Constant, M_.Linear, M_.Constant],
                               ref_shim.as_time(t),
                               rv, e1, rhk, e2, bis, e3)
                with contextlib.redirect_stdout(io.StringIO()):
                    mean, std, cov = gp.predict_gp(grids[gname], a, b, model=model)
                name = "%s_%s_%s" % (kname, gname, model)
                arrays[name + "_mean"], arrays[name + "_std"], arrays[name + "_cov"] = \
                    np.asarray(mean), np.asarray(std), np.asarray(cov)
                cases.append(dict(name=name, kernel=kname, grid=gname, model=model, a=a, b=b))
    jit = [0.2, 0.05, 0.11]
    for model in ("rv", "rhk", "bis"):
        gp = ref.BIGgpPlusJitt(K_.QuasiPeriodic, [M_.Constant, M_.Linear, M_.Constant],
                               ref_shim.as_time(t), rv, e1, rhk, e2, bis, e3)
        with contextlib.redirect_stdout(io.StringIO()):
            mean, std, cov = gp.predict_gp(grids["grid37"], hypers["QuasiPeriodic"], b, jit, model=model)
        name = "jitt_QuasiPeriodic_grid37_%s" % model
        arrays[name + "_mean"], arrays[name + "_std"], arrays[name + "_cov"] = \
            np.asarray(mean), np.asarray(std), np.asarray(cov)
        cases.append(dict(name=name, kernel="QuasiPeriodic", grid="grid37", model=model,
                          a=hypers["QuasiPeriodic"], b=b, jitter=jit))
    # predict_G / predict_Gdot / predict_rv / predict_rhk / predict_bis (BIGgp.py:381-549)
    for kname, a in hypers.items():
        for gname in ("grid37", "gridN"):
            gp = ref.BIGgp(getattr(K_, kname), [None, None, None], ref_shim.as_time(t),
                           rv, e1, rhk, e2, bis, e3)
            with contextlib.redirect_stdout(io.StringIO()):
                outs = {"G": gp.predict_G(grids[gname], rv, a), "Gdot": gp.predict_Gdot(grids[gname], bis, a),
                        "rv": gp.predict_rv(grids[gname], a), "rhk": gp.predict_rhk(grids[gname], a),
                        "bis": gp.predict_bis(grids[gname], a)}
            for what, out in outs.items():
                name = "base_%s_%s_%s" % (kname, gname, what)
                for k, v in enumerate(out):
                    arrays["%s_%d" % (name, k)] = np.asarray(v)
                cases.append(dict(name=name, framework="base", kernel=kname, grid=gname, what=what, a=a,
                                  nout=len(out)))
    # SMALLgp.predict_gp (SMALLgp.py:317-368): three series in the order rv, bis, rhk
    sa = {"QuasiPeriodic": [20.0, 1.1, 23.0, 0.08, 1.3, 0.4, 0.02, 0.9, 0.2, 0.01, 0.7, 0.5, 0.03]}
    sc = [4.0, 0.05, 0.3, 0.2, 0.4]                      # extra SquaredExponential (l, wn) + 3 amplitudes
    sb = [0.3, -0.2, 0.01]
    for extraname, cc in ((None, []), ("SquaredExponential", sc)):
        for gname in ("grid37", "gridN"):
            for model in (1, 2, 3):
                ek = None if extraname is None else getattr(K_, extraname)
                gp = ref.SMALLgp(K_.QuasiPeriodic, ek, [M_.Constant, None, M_.Linear], t,
                                 rv, e1, bis, e3, rhk, e2)
                with contextlib.redirect_stdout(io.StringIO()):
                    mean, std, cov = gp.predict_gp(grids[gname], sa["QuasiPeriodic"], sb, cc, model=model)
                name = "small_%s_%s_%d" % (extraname or "noextra", gname, model)
                arrays[name + "_mean"], arrays[name + "_std"], arrays[name + "_cov"] = \
                    np.asarray(mean), np.asarray(std), np.asarray(cov)
                cases.append(dict(name=name, framework="SMALLgp", kernel="QuasiPeriodic", extrakernel=extraname,
                                  grid=gname, model=model, a=sa["QuasiPeriodic"], b=sb, c=cc))
    np.savez_compressed(os.path.join(HERE, "golden_predict_v1.npz"), **arrays)
    with open(os.path.join(HERE, "golden_predict_v1.json"), "w") as f:
        json.dump(dict(cases=cases, means=["Constant", "Linear", "Constant"]), f, indent=1)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()

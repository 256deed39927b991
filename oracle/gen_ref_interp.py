#!/usr/bin/env python3
"""Golden vectors of the SK-table interpolation made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
SlakoEqGrid_interNew_ / SlakoEqGrid_interOld_ (dftb/slakoeqgrid.F90:183-300); their callees polyInterUniform, poly5ToZero
and freeCubicSpline are the machine translations of gen_ref_math.py.  TEST INFRASTRUCTURE ONLY.
Outputs -> tests/golden/interp_ref_golden.npz; tests/test_oracle_interp_ref.py compares orc_sk_eval with them."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402
import gen_ref_math  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/slakoeqgrid.F90"


class Grid:
    pass


def routines():
    mref = gen_ref_math.load(os.path.join(HERE, "_ref", "libmathref.so"))
    dp = ctypes.POINTER(ctypes.c_double)

    def poly_inter_uniform(xa, yb, xx):
        xp = np.ascontiguousarray(xa.a if isinstance(xa, fx.FArr) else xa, dtype=np.float64)
        y = np.asfortranarray(yb.a if isinstance(yb, fx.FArr) else yb, dtype=np.float64)     # (nv, nn)
        out = np.zeros(y.shape[0])
        mref.mref_poly_inter_u2(len(xp), xp.ctypes.data_as(dp), y.shape[0], y.ctypes.data_as(dp), float(xx), out.ctypes.data_as(dp))
        return out

    def poly5(y0, y0p, y0pp, xx, dx, invdx):
        return mref.mref_poly5_to_zero(y0, y0p, y0pp, xx, dx, invdx)

    def fcs(y0, y0p, y0pp, dx, ydx, xx):
        y, yp, ypp = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        mref.mref_free_cubic_spline(y0, y0p, y0pp, dx, ydx, xx, ctypes.byref(y), ctypes.byref(yp), ctypes.byref(ypp))
        return y.value, yp.value, ypp.value

    def fcs_call(a):
        pos = [x for x in a if "=" not in x]
        kw = dict(x.split("=") for x in a if "=" in x)
        if len(pos) == 7:                                   # ..., yy
            return "%s = freeCubicSpline(%s)[0]" % (pos[6], ", ".join(pos[:6]))
        return "%s, %s = freeCubicSpline(%s)[1:]" % (kw["yp"].strip(), kw["ypp"].strip(), ", ".join(pos[:6]))

    def err(msg):
        raise RuntimeError(msg)

    calls = {"freeCubicSpline": fcs_call, "error": lambda a: "error(%s)" % a[0]}
    env = dict(polyInterUniform=poly_inter_uniform, poly5ToZero=poly5, freeCubicSpline=fcs, error=err, nInterNew_=8, nInterOld_=3,
               nRightInterNew_=4, nRightInterOld_=2, deltaR_=1e-5, distFudge=1.0, distFudgeOld=0.3)
    src = open(REF).read()
    new = fx.compile_routine(src, "subroutine", "SlakoEqGrid_interNew_", calls, extra_env=env, extra_arrays=("skTab",))[0]
    old = fx.compile_routine(src, "subroutine", "SlakoEqGrid_interOld_", calls, extra_env=env, extra_arrays=("skTab",))[0]
    return new, old


def tables():
    rng = np.random.default_rng(31)
    out = {}
    for name, n_grid, n_int, dist in (("t500", 500, 10, 0.02), ("short", 9, 3, 0.05), ("t40", 40, 20, 0.1)):
        r = (np.arange(1, n_grid + 1) * dist)[:, None]
        tab = np.exp(-0.8 * r) * np.cos(r * (1.0 + np.arange(n_int)[None, :] * 0.3)) * rng.uniform(0.5, 1.5, n_int)[None, :]
        out[name] = (np.asfortranarray(tab), dist)
    return out


def distances(n_grid, dist):
    """through every branch: inside the table, near its ends, the tail, beyond the cutoff, both sides of grid points"""
    end = n_grid * dist
    # not on the list: r <= xp(1) of the interpolation window -- below the first grid point (both methods) and, for the
    # old method, r exactly on a grid point.  There polyInterU2 starts from yp(:, iCl) with iCl = ceiling(0) = 0, an
    # out-of-bounds read in the reference (math/interpolation.F90:205-207); restatements and kernels extrapolate instead.
    pts = [1.2 * dist, 1.5 * dist, 2.71 * dist, 4.0 * dist - 1e-9, 4.0 * dist + 1e-9, 0.3712345 * end, 0.5123 * end,
           end - 5.3 * dist, end - 2.3 * dist, end - 1.5 * dist, end - 1.0001 * dist, end - 0.5 * dist, end - 1e-9, end + 1e-9,
           end + 0.1, end + 0.2999, end + 0.3001, end + 0.31, end + 0.5, end + 0.999999, end + 1.0 + 1e-9, end + 3.0]
    return np.array(pts)


def main():
    if not os.path.exists(REF):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    new, old = routines()
    g = {}
    for name, (tab, dist) in tables().items():
        this = Grid()
        this.nGrid, this.nInteg, this.dist, this.skTab = tab.shape[0], tab.shape[1], dist, fx.FArr(tab)
        rr = distances(tab.shape[0], dist)
        for tag, f in (("new", new), ("old", old)):
            out = np.zeros((len(rr), tab.shape[1]))
            for k, r in enumerate(rr):
                dd = fx.FArr(np.full(tab.shape[1], 7.0))
                f(this, dd, float(r))
                out[k] = dd.a
            g["%s__%s" % (name, tag)] = out
        g[name + "__rr"] = rr
        print(name, "done", np.abs(g[name + "__new"]).max(), np.abs(g[name + "__old"]).max())
    path = os.path.join(ROOT, "tests", "golden", "interp_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Golden vectors of the Fermi filling made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
Efilling, electronCount, electronFill, middleGap (dftb/etemp.F90:59-525) with kahan (math/summation.F90:20-42) and the
machine-translated index_heap_sort.  Fermi distribution (distrib = fillingTypes%Fermi); the Methfessel-Paxton helper calls
are mapped to no-ops (never reached).  TEST INFRASTRUCTURE ONLY.  Outputs -> tests/golden/filling_ref_golden.npz;
tests/test_oracle_filling_ref.py compares oracle/scc_model.py (efilling) with them."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402
import gen_ref_math  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/etemp.F90"
REF_SUM = "/root/reference/src/dftbp/math/summation.F90"


class Enum:
    Fermi, Gaussian, Methfessel = 0, 1, 2


def routines():
    mref = gen_ref_math.load(os.path.join(HERE, "_ref", "libmathref.so"))
    src, ssrc = open(REF).read(), open(REF_SUM).read()
    kah = fx.compile_routine(ssrc, "subroutine", "kahan")[0]

    def kahan(s, x, c):
        loc = kah(s, x, c)
        return loc["summation"], loc["compensator"]

    def heap_sort(indx, arr):
        vals = np.ascontiguousarray(arr, dtype=np.float64)
        idx = np.zeros(len(vals), dtype=np.int32)
        mref.mref_index_heap_sort(len(vals), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                  vals.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 2.220446049250313e-16)
        indx.a[:] = idx

    def err(msg):
        raise RuntimeError(msg)

    calls = {"kahan": lambda a: "%s, %s = kahan(%s)" % (a[0], a[2], ", ".join(a)),
             "Kahan": lambda a: "%s, %s = kahan(%s)" % (a[0], a[2], ", ".join(a)),
             "Aweights": "pass", "hX": "pass", "error": lambda a: "error(%s)" % a[0],
             "index_heap_sort": lambda a: "index_heap_sort(%s)" % ", ".join(a),
             "electronFill": lambda a: "electronFill(%s)" % ", ".join(a)}
    env = dict(kahan=kahan, fillingTypes=Enum, erfcwrap=None, index_heap_sort=heap_sort, error=err, elecTol=1.0e-15, elecTolMax=1.0e-7,
               epsilon2=2.0 * 2.220446049250313e-16, mExpArg=36.043653389117154)
    cnt = fx.compile_routine(src, "function", "electronCount", calls, extra_env=env)[0]
    env["electronCount"] = lambda *a: cnt(*a)["electronCount"]
    gap = fx.compile_routine(src, "function", "middleGap", calls, extra_env=env)[0]
    env["middleGap"] = lambda *a: gap(*a)["middleGap"]
    fill = fx.compile_routine(src, "subroutine", "electronFill", calls, extra_env=env)[0]
    env["electronFill"] = fill
    ef = fx.compile_routine(src, "subroutine", "Efilling", calls, extra_env=env)[0]
    return ef


def run(f, eig, n_el, kT, kw):
    """eig [nLevels, nKpt, nSpin] -> dict(Ef, Ebs, TS, E0, filling)"""
    F = np.asfortranarray
    nl, nk, ns = eig.shape
    ebs, ts, e0 = (fx.FArr(np.full(ns, 7.0)) for _ in range(3))
    filling = fx.FArr(np.full((nl, nk, ns), 7.0, order="F"))
    loc = f(ebs, 0.0, ts, e0, filling, fx.FArr(F(eig)), float(n_el), float(kT), fx.FArr(np.asarray(kw, dtype=np.float64)), 0)
    return dict(Ef=float(loc["Ef"]), Ebs=ebs.a.copy(), TS=ts.a.copy(), E0=e0.a.copy(), filling=filling.a.copy())


def cases():
    rng = np.random.default_rng(2024)
    out = {}
    out["gap_T0"] = (np.array([-0.9, -0.5, -0.45, -0.1, 0.2, 0.3]).reshape(6, 1, 1), 4.0, 3.1668e-9, [1.0])
    out["degenerate_homo"] = (np.array([-1.0, 0.0, 0.0, 1.0]).reshape(4, 1, 1), 2.0, 1.0e-3, [1.0])
    out["metal_warm"] = (np.sort(rng.uniform(-1.0, 1.0, size=(24, 1, 1)), axis=0), 9.5, 0.02, [1.0])
    out["kpoints_spin"] = (np.sort(rng.uniform(-1.0, 1.0, size=(12, 3, 2)), axis=0), 13.25, 0.01, [0.5, 0.25, 0.25])
    out["fractional_cold"] = (np.sort(rng.uniform(-1.0, 1.0, size=(10, 2, 1)), axis=0), 3.7, 1.0e-4, [0.75, 0.25])
    out["empty"] = (np.sort(rng.uniform(-1.0, 1.0, size=(5, 1, 1)), axis=0), 0.0, 0.01, [1.0])
    out["full"] = (np.sort(rng.uniform(-1.0, 1.0, size=(5, 1, 2)), axis=0), 10.0, 0.01, [1.0])
    return out


def main():
    if not os.path.exists(REF):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    f = routines()
    g = {}
    for name, (eig, n_el, kT, kw) in cases().items():
        r = run(f, eig, n_el, kT, kw)
        for k, v in r.items():
            g["%s__%s" % (name, k)] = np.asarray(v)
        print(name, "Ef %.12f" % r["Ef"], "sum f", float((r["filling"] * np.asarray(kw)[None, :, None]).sum()))
    path = os.path.join(ROOT, "tests", "golden", "filling_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path)


if __name__ == "__main__":
    main()

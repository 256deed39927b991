"""The oracle's restatements of the reference's interpolation and tolerance-sort routines
(math/interpolation.F90:31-239 poly5ToZero / freeCubicSpline / polyInterU2, math/sorting.F90:142-199
index_heap_sort) pinned to the reference's own code: oracle/gen_ref_math.py machine-translates those routines
where they lie under /root/reference and compiles oracle/_ref/libmathref.so; its outputs for fixed inputs are
committed as tests/golden/math_ref_golden.npz.  Bit-exact: same operations in the same order."""
import ctypes
import os

import numpy as np
import pytest

from helpers import ROOT
import oracle
from oracle import gen_ref_math

GOLD = os.path.join(ROOT, "tests", "golden", "math_ref_golden.npz")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmathref.so")
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


def _orc():
    lib = oracle.oracle.lib()
    d = ctypes.c_double
    lib.orc_poly5_to_zero.restype = d
    lib.orc_poly5_to_zero.argtypes = [d] * 6
    lib.orc_free_cubic_spline.argtypes = [d] * 6 + [dp, dp, dp]
    lib.orc_poly_inter_u2.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp, d, dp]
    lib.orc_index_heap_sort.argtypes = [ctypes.c_int, ip, dp, d]
    return lib


def _run(lib, prefix, g):
    """evaluate the four routines of `lib` (prefix orc_ or mref_) on the golden inputs"""
    f5 = getattr(lib, prefix + "poly5_to_zero")
    fcs = getattr(lib, prefix + "free_cubic_spline")
    u2 = getattr(lib, prefix + "poly_inter_u2")
    srt = getattr(lib, prefix + "index_heap_sort")
    dx = 0.3
    out = {"p5_out": np.array([f5(a[0], a[1], a[2], x * dx, dx, 1.0 / dx) for a, x in zip(g["p5_in"], g["p5_x"])])}
    fo = np.zeros((len(g["fcs_x"]), 3))
    for k, (a, x) in enumerate(zip(g["fcs_in"], g["fcs_x"])):
        y, yp, ypp = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        fcs(a[0], a[1], a[2], 0.02, a[3], x, ctypes.byref(y), ctypes.byref(yp), ctypes.byref(ypp))
        fo[k] = (y.value, yp.value, ypp.value)
    out["fcs_out"] = fo
    uo = np.zeros((len(g["u2_xp0"]), g["u2_yp"].shape[1]))
    for k in range(len(g["u2_xp0"])):
        xp = g["u2_xp0"][k] + 0.02 * np.arange(8)
        yp = np.asfortranarray(g["u2_yp"][k])
        u2(8, xp.ctypes.data_as(dp), yp.shape[0], yp.ctypes.data_as(dp), xp[0] + 0.02 * g["u2_frac"][k],
           uo[k].ctypes.data_as(dp))
    out["u2_out"] = uo
    so_out, pos = [], 0
    for n in g["sort_len"]:
        v = np.ascontiguousarray(g["sort_vals"][pos:pos + n])
        idx = np.zeros(n, dtype=np.int32)
        srt(int(n), idx.ctypes.data_as(ip), v.ctypes.data_as(dp), 1.0e-10)
        so_out.append(idx)
        pos += n
    out["sort_out"] = np.concatenate(so_out)
    return out


def test_oracle_equals_golden_vectors_of_the_translated_reference():
    g = dict(np.load(GOLD))
    fresh = gen_ref_math.golden_inputs()
    for k, v in fresh.items():          # the committed inputs are the generator's
        assert np.array_equal(g[k], v), k
    got = _run(_orc(), "orc_", g)
    for k in ("p5_out", "fcs_out", "u2_out", "sort_out"):
        assert np.array_equal(got[k], g[k]), k
    # the sort really sorts to within the tolerance and is a permutation
    pos = 0
    for n in g["sort_len"]:
        idx, v = g["sort_out"][pos:pos + n] - 1, g["sort_vals"][pos:pos + n]
        assert sorted(idx.tolist()) == list(range(n))
        assert np.all(np.diff(v[idx]) >= -2.0e-10)      # inside a cluster of near-equal values the order is the algorithm's
        pos += n


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libmathref.so is built only where /root/reference exists")
def test_oracle_equals_translated_reference_on_fresh_inputs():
    ref = gen_ref_math.load(REF_SO)
    rng = np.random.default_rng(99)
    g = gen_ref_math.golden_inputs()
    for trial in range(5):
        for k in ("p5_in", "p5_x", "fcs_in", "u2_yp", "sort_vals"):
            g[k] = g[k] * rng.uniform(0.5, 2.0) + (rng.normal(size=g[k].shape) * 1e-3 if k != "sort_vals" else 0.0)
        a, b = _run(_orc(), "orc_", g), _run(ref, "mref_", g)
        for k in a:
            assert np.array_equal(a[k], b[k]), (trial, k)

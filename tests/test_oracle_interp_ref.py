"""The oracle's SK-table interpolation (orc_sk_eval: SlakoEqGrid_interNew_ / _interOld_ restated) against vectors made by
EXECUTING THE REFERENCE'S OWN SOURCE (oracle/gen_ref_interp.py: dftb/slakoeqgrid.F90:183-300 through oracle/fortran_exec.py,
callees = the machine translations of gen_ref_math.py): a distance scan through every branch of both methods -- inside the
table, next to both ends (window clamping), the 5th-order tail, the free cubic spline of the old method, beyond the cutoff
-- on three tables (500 x 10, a 9-row table, 40 x 20).  Bit for bit."""
import ctypes
import os
import sys

import numpy as np
import pytest

from helpers import oracle, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_interp as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "interp_ref_golden.npz")


@pytest.mark.parametrize("name", ["t500", "short", "t40"])
def test_oracle_interpolation_equals_the_executed_reference_source(name):
    gold = np.load(GOLD)
    tab, dist = G.tables()[name]
    rr = gold[name + "__rr"]
    assert np.array_equal(rr, G.distances(tab.shape[0], dist))
    L = oracle.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    for tag, method in (("new", 2), ("old", 1)):
        ref = gold["%s__%s" % (name, tag)]
        assert np.all(np.isfinite(ref))
        for k, r in enumerate(rr):
            out = np.zeros(tab.shape[1])
            L.orc_sk_eval(tab.shape[0], tab.shape[1], ctypes.c_double(dist), tab.ctypes.data_as(dp), method, ctypes.c_double(float(r)),
                          out.ctypes.data_as(dp))
            assert np.array_equal(out, ref[k]), (tag, float(r), np.abs(out - ref[k]).max())
        assert np.all(ref[-1] == 0.0) and np.any(ref[0] != 0.0)          # beyond the cutoff: zeros


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today():
    import fortran_exec as fx
    gold = np.load(GOLD)
    new, old = G.routines()
    tab, dist = G.tables()["short"]
    this = G.Grid()
    this.nGrid, this.nInteg, this.dist, this.skTab = tab.shape[0], tab.shape[1], dist, fx.FArr(tab)
    for tag, f in (("new", new), ("old", old)):
        for k, r in enumerate(gold["short__rr"]):
            dd = fx.FArr(np.full(tab.shape[1], 7.0))
            f(this, dd, float(r))
            assert np.array_equal(dd.a, gold["short__" + tag][k])

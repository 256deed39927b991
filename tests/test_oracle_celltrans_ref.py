"""getCellTranslations / getLatticePoints (dftb/periodic.F90:227-330) with the lattice point iterator of
type/latpointiter.F90 EXECUTED from the reference's source (oracle/fortran_exec.py) against the oracle's
orc_cell_translations and the product's host routine dftbp_b200_cell_translations (no GPU needed: it takes no context) --
the translation vectors every neighbour list, and the Ewald real-space sum, are built over, in the reference's order.
Needs the reference tree."""
import ctypes
import os
import sys

import numpy as np
import pytest

from helpers import oracle, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402
import gen_ref_ewald as GE  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/periodic.F90"
REF_L = "/root/reference/src/dftbp/type/latpointiter.F90"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_cell_translations_equal_the_executed_reference_source():
    lsrc, psrc = open(REF_L).read(), open(REF).read()
    join = lambda fn: (lambda a: "%s(%s)" % (fn, ", ".join(a)))  # noqa: E731
    arrays = ("latVecs", "ranges", "curPoint")
    env = dict(new_object=GE.new_object)
    lcalls = {"getRanges_": join("getRanges_"), "getNextPoint": join("getNextPoint"),
              "this%getNextPoint": lambda a: "tFinished = TLatPointIter_getNextPoint(this, %s, None)['tFinished']" % a[0]}
    for name in ("getRanges_", "getNextPoint", "TLatPointIter_init", "TLatPointIter_getNextPoint", "TLatPointIter_getAllPoints"):
        env[name] = fx.compile_routine(lsrc, "subroutine", name, lcalls, extra_env=env, extra_arrays=arrays)[0]
    pcalls = {"TLatPointIter_init": lambda a: "%s = new_object('latPointGen'); TLatPointIter_init(%s)" % (a[0], ", ".join(a)),
              "latPointGen%getAllPoints": lambda a: "%s = TLatPointIter_getAllPoints(latPointGen, None)['latticePoints']" % a[0],
              "move_alloc": lambda a: "%s = %s" % (a[1], a[0]),
              "getLatticePoints": lambda a: "%s = getLatticePoints(%s)['latPoint']" % (a[0], ", ".join(["None"] + a[1:])),
              "getHelicalPoints": "pass", "error": "pass"}
    env["getLatticePoints"] = fx.compile_routine(psrc, "subroutine", "getLatticePoints", pcalls, extra_env=env, extra_arrays=arrays)[0]
    gct = fx.compile_routine(psrc, "subroutine", "getCellTranslations", pcalls, extra_env=env, extra_arrays=arrays)[0]
    from dftbplus_b200 import _lib
    L = _lib.load()
    for lat, cutoff in ((np.diag([7.0, 8.0, 9.0]), 10.0), (np.array([[6.0, 0.4, 0.0], [0.3, 7.0, 0.5], [0.0, 0.6, 8.0]]), 15.3),
                        (np.array([[0.0, 5.1, 5.1], [5.1, 0.0, 5.1], [5.1, 5.1, 0.0]]), 9.0), (np.eye(3) * 40.0, 6.0)):
        rec = np.ascontiguousarray(np.linalg.inv(lat).T)
        loc = gct(None, None, fx.FArr(np.asfortranarray(lat.T)), fx.FArr(np.asfortranarray(rec.T)), float(cutoff), None)
        cell_ref, rcell_ref = np.asarray(loc["cellVec"].a if hasattr(loc["cellVec"], "a") else loc["cellVec"]).T, loc["rCellVec"].a.T
        cell, rcell = oracle.cell_translations(lat, cutoff)
        assert np.array_equal(cell, cell_ref) and np.array_equal(rcell, rcell_ref)
        # the product's host routine
        n = ctypes.c_int()
        latc = np.ascontiguousarray(lat)
        dp = ctypes.c_void_p
        assert L.dftbp_b200_cell_translations(dp(latc.ctypes.data), dp(rec.ctypes.data), float(cutoff), 0, ctypes.byref(n), None, None) == 0
        pc, pr = np.zeros((n.value, 3)), np.zeros((n.value, 3))
        assert L.dftbp_b200_cell_translations(dp(latc.ctypes.data), dp(rec.ctypes.data), float(cutoff), n.value, ctypes.byref(n),
                                              dp(pc.ctypes.data), dp(pr.ctypes.data)) == 0
        assert np.array_equal(pc, cell_ref) and np.array_equal(pr, rcell_ref)
        assert np.all(cell_ref[0] == 0.0)

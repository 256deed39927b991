"""foldCoordsToCell, pbc3d branch (dftb/boundarycond.F90:322-396) with invert33 / determinant33 (math/simplealgebra.F90)
EXECUTED from the reference's source (oracle/fortran_exec.py) against the oracle's orc_fold_coords_to_cell -- the routine the
device kernel nl_fold is compared with bit for bit.  Needs the reference tree."""
import os
import sys
import types

import numpy as np
import pytest

from helpers import oracle, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/boundarycond.F90"
REF_ALG = "/root/reference/src/dftbp/math/simplealgebra.F90"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_fold_equals_the_executed_reference_source():
    alg = open(REF_ALG).read()
    det = fx.compile_routine(alg, "function", "determinant33")[0]
    inv = fx.compile_routine(alg, "subroutine", "invertInOut", extra_env=dict(determinant33=lambda m: det(m)["determinant33"]))[0]
    calls = {"invert33": lambda a: "invert33(%s)" % ", ".join(a), "pseudoInv": "pass", "rotate3": "pass"}
    enum = types.SimpleNamespace(cluster=0, pbc3d=1, helical=2)
    fold = fx.compile_routine(open(REF).read(), "subroutine", "foldCoordsToCell", calls,
                              extra_env=dict(invert33=inv, boundaryCondsEnum=enum, zAxis=None), extra_arrays=("nonContactDirs",))[0]
    this = types.SimpleNamespace(iBoundaryCondition=1, areSIContactsPresent=False)
    rng = np.random.default_rng(12)
    for lat in (np.diag([7.0, 8.0, 9.0]), np.array([[6.0, 0.4, 0.0], [0.3, 7.0, 0.5], [0.0, 0.6, 8.0]]),
                np.array([[0.0, 5.1, 5.1], [5.1, 0.0, 5.1], [5.1, 5.1, 0.0]])):
        c = rng.uniform(-30.0, 30.0, (200, 3))
        c[:6] = [lat[0], lat[1] + lat[2], -lat[0], 0.5 * lat[2], lat[0] * (1 - 1e-17), np.zeros(3)]     # on faces and corners
        ref = np.asfortranarray(c.T.copy())
        fold(this, fx.FArr(ref), fx.FArr(np.asfortranarray(lat.T)))                   # latVec(:, i) = i-th lattice vector
        got = oracle.fold_coords_to_cell(c, lat)
        assert np.array_equal(got, ref.T), np.abs(got - ref.T).max()
        frac = np.linalg.solve(lat.T, got.T).T
        assert frac.min() > -1e-12 and frac.max() < 1.0 + 1e-12

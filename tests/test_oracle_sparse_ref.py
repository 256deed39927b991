"""The oracle's sparse-layout routines against vectors made by EXECUTING THE REFERENCE'S OWN SOURCE: oracle/fortran_exec.py
transpiles unpackHS_real / unpackHS_cmplx_kpts / packHS_real / packHS_cmplx_kpts (dftb/sparse2dense.F90), mullikenPerOrbital /
mullikenPerBlock (dftb/populations.F90) and addShift_atom / addShift_block (dftb/shift.F90) statement by statement into
NumPy with Fortran index semantics -- and buildH0 / buildS / buildDiatomicBlocks (dftb/nonscc.F90) with rotateH0
(dftb/sk.F90), and the force contractions derivativeNonSccEuclidian / derivative_blockEuclidean (dftb/forces.F90) with the
finite-difference derivator getFirstDerivFiniteDiff, getNonSCCStress / getBlockStress (dftb/stress.F90) and
fullDensityMatrix_real / fullEnergyDensityMatrix_real (dftb/densitymatrix.F90) the same way; oracle/gen_ref_sparse.py ran them on H2O, the Ce/O cluster (f shells), the 2-atom Si
cell (hundreds of images folding onto the same dense blocks) and bcc Fe/C 128 and wrote tests/golden/sparse_ref_golden.npz.
This pins index order, transposition, the on-site symmetrisation, the Bloch phase convention and accumulate / overwrite
semantics of the oracle (and through the GPU parity tests, of the kernels) to the reference's code."""
import os
import sys

import numpy as np
import pytest

from helpers import oracle, case, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_sparse as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "sparse_ref_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cmp(gold, name, key, got, tol):
    ref = gold["%s__%s" % (name, key)]
    if name in G.SAMPLED:
        got = G.sample(got)
    assert got.shape == ref.shape, (key, got.shape, ref.shape)
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max())
    assert err <= tol * scale, (name, key, err)


@pytest.mark.parametrize("name", G.CASES)
def test_oracle_equals_the_executed_reference_source(name, gold):
    w, sk, geo, model, selfegy = case(name)
    d = G.inputs(name)
    n, nA, mOrb = geo["nOrb"], geo["nAtom"], d["mOrb"]
    # unpack: pure placement (sums only where images fold onto one block) -> exact without folding
    tol_fold = 0.0 if name in ("h2o", "ce_o") else 1e-15
    _cmp(gold, name, "unpack_real", np.asfortranarray(oracle.unpack_hs_real(geo, d["orig"])), tol_fold)
    _cmp(gold, name, "unpack_cplx", np.asfortranarray(oracle.unpack_hs_cplx(geo, d["orig"], d["kpoint"])), 1e-14)
    _cmp(gold, name, "pack_real", oracle.pack_hs_real(geo, d["rho"].copy(), d["square"]), 0.0)
    _cmp(gold, name, "pack_cplx", oracle.pack_hs_cplx(geo, d["rho"].copy(), d["square_c"], d["kpoint"], d["kweight"]), 1e-14)
    qq = oracle.mulliken_per_orbital(model, geo, d["over"], d["rho"])
    _cmp(gold, name, "mulliken_orbital", np.asfortranarray(qq.T[:mOrb]), 1e-14)          # Fortran (mOrb, nAtom)
    qb = oracle.mulliken_per_block(model, geo, d["over"], d["rho"])                       # [nAtom, b, a] -> Fortran (a, b, atom)
    _cmp(gold, name, "mulliken_block", np.asfortranarray(np.transpose(qb, (2, 1, 0))[:mOrb, :mOrb]), 1e-14)
    # buildH0 / buildS: the reference's loops (nonscc.F90), its rotation driver (sk.F90 rotateH0) and its own shell-pair
    # formulas (machine translated), integrals from the oracle's interpolation; the oracle rotates with an independent
    # formulation, hence 1e-15 of the largest element (observed 5e-17) instead of bit equality
    _cmp(gold, name, "build_h0", oracle.build_h0(model, geo, selfegy), 1e-15)
    _cmp(gold, name, "build_s", oracle.build_s(model, geo), 1e-15)
    # forces: derivativeNonSccEuclidian / derivative_blockEuclidean (dftb/forces.F90) with getFirstDerivFiniteDiff
    # (dftb/nonscc.F90); finite differences amplify the last-digit difference of the two rotation formulations by 1 / delta
    g1 = oracle.derivative_shift(model, geo, d["rho"], d["over"])
    _cmp(gold, name, "deriv_nonscc", np.asfortranarray(g1.T), 1e-11)
    dm2 = np.stack([d["rho"], 0.3 * d["over"]])
    g2 = oracle.derivative_shift(model, geo, dm2, d["orig"], shift=d["shift_block"])
    _cmp(gold, name, "deriv_block", np.asfortranarray(g2.T), 1e-11)
    if geo["nAtom"] <= 8:
        # getFirstDerivRichardson (dftb/nonscc.F90:488-600): element-wise extrapolation tables, masked updates, frozen elements
        g3 = oracle.derivative_shift(model, geo, d["rho"], d["over"], diff_type=2)
        _cmp(gold, name, "deriv_nonscc_richardson", np.asfortranarray(g3.T), 1e-11)
    # stress: getNonSCCStress / getBlockStress (dftb/stress.F90); the oracle returns st[ii, jj] = reference st(jj, ii)
    s1 = oracle.stress(model, geo, d["rho"], d["over"], 1000.0)
    _cmp(gold, name, "stress_nonscc", np.asfortranarray(s1.T), 1e-11)
    s2 = oracle.stress(model, geo, dm2, d["orig"], 1000.0, shift=d["shift_block"])
    _cmp(gold, name, "stress_block", np.asfortranarray(s2.T), 1e-11)
    # density matrices: fullDensityMatrix_real / fullEnergyDensityMatrix_real (dftb/densitymatrix.F90), both branches each
    if name not in G.SAMPLED:
        C, eig = gold[name + "__dmin_C"], gold[name + "__dmin_eig"]
        for tag in ("pos", "mixed"):
            fill = gold["%s__dmin_%s_fill" % (name, tag)]
            _cmp(gold, name, "dm_" + tag, np.asfortranarray(oracle.density_matrix(C, fill)), 1e-14)
            e = eig if tag == "mixed" else -np.abs(eig) - 0.1
            _cmp(gold, name, "edm_" + tag, np.asfortranarray(oracle.energy_density_matrix(C, fill, e)), 1e-14)
            Cz = C + 1j * C[::-1, ::-1]
            _cmp(gold, name, "dmc_" + tag, np.asfortranarray(oracle.density_matrix(Cz, fill)), 1e-14)
            _cmp(gold, name, "edmc_" + tag, np.asfortranarray(oracle.energy_density_matrix(Cz, fill, e)), 1e-14)
            # the reference scales the eigenvectors and scales them back: they return to rounding
            assert np.abs(gold["%s__dm_%s_cback" % (name, tag)] - C).max() <= 1e-14 * np.abs(C).max()
    _cmp(gold, name, "mulliken_atom", oracle.mulliken_per_atom(model, geo, d["over"], d["rho"], qq=np.linspace(-1.0, 1.0, nA)), 1e-14)
    _cmp(gold, name, "onsite_population", oracle.onsite_population(geo, d["rho"]), 1e-15)
    for zero in (True, False):
        ham = oracle.add_shift(model, geo, d["ham0"].copy(), d["over"], d["shift_atom"], is_input_zero=zero)
        _cmp(gold, name, "add_shift_atom_%d" % zero, np.asfortranarray(ham.T), 1e-15)
        ham = oracle.add_shift(model, geo, d["ham0"].copy(), d["over"], d["shift_block"], is_input_zero=zero)
        _cmp(gold, name, "add_shift_block_%d" % zero, np.asfortranarray(ham.T), 1e-14)


@pytest.mark.skipif(not os.path.exists(G.REF + "sparse2dense.F90"), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today(gold):
    R = G.routines()
    for name in ("h2o", "si2"):
        out = G.run_case(R, G.inputs(name))
        for k, v in out.items():
            assert np.array_equal(v, gold["%s__%s" % (name, k)]), (name, k)


def test_transpiler_index_semantics():
    """FArr: lower bounds, inclusive sections, column-major reshape, conformance checks."""
    import fortran_exec as fx
    a = fx.FArr(np.arange(12, dtype=float).reshape((3, 4), order="F"), lower=(0, 1))
    assert a[0, 1] == 0.0 and a[2, 4] == 11.0
    assert np.array_equal(a[1:2, 2], [4.0, 5.0])
    assert np.array_equal(fx.f_reshape(np.arange(6.0), [2, 3]), [[0, 2, 4], [1, 3, 5]])
    a[0:1, 1:2] = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert a[1, 1] == 3.0 and a[0, 2] == 2.0
    with pytest.raises(AssertionError):
        a[0:1, 1:2] = np.zeros((3, 2))
    with pytest.raises(AssertionError):
        a[3, 1]
    src = "subroutine t(x, n)\n real(dp), intent(inout) :: X(0:)\n integer :: I\n do i = 0, N - 1\n  if (x(i) /= 0.0_dp) x(I) = x(i) * 2.0_dp\n end do\nend subroutine t"
    f, args, code = fx.compile_routine(src, "subroutine", "t")
    x = np.array([1.0, 0.0, 3.0])
    f(fx.wrap(f, "x", x), 3)
    assert np.array_equal(x, [2.0, 0.0, 6.0])

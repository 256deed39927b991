"""The CUDA path against vectors made by executing the reference's own source (tests/golden/sparse_ref_golden.npz,
neigh_ref_golden.npz, filling_ref_golden.npz; oracle/gen_ref_sparse.py, gen_ref_neigh.py, gen_ref_filling.py): the direct
comparison, without the oracle in between.  Same inputs as the generators (their `inputs()` / `cases()` functions)."""
import os
import sys

import numpy as np
import pytest

from helpers import case, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_sparse as GS  # noqa: E402
import gen_ref_neigh as GN  # noqa: E402
import gen_ref_filling as GF  # noqa: E402
import gen_ref_gamma as GG  # noqa: E402
import gen_ref_ewald as GE  # noqa: E402

pytestmark = pytest.mark.gpu


def _close(got, ref, tol, what):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = float(np.abs(got - ref).max())
    assert err <= tol * max(1.0, float(np.abs(ref).max())), (what, err)


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128"])
def test_kernels_equal_the_executed_reference_source(name):
    from dftbplus_b200.api import HotPath
    gold = np.load(os.path.join(ROOT, "tests", "golden", "sparse_ref_golden.npz"))
    w, sk, geo, model, selfegy = case(name)
    d = GS.inputs(name)
    mOrb, nA, ss = d["mOrb"], geo["nAtom"], geo["sparseSize"]

    def ref(key):
        return gold["%s__%s" % (name, key)]

    def cmp(key, got, tol):
        _close(GS.sample(got) if name in GS.SAMPLED else got, ref(key), tol, name + " " + key)

    hp = HotPath(0)
    try:
        sk.upload(hp)
        hp.set_geometry(geo)
        h0, s = hp.buildH0S(sk.selfegy)
        cmp("build_h0", h0, 1e-13)
        cmp("build_s", s, 1e-13)
        cmp("unpack_real", np.asfortranarray(hp.unpackHS(d["orig"])), 1e-15)
        cmp("unpack_cplx", np.asfortranarray(hp.unpackHS(d["orig"], kpoint=d["kpoint"])), 1e-13)
        cmp("pack_real", hp.packHS(d["rho"].copy(), d["square"].copy(order="F")), 0.0 if name in ("h2o", "ce_o") else 1e-15)
        cmp("pack_cplx", hp.packHS(d["rho"].copy(), d["square_c"].copy(order="F"), kpoint=d["kpoint"], kweight=d["kweight"]), 1e-13)
        qq = hp.mulliken(d["over"], d["rho"])
        cmp("mulliken_orbital", np.asfortranarray(qq.T[:mOrb]), 1e-13)
        cmp("mulliken_atom", hp.mullikenPerAtom(d["over"], d["rho"], qq=np.linspace(-1.0, 1.0, nA)), 1e-13)
        cmp("onsite_population", hp.getOnsitePopulation(d["rho"]), 1e-15)
        for zero in (True, False):
            ham = hp.addShift(d["ham0"].copy(), d["over"], np.ascontiguousarray(d["shift_atom"]), is_input_zero=zero)
            cmp("add_shift_atom_%d" % zero, np.asfortranarray(ham.T), 1e-14)
            ham = hp.addShift(d["ham0"].copy(), d["over"], d["shift_block"], is_input_zero=zero)
            cmp("add_shift_block_%d" % zero, np.asfortranarray(ham.T), 1e-13)
        g1 = hp.derivative_shift(d["rho"], d["over"])
        cmp("deriv_nonscc", np.asfortranarray(g1.T), 1e-9)
        g2 = hp.derivative_shift(np.stack([d["rho"], 0.3 * d["over"]]), d["orig"], shift=d["shift_block"])
        cmp("deriv_block", np.asfortranarray(g2.T), 1e-9)
        if geo["nAtom"] <= 8:
            g3 = hp.derivative_shift(d["rho"], d["over"], diff_type=2)           # Richardson derivator
            cmp("deriv_nonscc_richardson", np.asfortranarray(g3.T), 1e-8)
        st = hp.stress(np.stack([d["rho"], 0.3 * d["over"]]), d["orig"], 1000.0, shift=d["shift_block"])
        cmp("stress_block", np.asfortranarray(np.asarray(st).T), 1e-9)
        if name not in GS.SAMPLED:
            C, eig = ref("dmin_C"), ref("dmin_eig")
            for tag in ("pos", "mixed"):
                fill = ref("dmin_%s_fill" % tag)
                cmp("dm_" + tag, np.asfortranarray(hp.getDensityMatrix(np.asfortranarray(C), fill)), 1e-13)
                e = eig if tag == "mixed" else -np.abs(eig) - 0.1
                cmp("edm_" + tag, np.asfortranarray(hp.getEDensityMatrix(np.asfortranarray(C), fill, e)), 1e-13)
                Cz = np.asfortranarray(C + 1j * C[::-1, ::-1])
                cmp("dmc_" + tag, np.asfortranarray(hp.getDensityMatrix(Cz, fill)), 1e-13)
                cmp("edmc_" + tag, np.asfortranarray(hp.getEDensityMatrix(Cz, fill, e)), 1e-13)
    finally:
        hp.close()


@pytest.mark.parametrize("name", ["si2", "fec128", "crystal", "crystal_symmetric"])
def test_device_neighbour_list_equals_the_executed_reference_source(name):
    from dftbplus_b200.api import HotPath
    gold = np.load(os.path.join(ROOT, "tests", "golden", "neigh_ref_golden.npz"))
    c0, rcell, cutoff, sym = GN.cases()[name]
    lat = {"si2": "si2", "fec128": "fec128"}.get(name)
    if lat:
        latvec = case(lat)[0]["latvec"]
    else:
        latvec = np.diag([2 * 5.42] * 3)
    species = np.array([1 + i % 3 for i in range(len(c0))], dtype=np.int32)
    hp = HotPath(0)
    try:
        hp.set_orbitals([[0, 1, 2], [0, 1], [0]])          # 9 / 4 / 1 orbitals: gen_ref_neigh.orbitals_of
        got = hp.updateNeighbourListAndSpecies(c0, species, cutoff, latvec, adopt=False, symmetric=sym)
    finally:
        hp.close()
    keys = ("coords", "img2CentCell", "iCellVec", "nNeighbour", "iNeighbour", "neighDist2") + (() if sym else ("iPair",))
    for k in keys:
        assert np.array_equal(got[k], gold["%s__%s" % (name, k)]), k          # bit for bit
    if not sym:
        assert got["sparseSize"] == int(gold[name + "__sparseSize"])


def test_device_efilling_equals_the_executed_reference_source():
    from dftbplus_b200.api import HotPath
    gold = np.load(os.path.join(ROOT, "tests", "golden", "filling_ref_golden.npz"))
    hp = HotPath(0)
    try:
        for name, (eig, n_el, kT, kw) in GF.cases().items():
            if n_el == 0.0:
                continue
            e = np.ascontiguousarray(np.transpose(eig, (2, 1, 0)))            # [nSpin][nKpt][nLevels] in memory
            r = hp.Efilling(e, n_el, kT, k_weight=kw)
            fill = np.transpose(r["filling"], (2, 1, 0))
            assert abs(r["Ef"] - float(gold[name + "__Ef"])) <= 1e-12, name
            assert np.abs(fill - gold[name + "__filling"]).max() <= 1e-12, name
            assert np.abs(r["Ebs"] - gold[name + "__Ebs"]).max() <= 1e-12 * max(1.0, np.abs(gold[name + "__Ebs"]).max()), name
    finally:
        hp.close()


@pytest.mark.parametrize("name", ["h2o", "ce_o", "fec128", "fec128r"])
def test_device_short_gamma_equals_the_executed_reference_source(name):
    """dftbp_b200_short_gamma_update / _shifts on the device neighbour list against updateNrOfNeighbours_ /
    updateShortGammaValues_ / buildShifts_ executed from the reference's source."""
    from dftbplus_b200.api import HotPath
    gold = np.load(os.path.join(ROOT, "tests", "golden", "gamma_ref_golden.npz"))
    nl, hub, cuts, n_shell, dq = GG.cases()[name]
    w, sk, geo, model, selfegy = case(name.rstrip("r"))
    cutoff = max(sk.cutoff(), float(cuts.max()))
    hp = HotPath(0)
    try:
        sk.upload(hp)
        g = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], sk_cutoff=sk.cutoff(), adopt=False)
        assert np.array_equal(g["iNeighbour"], nl["iNeighbour"])
        got = hp.shortGammaUpdate(hub["nHubbU"], hub["uniqHubbU"], cuts, fetch=True)
        assert np.array_equal(got["nNeighU"], gold[name + "__nNeighU"])
        gam = got["shortGamma"]
        if name + "__shortGamma_shape" in gold:
            shape = tuple(int(x) for x in gold[name + "__shortGamma_shape"])
            dev = GG.sample(np.ascontiguousarray(gam[:, :shape[1]]))
        else:
            dev = gam[:, :gold[name + "__shortGamma"].shape[1]]
        ref = gold[name + "__shortGamma"]
        assert np.abs(dev - ref).max() <= 1e-13 * np.abs(ref).max()
        sh = hp.shortGammaShifts(hub["iHubbU"], dq)
        _close(sh, gold[name + "__shiftShell"], 1e-12, name + " shiftShell")
    finally:
        hp.close()


@pytest.mark.parametrize("name", ["nacl", "skew", "fec16"])
def test_device_ewald_matrix_equals_the_executed_reference_source(name):
    """dftbp_b200_invr_periodic / _invr_cluster against invRPeriodicSerial / invRClusterSerial executed from the reference's
    source (with its dynamic neighbour list iterator)."""
    from dftbplus_b200.api import HotPath
    from dftbplus_b200 import coulomb as pc
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ewald_ref_golden.npz"))
    c0, lat = GE.cases()[name]
    ew = pc.ewald_setup(lat)
    hp = HotPath(0)
    try:
        got = hp.invRPeriodic(np.ascontiguousarray(c0), ew["rLat"], ew["gLat"], ew["alpha"], ew["volume"], ew["maxR"])
        _close(got, gold[name + "__invr"], 1e-13, name + " invR")
        _close(hp.invRCluster(np.ascontiguousarray(c0)), gold[name + "__invr_cluster"], 1e-15, name + " invR cluster")
    finally:
        hp.close()

"""SURVEY 8(f) N2 on the GPU: updateNeighbourList / getNrOfNeighboursForAll / getSparseDescriptor
through the C ABI (dftbp_b200_update_neighbour_list, _get_neighbour_list, _adopt_neighbour_list)
against the oracle restatement -- integer arrays AND coordinates / squared distances bit for bit
(the reference's arithmetic order, no FMA contraction), including the order inside shells of equal
distance, which fixes the iPair layout.

The same kernels run on the CPU in tests/test_neighbour_list_host.py; this file adds the CUDA
backend (launch shapes, cub scans, copies) and the chain into the pair kernels.  It sorts last on
purpose: it was written after the round's GPU budget was spent, so it runs after the parity
tests that were already measured."""
import numpy as np
import pytest

from helpers import oracle, case, assert_sparse_close
from dftbplus_b200 import workloads

pytestmark = pytest.mark.gpu

ARRAYS = ("coords", "species", "img2CentCell", "iCellVec", "nNeighbour", "iNeighbour", "neighDist2",
          "nNeighbourSK", "iPair", "iAtomStart")


@pytest.fixture(scope="module")
def hp_for():
    from dftbplus_b200.api import HotPath
    cache = {}

    def get(name):
        if name not in cache:
            w, sk, geo, model, selfegy = case(name)
            hp = HotPath(0)
            sk.upload(hp)
            cache[name] = hp
        return cache[name]
    yield get
    for hp in cache.values():
        hp.close()


def assert_same_list(got, ref):
    for k in ("nAllAtom", "mNeigh", "sparseSize"):
        assert got[k] == ref[k], k
    for k in ARRAYS:
        assert got[k].shape == ref[k].shape, k
        assert np.array_equal(got[k], ref[k]), k


def _rcell(w, cutoff):
    return np.zeros((1, 3)) if w["latvec"] is None else oracle.cell_translations(w["latvec"], cutoff)[1]


@pytest.mark.parametrize("name", ["h2o", "si2", "fec128"])
def test_k4_per_atom_populations(name, hp_for):
    """mullikenPerAtom (populations.F90:68-112) and getOnsitePopulation (:1195-1222), SURVEY 8(a) a15."""
    w, sk, geo, model, selfegy = case(name)
    hp = hp_for(name)
    hp.set_geometry(geo)
    rng = np.random.default_rng(23)
    over = oracle.build_s(model, geo)
    rho = over * rng.normal(size=over.shape)
    q0 = rng.normal(size=geo["nAtom"])
    got = hp.mullikenPerAtom(over, rho, qq=q0.copy())
    ref = oracle.mulliken_per_atom(model, geo, over, rho, qq=q0)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    assert np.array_equal(hp.getOnsitePopulation(rho), oracle.onsite_population(geo, rho))   # pure gather + ordered sum


@pytest.mark.parametrize("name", ["h2o", "ce_o", "scan", "si2", "fec128", "si512", "water1002"])
def test_device_neighbour_list_equals_the_oracle(name, hp_for):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_for(name)
    cutoff = sk.cutoff()
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], _rcell(w, cutoff), cutoff, cutoff)
    got = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], adopt=False)
    assert_same_list(got, ref)
    assert np.array_equal(got["cellVec"], geo["cellVec"])


def test_device_list_with_a_longer_list_cutoff_and_repeated_updates(hp_for):
    """mCutOff > skCutOff, and the grow-only buffers reused across updates of different size."""
    w, sk, geo, model, selfegy = case("fec128")
    hp = hp_for("fec128")
    skc = sk.cutoff()
    for cutoff in (14.5, skc, 12.0, 16.0):
        ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], _rcell(w, cutoff), cutoff, skc)
        got = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], sk_cutoff=skc,
                                               adopt=False)
        assert_same_list(got, ref)
        assert got["sparseSize"] == geo["sparseSize"]


def test_perfect_crystal_tie_order_on_the_device(hp_for):
    a = 5.42
    basis = np.array([[0, 0, 0], [.5, .5, .5]])
    cc = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(2)])
    coord0 = (cc[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    species0 = np.array([1 + (i % 3 == 0) for i in range(len(coord0))], dtype=np.int32)
    lat = np.diag([3 * a, 3 * a, 2 * a])
    hp = hp_for("fec128")                           # Fe (spd) / C (sp)
    w, sk, geo, model, selfegy = case("fec128")
    ref = oracle.neighbour_list_full(sk.n_orb_species, coord0, species0, oracle.cell_translations(lat, 11.0)[1],
                                     11.0, 11.0)
    got = hp.updateNeighbourListAndSpecies(coord0, species0, 11.0, lat, adopt=False)
    assert_same_list(got, ref)


def test_device_coordinates_as_cuda_tensors(hp_for):
    """A device-resident driver (MD) passes coord0 / species0 as device pointers."""
    import torch
    w, sk, geo, model, selfegy = case("water1002")
    hp = hp_for("water1002")
    cutoff = sk.cutoff()
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], _rcell(w, cutoff), cutoff, cutoff)
    c = torch.from_numpy(np.ascontiguousarray(w["coord0"])).cuda()
    s = torch.from_numpy(np.ascontiguousarray(w["species0"], dtype=np.int32)).cuda()
    torch.cuda.synchronize()
    got = hp.updateNeighbourListAndSpecies(c, s, cutoff, w["latvec"], adopt=False)
    assert_same_list(got, ref)


@pytest.mark.parametrize("name", ["ce_o", "si2", "fec128"])
def test_adopted_list_feeds_the_pair_kernels(name, hp_for):
    """update -> adopt -> buildH0/buildS/forces on the adopted geometry, against the oracle on the
    oracle's own list (same layout, so the sparse arrays compare element by element)."""
    w, sk, geo, model, selfegy = case(name)
    hp = hp_for(name)
    cutoff = sk.cutoff()
    g = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], adopt=True)
    ogeo = oracle.neighbour_geometry(model, w["coord0"], w["species0"], w["latvec"])
    assert g["sparseSize"] == ogeo["sparseSize"] and hp.num_pairs == int(ogeo["nNeighbourSK"].sum())
    h, s = hp.buildH0S(sk.selfegy)
    assert_sparse_close(h, oracle.build_h0(model, ogeo, selfegy), what="H0 on the adopted list " + name)
    assert_sparse_close(s, oracle.build_s(model, ogeo), what="S on the adopted list " + name)
    rng = np.random.default_rng(3)
    dm, edm = rng.normal(size=(1, g["sparseSize"])), rng.normal(size=g["sparseSize"])
    f = hp.derivative_shift(dm, edm)
    f_ref = oracle.derivative_shift(model, ogeo, dm, edm)
    assert np.abs(f - f_ref).max() < 1e-7 * max(1.0, np.abs(f_ref).max())


@pytest.mark.parametrize("name", ["ce_o", "si2", "fec128", "water1002"])
def test_device_adopted_geometry_equals_the_host_built_one(name, hp_for, monkeypatch):
    """adopt_neighbour_list builds work list, block descriptors, span index and hasDuplicates on the device
    (geo_* kernels); every kernel family that walks them must give what it gives on the geometry that
    set_geometry builds on the host from the same arrays: K2 unpack / pack (duplicates -> atomics, atom
    ranges -> blockStartOfAtom), K4, K6 addShift (span index), SCC forces with diagonal block shifts (span
    pipeline) also as pair shares (loadElementRange read from the device descriptors)."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case(name)
    cutoff = sk.cutoff()
    dev = hp_for(name)
    g = dev.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], adopt=True)
    ogeo = oracle.neighbour_geometry(model, w["coord0"], w["species0"], w["latvec"])
    host = HotPath(0)
    sk.upload(host)
    host.set_geometry(ogeo)
    rng = np.random.default_rng(71)
    n, ss, nA, mOrb = ogeo["nOrb"], ogeo["sparseSize"], ogeo["nAtom"], dev.mOrb
    over = oracle.build_s(model, ogeo)
    rho = rng.normal(size=(2, ss))
    edm = rng.normal(size=ss)
    # K2 (blocks that fold onto one dense block are added with atomics: the order of the sum is not fixed)
    ud, uh = dev.unpackHS(over), host.unpackHS(over)
    assert np.array_equal(ud, uh) if name in ("ce_o", "water1002") else np.abs(ud - uh).max() <= 1e-15
    sq = np.asfortranarray(rng.normal(size=(n, n)))
    assert np.array_equal(dev.packHS(np.zeros(ss), sq), host.packHS(np.zeros(ss), sq))
    a0, a1 = max(1, nA // 3), max(1, 2 * nA // 3)
    assert np.array_equal(dev.packHS(np.zeros(ss), sq, atoms=(a0, a1)), host.packHS(np.zeros(ss), sq, atoms=(a0, a1)))
    # K4
    qd, qh = dev.mulliken(over, rho[0]), host.mulliken(over, rho[0])
    assert np.abs(qd - qh).max() <= 1e-13 * max(1.0, np.abs(qh).max())
    # K6: atom shifts and diagonal block shifts through the span pipeline
    sh_atom = rng.normal(size=(2, nA))
    hd, hh = dev.addShift(np.zeros((2, ss)), over, sh_atom, True), host.addShift(np.zeros((2, ss)), over, sh_atom, True)
    assert np.array_equal(hd, hh)
    sh = np.zeros((2, nA, mOrb, mOrb))
    for s_ in range(2):
        for a in range(nA):
            sh[s_, a][np.diag_indices(mOrb)] = rng.normal(size=mOrb)
    hd, hh = dev.addShift(np.zeros((2, ss)), over, sh, True), host.addShift(np.zeros((2, ss)), over, sh, True)
    assert np.array_equal(hd, hh)
    # K5, SCC form, whole and as three shares
    fd, fh = dev.derivative_shift(rho, edm, shift=sh), host.derivative_shift(rho, edm, shift=sh)
    scale = max(1.0, np.abs(fh).max())
    assert np.abs(fd - fh).max() <= 1e-11 * scale
    parts = np.zeros_like(fd)
    for r in range(3):
        dev.set_pair_partition(r, 3)
        parts += dev.derivative_shift(rho, edm, shift=sh)
    dev.set_pair_partition(0, 1)
    assert np.abs(parts - fh).max() <= 1e-11 * scale
    # the host-side adoption (several ranks, or the A/B switch) is still the same geometry
    assert g["sparseSize"] == ss
    host.close()


def test_too_close_atoms_on_the_device(hp_for):
    from dftbplus_b200.api import DftbB200Error
    w, sk, geo, model, selfegy = case("fec128")
    hp = hp_for("fec128")
    coord0 = w["coord0"].copy()
    coord0[77] = coord0[12] + [0.003, 0.0, -0.002]
    coord0[90] = coord0[31] + [0.0, 0.004, 0.0]
    with pytest.raises(DftbB200Error) as e:
        hp.updateNeighbourListAndSpecies(coord0, w["species0"], sk.cutoff(), w["latvec"])
    assert e.value.code == 6 and "Atoms 78 and 13 too close together" in str(e.value)
    # the context stays usable
    got = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], sk.cutoff(), w["latvec"], adopt=False)
    assert got["sparseSize"] == geo["sparseSize"]


def test_md_like_sequence_of_updates(hp_for):
    """Coordinates drift, the list is rebuilt on the device every step (config 5's pattern)."""
    w, sk, geo, model, selfegy = case("water1002")
    hp = hp_for("water1002")
    rng = np.random.default_rng(8)
    cutoff = sk.cutoff()
    rcell = _rcell(w, cutoff)
    coord0 = w["coord0"].copy()
    for step in range(3):
        coord0 = coord0 + rng.normal(scale=0.02, size=coord0.shape)
        ref = oracle.neighbour_list_full(sk.n_orb_species, coord0, w["species0"], rcell, cutoff, cutoff)
        got = hp.updateNeighbourListAndSpecies(coord0, w["species0"], cutoff, w["latvec"], adopt=True)
        assert_same_list(got, ref)
        s = hp.buildS()
        assert s.shape[0] == ref["sparseSize"] and np.isfinite(s).all()


def test_fold_coords_to_cell_on_the_device(hp_for):
    """foldCoordsToCell (boundarycond.F90:369-380) bit for bit, host array and CUDA tensor."""
    import torch
    hp = hp_for("fec128")
    rng = np.random.default_rng(21)
    for lat in (np.diag([86.8, 86.8, 43.4]), np.array([[10.0, 0, 0], [2.0, 9.0, 0], [0, 1.0, 12.0]]),
                0.5 * 10.263 * np.array([[0.0, 1, 1], [1, 0, 1], [1, 1, 0]])):
        frac = rng.uniform(-3.0, 3.0, (700, 3))
        frac[:20] = rng.integers(-2, 3, (20, 3))
        coord = frac @ lat
        ref = oracle.fold_coords_to_cell(coord, lat)
        assert np.array_equal(hp.foldCoordsToCell(coord, lat), ref)
        t = torch.from_numpy(np.ascontiguousarray(coord)).cuda()
        torch.cuda.synchronize()
        hp.foldCoordsToCell(t, lat)
        hp.synchronize()
        assert np.array_equal(t.cpu().numpy(), ref)
    from dftbplus_b200.api import DftbB200Error
    with pytest.raises(DftbB200Error):
        hp.foldCoordsToCell(np.zeros((2, 3)), np.array([[1.0, 0, 0], [2.0, 0, 0], [0, 0, 1.0]]))


def test_md_step_from_unfolded_device_coordinates(hp_for):
    """The whole handleCoordinateChange chain on device-resident coordinates: fold -> list -> adopt ->
    H0/S, against the oracle's chain on the same (drifted, partly outside the cell) coordinates."""
    import torch
    w, sk, geo, model, selfegy = case("fec128")
    hp = hp_for("fec128")
    rng = np.random.default_rng(4)
    lat = np.asarray(w["latvec"])
    coord = w["coord0"] + rng.normal(scale=0.05, size=w["coord0"].shape) + rng.integers(-1, 2, (len(w["coord0"]), 3)) @ lat
    folded = oracle.fold_coords_to_cell(coord, lat)
    ogeo = oracle.neighbour_geometry(model, folded, w["species0"], lat)
    t = torch.from_numpy(np.ascontiguousarray(coord)).cuda()
    torch.cuda.synchronize()
    hp.foldCoordsToCell(t, lat)
    g = hp.updateNeighbourListAndSpecies(t, w["species0"], sk.cutoff(), lat, adopt=True)
    assert g["sparseSize"] == ogeo["sparseSize"] and np.array_equal(g["iPair"][:, :ogeo["mNeigh"] + 1], ogeo["iPair"])
    assert np.array_equal(g["coords"], ogeo["coords"])
    h, s = hp.buildH0S(sk.selfegy)
    assert_sparse_close(h, oracle.build_h0(model, ogeo, selfegy), what="H0 after fold + device list")
    assert_sparse_close(s, oracle.build_s(model, ogeo), what="S after fold + device list")


def test_md_steps_device_vs_oracle():
    """BASELINE config 5's pattern (MD with SCC restart per step) on a 24-atom water cluster, 4 steps:
    the device arm rebuilds the neighbour list on the GPU every step (update + adopt) and runs the
    whole chain through the C ABI; identical SCC iteration counts in every step, energies within
    1e-8 Ha, final positions / velocities within 1e-8."""
    from dftbplus_b200.api import HotPath
    import scc_harness as sh
    from test_scc_end_to_end import _setup, _oracle_arm_factory, TOL_ENERGY, TOL_FORCE
    w, sk, geo, model, selfegy = _setup("watercluster24")
    ref = sh.velocity_verlet(_oracle_arm_factory(model, selfegy, w["species0"]), w, sk, model.mOrb, n_steps=4)
    hp = HotPath(0)
    try:
        sk.upload(hp)

        def make(x):
            g = hp.updateNeighbourListAndSpecies(x, w["species0"], sk.cutoff(), None, adopt=True)
            return sh.DeviceArm(hp, sk, g), g
        got = sh.velocity_verlet(make, w, sk, model.mOrb, n_steps=4)
    finally:
        hp.close()
    for a, b in zip(got["steps"], ref["steps"]):
        assert a["n_iter"] == b["n_iter"] and a["sparseSize"] == b["sparseSize"], (a, b)
        assert abs(a["energy"] - b["energy"]) < TOL_ENERGY
    assert np.abs(got["x"] - ref["x"]).max() < 1e-8 and np.abs(got["v"] - ref["v"]).max() < 1e-8
    assert np.abs(got["forces"] - ref["forces"]).max() < TOL_FORCE


def test_symmetric_list_on_the_device(hp_for):
    """The reference's `symmetric` option (periodic.F90:580-584): equal to the oracle; such a list is for
    the host's use and cannot become the geometry of the pair kernels."""
    from dftbplus_b200.api import DftbB200Error
    w, sk, geo, model, selfegy = case("fec128")
    hp = hp_for("fec128")
    cutoff = sk.cutoff()
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], _rcell(w, cutoff), cutoff, cutoff,
                                     symmetric=True)
    got = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], symmetric=True)
    assert_same_list(got, ref)
    with pytest.raises(DftbB200Error):
        hp._ck(hp._L.dftbp_b200_adopt_neighbour_list(hp._h))
    # the context takes a normal list again afterwards
    got = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], adopt=True)
    assert got["sparseSize"] == geo["sparseSize"]


def test_nr_of_neighbours_for_other_cutoffs_on_the_device(hp_for):
    """getNrOfNeighboursForAll (periodic.F90:996-1054) on the device list for shorter cutoffs."""
    import ctypes
    w, sk, geo, model, selfegy = case("si512")
    hp = hp_for("si512")
    cutoff = sk.cutoff()
    g = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], adopt=True)
    a = 5.431 / 0.529177249
    L = oracle.lib()
    for c in (0.0, a * np.sqrt(3) / 4, a / np.sqrt(2), 7.3, cutoff):
        ref = np.zeros(g["nAtom"], dtype=np.int32)
        L.orc_nr_of_neighbours(g["nAtom"], g["nNeighbour"].ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                               g["neighDist2"].ctypes.data_as(ctypes.POINTER(ctypes.c_double)), g["mNeigh"],
                               ctypes.c_double(c), ref.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        assert np.array_equal(hp.getNrOfNeighboursForAll(c), ref)

"""The PRODUCT's neighbour-list kernels and launch sequence (dftbplus_b200/csrc/n2_neighbour_kernels.cuh,
n2_neighbour_driver.h) compiled for the host and run thread by thread in shuffled order
(tests/host_eval/host_nl.cpp), against the oracle restatement of updateNeighbourList /
getNrOfNeighboursForAll / getSparseDescriptor -- every array bit for bit, no GPU needed.

The `-m gpu` tests (test_neighbour_list_gpu.py) cover the CUDA backend around the same kernels."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import oracle, case, ROOT
from dftbplus_b200 import workloads

SRC = os.path.join(ROOT, "tests", "host_eval", "host_nl.cpp")
SO = os.path.join(ROOT, "oracle", "_build", "libhostnl.so")
DEPS = [SRC] + [os.path.join(ROOT, "dftbplus_b200", "csrc", f)
                for f in ("n2_neighbour_kernels.cuh", "n2_neighbour_driver.h", "emu_kernel.cuh", "k4_atom_population.cuh")]

ARRAYS = ("coords", "species", "img2CentCell", "iCellVec", "nNeighbour", "iNeighbour", "neighDist2",
          "nNeighbourSK", "iPair", "iAtomStart")


def build_host_so():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, SRC])
    return SO


@pytest.fixture(scope="module")
def host():
    return ctypes.CDLL(build_host_so())


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def host_list(lib, coord0, species0, n_orb_species, cell, rcell, cutoff, sk_cutoff, n_threads=61, seed=1,
              symmetric=False):
    coord0 = np.ascontiguousarray(coord0, dtype=np.float64)
    species0 = np.ascontiguousarray(species0, dtype=np.int32)
    nos = np.ascontiguousarray(n_orb_species, dtype=np.int32)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    rcell = np.ascontiguousarray(rcell, dtype=np.float64)
    n = coord0.shape[0]
    out = np.zeros(8, dtype=np.int32)
    di, dd = ctypes.c_int, ctypes.c_double
    rc = lib.host_nl_update(n, _p(coord0, dd), _p(species0, di), len(nos), _p(nos, di), cell.shape[0], _p(cell, dd),
                            _p(rcell, dd), dd(cutoff), dd(sk_cutoff), n_threads, ctypes.c_uint(seed),
                            1 if symmetric else 0, _p(out, di))
    assert rc == 0
    nAll, mN, sparse, nOrb, e1, e2, valid, launches = (int(v) for v in out)
    if e1:
        raise RuntimeError("Atoms %d and %d too close together" % (e1, e2))
    assert valid
    g = dict(nAtom=n, nAllAtom=nAll, mNeigh=mN, sparseSize=sparse, nOrb=nOrb, launches=launches,
             coords=np.zeros((nAll, 3)), species=np.zeros(nAll, np.int32), img2CentCell=np.zeros(nAll, np.int32),
             iCellVec=np.zeros(nAll, np.int32), nNeighbour=np.zeros(n, np.int32),
             iNeighbour=np.zeros((n, mN + 1), np.int32), neighDist2=np.zeros((n, mN + 1)),
             nNeighbourSK=np.zeros(n, np.int32), iPair=np.zeros((n, mN + 1), np.int32),
             iAtomStart=np.zeros(n + 1, np.int32))
    lib.host_nl_get(_p(g["coords"], dd), _p(g["species"], di), _p(g["img2CentCell"], di), _p(g["iCellVec"], di),
                    _p(g["nNeighbour"], di), _p(g["iNeighbour"], di), _p(g["neighDist2"], dd),
                    _p(g["nNeighbourSK"], di), _p(g["iPair"], di), _p(g["iAtomStart"], di))
    return g


def assert_same_list(got, ref):
    for k in ("nAllAtom", "mNeigh", "sparseSize", "nOrb"):
        assert got[k] == ref[k], k
    for k in ARRAYS:
        assert got[k].shape == ref[k].shape, k
        assert np.array_equal(got[k], ref[k]), k      # bit for bit, doubles included


def cells_of(w, cutoff):
    if w["latvec"] is None:
        return np.zeros((1, 3)), np.zeros((1, 3))
    return oracle.cell_translations(w["latvec"], cutoff)


@pytest.mark.parametrize("name", ["h2o", "ce_o", "scan", "si2", "fec128", "si512", "water1002"])
def test_host_run_of_the_kernels_equals_the_oracle(name, host):
    w, sk, geo, model, selfegy = case(name)
    cutoff = sk.cutoff()
    cell, rcell = cells_of(w, cutoff)
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, cutoff)
    got = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    assert_same_list(got, ref)
    # the trimmed arrays are what the harness' own producer (geometry.py) and the parity tests use
    assert got["sparseSize"] == geo["sparseSize"] and got["nAllAtom"] == geo["nAllAtom"]


def test_perfect_crystal_equal_distance_shells_keep_the_reference_order(host):
    """Unjittered lattices: every shell holds many neighbours at distances equal within tolSameDist2,
    where the order is whatever the reference's tolerance heap sort leaves -- and it fixes iPair."""
    a = 5.42
    basis = np.array([[0, 0, 0], [.5, .5, .5]])
    cc = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(2)])
    coord0 = (cc[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    species0 = np.array([1 + (i % 3 == 0) for i in range(len(coord0))], dtype=np.int32)
    lat = np.diag([3 * a, 3 * a, 2 * a])
    cell, rcell = oracle.cell_translations(lat, 11.0)
    ref = oracle.neighbour_list_full([9, 4], coord0, species0, rcell, 11.0, 11.0)
    assert np.any(np.abs(np.diff(ref["neighDist2"][:, 1:], axis=1)) < 1e-10)
    for seed in (1, 2):
        got = host_list(host, coord0, species0, [9, 4], cell, rcell, 11.0, 11.0, n_threads=17 + 40 * seed, seed=seed)
        assert_same_list(got, ref)


def test_list_cutoff_larger_than_sk_cutoff(host):
    """mCutOff > skCutOff (repulsive / dispersion reach further): rows keep all neighbours of the list
    cutoff, nNeighbourSK and iPair only those of the SK cutoff (main.F90:2562-2578)."""
    w, sk, geo, model, selfegy = case("fec128")
    cutoff, skc = 14.5, sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, skc)
    got = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, skc)
    assert_same_list(got, ref)
    assert np.all(got["nNeighbourSK"] <= got["nNeighbour"]) and got["nNeighbourSK"].sum() < got["nNeighbour"].sum() // 2
    assert got["sparseSize"] == geo["sparseSize"]


def test_sk_cutoff_inside_a_shell_of_equal_distances(host):
    """search_asc_real_geq with tolSameDist2: a cutoff that cuts a shell of a perfect lattice within
    the tolerance, once from below and once from above."""
    a = 4.0
    coord0 = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], dtype=np.float64) * a
    species0 = np.ones(27, dtype=np.int32)
    cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    for skc in (np.sqrt(32.0 - 0.5e-10), np.sqrt(32.0 + 0.5e-10), np.sqrt(32.0 - 3e-10), 4.0, 0.5):
        ref = oracle.neighbour_list_full([4], coord0, species0, rcell, 9.0, skc)
        got = host_list(host, coord0, species0, [4], cell, rcell, 9.0, skc)
        assert_same_list(got, ref)


@pytest.mark.parametrize("shape", ["planar", "line", "single", "wide"])
def test_degenerate_bounding_boxes(shape, host):
    rng = np.random.default_rng(5)
    if shape == "planar":
        coord0 = np.c_[rng.uniform(0, 40, (60, 2)), np.full(60, 1.5)]
    elif shape == "line":
        coord0 = np.c_[np.zeros(40), np.zeros(40), np.arange(40) * 1.7]
    elif shape == "single":
        coord0 = np.array([[0.3, -0.2, 0.9]])
    else:  # more than kMaxBinsPerDim cutoffs wide: the bin edge grows instead of the bin count
        coord0 = np.c_[np.arange(120) * 3.9, rng.uniform(0, 3, 120), rng.uniform(0, 3, 120)]
    species0 = np.ones(len(coord0), dtype=np.int32)
    cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    ref = oracle.neighbour_list_full([4], coord0, species0, rcell, 11.0, 11.0)
    got = host_list(host, coord0, species0, [4], cell, rcell, 11.0, 11.0)
    assert_same_list(got, ref)
    if shape == "single":
        assert got["mNeigh"] == 0 and got["nAllAtom"] == 1 and got["sparseSize"] == 16


def test_small_cell_periodic_images_of_the_atom_itself(host):
    """One atom in a small cell: all neighbours are its own images (fac = 1 pairs of the stress)."""
    lat = np.array([[4.1, 0.0, 0.0], [1.0, 4.3, 0.0], [0.2, 0.7, 5.0]])
    coord0 = np.array([[0.4, 0.1, 0.2]])
    cell, rcell = oracle.cell_translations(lat, 9.0)
    ref = oracle.neighbour_list_full([9], coord0, [1], rcell, 9.0, 9.0)
    got = host_list(host, coord0, [1], [9], cell, rcell, 9.0, 9.0)
    assert_same_list(got, ref)
    assert np.all(got["img2CentCell"] == 1) and got["mNeigh"] > 20


def test_results_do_not_depend_on_the_thread_order_or_count(host):
    w, sk, geo, model, selfegy = case("fec128")
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    first = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff, 1, 0)
    for n_threads, seed in ((7, 3), (256, 4), (5000, 5)):
        again = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff,
                          n_threads, seed)
        assert_same_list(again, first)


def test_too_close_atoms_report_the_reference_pair(host):
    """pairError = first (cell, iAtom1, iAtom2) in the reference's loop order (periodic.F90:616-628)."""
    w, sk, geo, model, selfegy = case("fec128")
    coord0 = w["coord0"].copy()
    coord0[77] = coord0[12] + [0.003, 0.0, -0.002]
    coord0[90] = coord0[31] + [0.0, 0.004, 0.0]
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    with pytest.raises(RuntimeError) as ref:
        oracle.neighbour_list_full(sk.n_orb_species, coord0, w["species0"], rcell, cutoff, cutoff)
    with pytest.raises(RuntimeError) as got:
        host_list(host, coord0, w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    assert str(got.value) == str(ref.value) == "Atoms 78 and 13 too close together"
    # an atom on top of its own periodic image is an error as well (only cell 1 is exempt)
    lat = np.diag([0.008, 30.0, 30.0])
    cell, rcell = oracle.cell_translations(lat, 2.0)
    with pytest.raises(RuntimeError) as ref:
        oracle.neighbour_list_full([4], np.zeros((1, 3)), [1], rcell, 2.0, 2.0)
    with pytest.raises(RuntimeError) as got:
        host_list(host, np.zeros((1, 3)), [1], [4], cell, rcell, 2.0, 2.0)
    assert str(got.value) == str(ref.value)


def test_grid_never_has_bins_narrower_than_the_cutoff(host):
    w, sk, geo, model, selfegy = case("si512")
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    nb, lo, inv = np.zeros(3, np.int32), np.zeros(3), np.zeros(3)
    host.host_nl_grid(_p(nb, ctypes.c_int), _p(lo, ctypes.c_double), _p(inv, ctypes.c_double))
    assert np.all(1.0 / inv >= cutoff * 1.0001 * (1 - 1e-15)) and np.all(nb >= 1) and np.all(nb <= 32)
    ext = w["coord0"].max(axis=0) - w["coord0"].min(axis=0)
    assert np.all(nb / inv >= ext * (1 - 1e-15)) and np.allclose(lo, w["coord0"].min(axis=0), rtol=0, atol=0)


def test_cell_translations_through_the_c_abi_equal_the_oracle():
    """dftbp_b200_cell_translations is host code of the shipped library (no device needed)."""
    from dftbplus_b200 import _lib, build as build_mod
    if not os.path.exists(_lib.SO_PATH):
        build_mod.build()
    L = _lib.load()
    dd = ctypes.c_double
    for lat, cutoff in ((np.array([[10.0, 0, 0], [2.0, 9.0, 0], [0, 1.0, 12.0]]), 11.0),
                        (0.5 * 10.263 * np.array([[0.0, 1, 1], [1, 0, 1], [1, 1, 0]]), 11.0),
                        (np.diag([86.8, 86.8, 43.4]), 11.0), (np.diag([3.0, 50.0, 7.0]), 0.0)):
        lat = np.ascontiguousarray(lat)
        rec = np.ascontiguousarray(np.linalg.inv(lat).T)
        n = ctypes.c_int()
        assert L.dftbp_b200_cell_translations(_p(lat, dd), _p(rec, dd), cutoff, 0, ctypes.byref(n), None, None) == 0
        cell, rcell = np.zeros((n.value, 3)), np.zeros((n.value, 3))
        assert L.dftbp_b200_cell_translations(_p(lat, dd), _p(rec, dd), cutoff, n.value, ctypes.byref(n),
                                              _p(cell, dd), _p(rcell, dd)) == 0
        c_ref, r_ref = oracle.cell_translations(lat, cutoff)
        assert np.array_equal(cell, c_ref) and np.array_equal(rcell, r_ref)
        # too small an output array is refused, not overrun
        assert L.dftbp_b200_cell_translations(_p(lat, dd), _p(rec, dd), cutoff, n.value - 1, ctypes.byref(n),
                                              _p(cell, dd), _p(rcell, dd)) == 1


def test_fold_coords_to_cell_equals_the_oracle(host):
    """foldCoordsToCell (boundarycond.F90:369-380): the step before the list in handleCoordinateChange."""
    rng = np.random.default_rng(21)
    dd = ctypes.c_double
    lats = [np.diag([86.8, 86.8, 43.4]), np.array([[10.0, 0, 0], [2.0, 9.0, 0], [0, 1.0, 12.0]]),
            0.5 * 10.263 * np.array([[0.0, 1, 1], [1, 0, 1], [1, 1, 0]])]
    for lat in lats:
        lat = np.ascontiguousarray(lat)
        frac = rng.uniform(-3.0, 3.0, (500, 3))
        frac[:20] = rng.integers(-2, 3, (20, 3))                   # lattice points: the frac = 1 - eps guard
        frac[20:40] = np.round(frac[20:40] * 4) / 4
        coord = frac @ lat
        ref = oracle.fold_coords_to_cell(coord, lat)
        got = np.array(coord, copy=True)
        assert host.host_nl_fold(len(got), _p(got, dd), _p(lat, dd), 37, 3) == 0
        assert np.array_equal(got, ref)
        back = np.linalg.solve(lat.T, ref.T).T
        assert back.min() > -1e-12 and back.max() < 1 + 1e-12
        shift = np.linalg.solve(lat.T, (coord - ref).T).T
        assert np.abs(shift - np.round(shift)).max() < 1e-9        # moved by lattice vectors only
    sing = np.ascontiguousarray(np.array([[1.0, 0, 0], [2.0, 0, 0], [0, 0, 1.0]]))
    c = np.zeros((2, 3))
    assert host.host_nl_fold(2, _p(c, dd), _p(sing, dd), 4, 1) == 1


def test_sm100a_code_of_the_distance_kernels_has_no_fused_multiply_add():
    """Bit-exact distances need the reference's separate multiplies and adds: the SASS of the built
    library must not contain DFMA in the kernels that produce coordinates / squared distances."""
    import shutil
    from dftbplus_b200 import build as build_mod
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    build_mod.build()
    obj = os.path.join(ROOT, "dftbplus_b200", "lib", "n2_neighbour_list.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fn, dfma, seen = None, {}, set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            seen.add(fn)
        elif "DFMA" in line and fn:
            dfma[fn] = dfma.get(fn, 0) + 1
    for kernel in ("nl_count", "nl_fill", "nl_rank", "nl_sort", "nl_fold"):
        names = [f for f in seen if kernel in f]
        assert names, kernel
        assert all(dfma.get(f, 0) == 0 for f in names), (kernel, dfma)


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128"])
def test_per_atom_population_kernels_equal_the_oracle(name, host):
    """mullikenPerAtom's sum over orbitals and getOnsitePopulation, host run of
    k4_atom_population.cuh: bit for bit."""
    w, sk, geo, model, selfegy = case(name)
    rng = np.random.default_rng(17)
    dd, di = ctypes.c_double, ctypes.c_int
    over = oracle.build_s(model, geo)
    rho = over * rng.normal(size=over.shape)
    q_orb = np.ascontiguousarray(oracle.mulliken_per_orbital(model, geo, over, rho))
    q0 = rng.normal(size=geo["nAtom"])
    qq = q0.copy()
    host.host_pop_atom_sum(geo["nAtom"], q_orb.shape[1], _p(q_orb, dd), _p(qq, dd), 13, 2)
    assert np.array_equal(qq, oracle.mulliken_per_atom(model, geo, over, rho, qq=q0))
    assert abs(oracle.mulliken_per_atom(model, geo, over, rho).sum() - q_orb.sum()) < 1e-10
    ipair = np.ascontiguousarray(geo["iPair"], dtype=np.int32)
    start = np.ascontiguousarray(geo["iAtomStart"], dtype=np.int32)
    out = np.full(geo["nAtom"], np.nan)
    host.host_pop_onsite(geo["nAtom"], ipair.shape[1], _p(ipair, di), _p(start, di), _p(rho, dd), _p(out, dd), 29, 3)
    ref = oracle.onsite_population(geo, rho)
    assert np.array_equal(out, ref)
    # independent statement: dense diagonal of the unpacked matrix, atom by atom (gamma point, no self images)
    if name in ("h2o", "ce_o"):
        dense = oracle.unpack_hs_real(geo, rho)
        d = np.diag(dense)
        assert np.allclose(ref, [d[start[i] - 1:start[i + 1] - 1].sum() for i in range(geo["nAtom"])], atol=1e-13)


def test_zero_cutoff_and_pairs_exactly_at_the_cutoff(host):
    """dist2 > cutoff2 skips, so dist2 == cutoff2 is a neighbour (periodic.F90:598); cutoff = 0 leaves
    every atom with itself only."""
    cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    chain = np.c_[np.arange(40) * 8.0, np.zeros(40), np.zeros(40)]           # spaced exactly one cutoff apart
    species0 = np.ones(40, dtype=np.int32)
    ref = oracle.neighbour_list_full([4], chain, species0, rcell, 8.0, 8.0)
    got = host_list(host, chain, species0, [4], cell, rcell, 8.0, 8.0)
    assert_same_list(got, ref)
    assert got["nNeighbour"].sum() == 39 and np.all(got["neighDist2"][:-1, 1] == 64.0)
    ref = oracle.neighbour_list_full([4], chain, species0, rcell, 0.0, 0.0)
    got = host_list(host, chain, species0, [4], cell, rcell, 0.0, 0.0)
    assert_same_list(got, ref)
    assert got["mNeigh"] == 0 and got["nAllAtom"] == 40 and got["sparseSize"] == 40 * 16
    # periodic, images exactly at the cutoff: cubic cell of edge = cutoff, one atom
    lat = np.eye(3) * 6.0
    cell, rcell = oracle.cell_translations(lat, 6.0)
    ref = oracle.neighbour_list_full([1], np.zeros((1, 3)), [1], rcell, 6.0, 6.0)
    got = host_list(host, np.zeros((1, 3)), [1], [1], cell, rcell, 6.0, 6.0)
    assert_same_list(got, ref)
    assert got["nNeighbour"][0] == 6            # the six face neighbours at exactly 6.0


def test_reference_known_answers_of_the_bisection(host):
    """test/src/dftbp/unit/math/binarysearch.F90, "binarysearch_real_simple" (search_asc_real_geq part):
    the reference's own data, probes and expected results, for the oracle and the product's routine."""
    L = oracle.lib()
    dd = ctypes.c_double
    tol = float(np.finfo(np.float32).eps)                      # real(epsilon(0.0_rsp), dp)
    data = np.array([0.1, 0.2, 0.3, 0.4, 0.4, 0.5, 0.6, 0.8])
    x = [-1.0, 0.1 + tol / 2.0, 10.0, 0.8 - tol / 2.0, 0.4]
    expect_geq = [0, 1, 8, 8, 5]
    host.host_search_asc_real_geq.restype = ctypes.c_int
    L.orc_search_asc_real_geq.restype = ctypes.c_int
    for xv, e in zip(x, expect_geq):
        assert L.orc_search_asc_real_geq(len(data), _p(data, dd), dd(xv), dd(tol)) == e
        assert host.host_search_asc_real_geq(len(data), _p(data, dd), dd(xv), dd(tol)) == e
    for n in (0, 1):                                           # empty and one-element arrays
        for xv in (0.05, 0.1, 0.2):
            assert (host.host_search_asc_real_geq(n, _p(data, dd), dd(xv), dd(tol))
                    == L.orc_search_asc_real_geq(n, _p(data, dd), dd(xv), dd(tol)) == (1 if n == 1 and xv >= 0.1 else 0))


# data of the reference's sorting unit test (test/src/dftbp/unit/math/sorting.F90, from C60_OscWindow):
# clusters of values equal within tolerance
SORT_DATA = [
    1.0009256850962192E-1, 1.0009256850980475E-1, 1.0009256850996931E-1, 1.0009256824902155E-1,
    1.0009256824920437E-1, 1.0009256824936894E-1, 1.0009256824896715E-1, 1.0009256824914997E-1,
    1.0009256824931453E-1, 1.0009256824891125E-1, 1.0009256824909407E-1, 1.0009256824925863E-1,
    6.6042780967860620E-2, 6.6042780968043446E-2, 6.6042780968208009E-2, 9.4447988867716701E-2,
    9.4447988867874991E-2, 9.4447988868041677E-2, 6.6042780967824899E-2, 6.6042780968007725E-2,
    6.6042780968172288E-2, 9.4447988867680979E-2, 9.4447988867839269E-2, 9.4447988868005955E-2,
    6.6042780967797060E-2, 6.6042780967979886E-2, 6.6042780968144449E-2, 9.4447988867653140E-2,
    9.4447988867811430E-2, 9.4447988867978117E-2, 6.6042780825941200E-2, 6.6042780826124026E-2,
    6.6042780826288588E-2, 9.4447988725797280E-2, 9.4447988725955570E-2, 9.4447988726122256E-2,
    6.6042780825940506E-2, 6.6042780826123332E-2, 6.6042780826287895E-2, 9.4447988725796586E-2,
    9.4447988725954876E-2, 9.4447988726121562E-2]


def test_record_heap_sort_equals_the_index_heap_sort_on_the_reference_data(host):
    """The product sorts records where the reference sorts an index vector (math/sorting.F90:142-199):
    same permutation on the reference's own test data (near-equal clusters), for several tolerances
    and for random data with ties; the result is ascending within the tolerance."""
    L = oracle.lib()
    dd, di = ctypes.c_double, ctypes.c_int
    rng = np.random.default_rng(31)
    sets = [np.array(SORT_DATA), np.round(rng.uniform(0, 1, 200), 2), rng.uniform(0, 1, 77), np.zeros(9),
            np.array([3.0]), np.array([2.0, 1.0])]
    for data in sets:
        for tol in (float(np.finfo(np.float32).eps), 1e-10, 0.0, 1e-3):
            n = len(data)
            indx = np.zeros(n, dtype=np.int32)
            L.orc_index_heap_sort(n, _p(indx, di), _p(np.ascontiguousarray(data), dd), dd(tol))
            keys = np.arange(1, n + 1, dtype=np.int32)
            vals = np.array(data, copy=True)
            host.host_heap_sort_records(n, _p(keys, di), _p(vals, dd), dd(tol))
            assert np.array_equal(keys, indx)
            assert np.array_equal(vals, data[indx - 1])
            assert sorted(keys) == list(range(1, n + 1))
            assert np.all(vals[:-1] - vals[1:] <= tol)         # never descending by more than the tolerance


def test_randomized_geometries_against_the_oracle(host):
    """Seeded sweep: triclinic and very flat cells, tiny cells with hundreds of images, sparse clusters,
    atoms on exact multiples of the bin edge, several list / SK cutoffs and species mixes."""
    rng = np.random.default_rng(2025)
    n_case = 0
    for trial in range(100):
        n_atom = int(rng.integers(1, 40))
        n_orb_species = [int(v) for v in rng.choice([1, 4, 9, 16], size=int(rng.integers(1, 4)))]
        species0 = rng.integers(1, len(n_orb_species) + 1, n_atom).astype(np.int32)
        cutoff = float(rng.choice([2.5, 4.0, 6.3, 9.0, 11.0]))
        skc = cutoff if trial % 3 else float(cutoff * rng.uniform(0.3, 1.0))
        kind = trial % 4
        if kind == 0:      # cluster, sparse and wide
            coord0 = rng.uniform(-1, 1, (n_atom, 3)) * rng.choice([3.0, 30.0, 300.0])
            lat = None
        elif kind == 1:    # triclinic cell
            lat = np.diag(rng.uniform(3.0, 14.0, 3)) + rng.uniform(-2.0, 2.0, (3, 3))
            coord0 = rng.uniform(0, 1, (n_atom, 3)) @ lat
        elif kind == 2:    # flat / needle cells
            lat = np.diag(rng.permutation([rng.uniform(2.5, 4.0), rng.uniform(8.0, 12.0), rng.uniform(25.0, 40.0)]))
            coord0 = rng.uniform(0, 1, (n_atom, 3)) @ lat
        else:              # atoms on exact multiples of 1.0001 * cutoff (bin edges), non-periodic
            coord0 = rng.integers(0, 6, (n_atom, 3)) * (1.0001 * cutoff) + rng.choice([0.0, 1e-13, -1e-13], (n_atom, 3))
            lat = None
        if trial % 5 == 4:  # simple cubic lattice, periodic: exact ties in every shell
            a, m = float(rng.uniform(2.0, 5.0)), int(rng.integers(1, 4))
            coord0 = np.array([[i, j, k] for i in range(m) for j in range(m) for k in range(m)], dtype=float) * a
            n_atom, lat = len(coord0), np.eye(3) * a * m
            species0 = rng.integers(1, len(n_orb_species) + 1, n_atom).astype(np.int32)
        symmetric = trial % 7 == 0
        if lat is None:
            cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
        else:
            cell, rcell = oracle.cell_translations(lat, cutoff)
        if cell.shape[0] * n_atom > 300000:
            continue
        try:
            ref = oracle.neighbour_list_full(n_orb_species, coord0, species0, rcell, cutoff, skc, symmetric=symmetric)
        except RuntimeError as e:
            with pytest.raises(RuntimeError) as got:
                host_list(host, coord0, species0, n_orb_species, cell, rcell, cutoff, skc, 23, trial, symmetric=symmetric)
            assert str(got.value) == str(e)
            continue
        got = host_list(host, coord0, species0, n_orb_species, cell, rcell, cutoff, skc, n_threads=23 + trial, seed=trial,
                        symmetric=symmetric)
        assert_same_list(got, ref)
        n_case += 1
    assert n_case > 60


def test_randomized_many_bin_grids_against_the_oracle(host):
    """Larger random systems whose grid has many bins per dimension (up to the 32 cap and beyond)."""
    rng = np.random.default_rng(77)
    for trial in range(8):
        n_atom = int(rng.integers(150, 500))
        cutoff = float(rng.choice([3.0, 5.0, 7.5]))
        box = rng.uniform(8.0, 45.0, 3) * cutoff / 3.0 * np.array([1.0, 1.0, rng.choice([0.3, 1.0, 4.0])])
        coord0 = rng.uniform(0, 1, (n_atom, 3)) * box
        # keep atoms apart (no proximity error): drop the later atom of every pair closer than 0.3
        keep = np.ones(n_atom, dtype=bool)
        for i in range(n_atom):
            if keep[i]:
                d = np.linalg.norm(coord0[i + 1:] - coord0[i], axis=1)
                keep[i + 1:] &= d > 0.3
        coord0 = coord0[keep]
        species0 = rng.integers(1, 3, len(coord0)).astype(np.int32)
        periodic = trial % 2 == 0
        lat = np.diag(box + 0.4) if periodic else None
        cell, rcell = oracle.cell_translations(lat, cutoff) if periodic else (np.zeros((1, 3)), np.zeros((1, 3)))
        ref = oracle.neighbour_list_full([4, 9], coord0, species0, rcell, cutoff, cutoff)
        got = host_list(host, coord0, species0, [4, 9], cell, rcell, cutoff, cutoff, n_threads=97, seed=trial)
        assert_same_list(got, ref)
        nb = np.zeros(3, np.int32)
        lo, inv = np.zeros(3), np.zeros(3)
        host.host_nl_grid(_p(nb, ctypes.c_int), _p(lo, ctypes.c_double), _p(inv, ctypes.c_double))
        assert nb.max() >= 3


def test_reference_neighbourdistance_unit_test_on_the_product_kernels(host):
    """test/src/dftbp/unit/dftb/periodic.F90 "neighbourdistance": two atoms at minNeighDist = 1e-2 bohr
    are accepted, at minNeighDist**2 the list must fail with "too close" -- for the product's kernels."""
    cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    min_neigh_dist = 1.0e-2                                     # common/accuracy.F90:69
    ok = host_list(host, np.array([[0.0, 0, 0], [min_neigh_dist, 0, 0]]), [1, 1], [1], cell, rcell, 1.0, 1.0)
    assert ok["nNeighbour"].tolist() == [1, 0] and ok["nAllAtom"] == 2 and ok["sparseSize"] == 3
    assert ok["neighDist2"][0, 1] == min_neigh_dist ** 2
    with pytest.raises(RuntimeError, match="Atoms 2 and 1 too close together"):
        host_list(host, np.array([[0.0, 0, 0], [min_neigh_dist ** 2, 0, 0]]), [1, 1], [1], cell, rcell, 1.0, 1.0)


@pytest.mark.parametrize("name", ["ce_o", "si2", "fec128"])
def test_symmetric_list_equals_the_oracle(name, host):
    """The reference's `symmetric` option (periodic.F90:580-584; symNeighbourList of the hybrid
    functionals): iAtom2 over all atoms -- every pair appears in both rows, more images are numbered."""
    w, sk, geo, model, selfegy = case(name)
    cutoff = sk.cutoff()
    cell, rcell = cells_of(w, cutoff)
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, cutoff, symmetric=True)
    got = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff, symmetric=True)
    assert_same_list(got, ref)
    half = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, cutoff)
    n_self = int(sum((half["img2CentCell"][half["iNeighbour"][i, 1:half["nNeighbour"][i] + 1] - 1] == i + 1).sum()
                     for i in range(half["nAtom"])))
    # every off-site pair twice; images of the atom itself appear once per direction in the half list already
    assert got["nNeighbour"].sum() == 2 * (half["nNeighbour"].sum() - n_self) + n_self
    assert got["nAllAtom"] >= half["nAllAtom"]


def test_both_ordering_paths_are_exercised(host):
    """Rows without near-ties are placed by rank (nl_rank), rows with near-ties go through the reference's
    heap sort (nl_sort): jittered systems take the first path only, perfect lattices the second."""
    out = np.zeros(2, dtype=np.int32)
    w, sk, geo, model, selfegy = case("fec128")
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    host.host_nl_tie_rows(_p(out, ctypes.c_int))
    assert out[0] == 0 and out[1] == 128
    w, sk, geo, model, selfegy = case("si512")                 # unjittered diamond: every shell is a tie
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    host.host_nl_tie_rows(_p(out, ctypes.c_int))
    assert out[0] > 400
    # one tied pair in an otherwise generic cluster: only the rows that see both atoms at equal distance
    rng = np.random.default_rng(3)
    coord0 = rng.uniform(0, 12, (30, 3))
    coord0[7] = coord0[0] + [2.0, 0.0, 0.0]
    coord0[9] = coord0[0] + [0.0, 2.0, 0.0]                     # atoms 8 and 10 equidistant from atom 1
    species0 = np.ones(30, dtype=np.int32)
    cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    ref = oracle.neighbour_list_full([4], coord0, species0, rcell, 9.0, 9.0)
    got = host_list(host, coord0, species0, [4], cell, rcell, 9.0, 9.0)
    assert_same_list(got, ref)
    host.host_nl_tie_rows(_p(out, ctypes.c_int))
    assert 1 <= out[0] <= 3


def test_long_rows_skip_the_counting(host):
    """Rows longer than rankMax are ordered by the two heap sorts (n log n) instead of by counting (n^2):
    same result whatever the threshold."""
    w, sk, geo, model, selfegy = case("fec128")
    cutoff = 14.5
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, sk.cutoff())
    out = np.zeros(2, dtype=np.int32)
    try:
        for rank_max, some_long in ((1, True), (40, True), (100000, False)):
            host.host_nl_set_rank_max(rank_max)
            got = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, sk.cutoff())
            assert_same_list(got, ref)
            host.host_nl_tie_rows(_p(out, ctypes.c_int))
            assert (out[0] > 0) == some_long
        # a perfect lattice with a low threshold: long rows and tie rows together
        w2, sk2, geo2, model2, selfegy2 = case("si512")
        c2, r2 = oracle.cell_translations(w2["latvec"], sk2.cutoff())
        ref2 = oracle.neighbour_list_full(sk2.n_orb_species, w2["coord0"], w2["species0"], r2, sk2.cutoff(), sk2.cutoff())
        host.host_nl_set_rank_max(20)
        assert_same_list(host_list(host, w2["coord0"], w2["species0"], sk2.n_orb_species, c2, r2, sk2.cutoff(),
                                   sk2.cutoff()), ref2)
    finally:
        host.host_nl_set_rank_max(0)


def test_nr_of_neighbours_for_other_cutoffs(host):
    """getNrOfNeighboursForAll on the finished list for shorter cutoffs (repulsive / CAM ranges)."""
    w, sk, geo, model, selfegy = case("si512")                  # perfect lattice: cutoffs inside tie shells
    cutoff = sk.cutoff()
    cell, rcell = oracle.cell_translations(w["latvec"], cutoff)
    got = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    a = 5.431 / 0.529177249
    L = oracle.lib()
    for c in (0.0, a * np.sqrt(3) / 4, a / np.sqrt(2), a / np.sqrt(2) - 1e-6, 7.3, cutoff):
        ref = np.zeros(got["nAtom"], dtype=np.int32)
        L.orc_nr_of_neighbours(got["nAtom"], _p(got["nNeighbour"], ctypes.c_int), _p(got["neighDist2"], ctypes.c_double),
                               got["mNeigh"], ctypes.c_double(c), _p(ref, ctypes.c_int))
        out = np.full(got["nAtom"], -1, dtype=np.int32)
        host.host_nl_nr_of_neighbours(ctypes.c_double(c), _p(out, ctypes.c_int), 19, 5)
        assert np.array_equal(out, ref)
    assert np.array_equal(out, got["nNeighbour"])               # the list cutoff itself: everything


def test_kernels_are_clean_under_address_sanitizer(host):
    """The same kernels and launch sequence under ASan / UBSan: the CPU backend allocates every buffer
    with exactly the size the GPU backend would, so an out-of-range index (an illegal address on the
    device) fails here."""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = os.path.join(ROOT, "oracle", "_build", "libhostnl_asan.so")
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in DEPS):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", so, SRC])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_eval", "asan_run.py"), so], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "asan/ubsan run finished" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-4000:]


def _host_descriptors(g, group=32):
    """set_geometry's host-side work list / BlockDesc records / element -> descriptor index (context.cu,
    flat_stream.cuh buildGroupIndex) restated with plain loops."""
    n, ld = g["nAtom"], g["mNeigh"] + 1
    work, desc, dup = [], [], False
    start, img, vec = g["iAtomStart"], g["img2CentCell"], g["iCellVec"]
    for i in range(n):
        seen = {i}
        for k in range(g["nNeighbourSK"][i] + 1):
            a2 = g["iNeighbour"][i, k] - 1
            f = img[a2] - 1
            work.append(k + ld * i)
            desc.append((g["iPair"][i, k], start[i] - 1, start[f] - 1,
                         (start[i + 1] - start[i]) | ((start[f + 1] - start[f]) << 8), i, f, vec[a2] - 1, k))
            if k >= 1:
                dup = dup or f in seen
                seen.add(f)
    desc = np.array(desc, dtype=np.int32)
    n_groups = (g["sparseSize"] + group - 1) // group
    first = np.searchsorted(desc[:, 0], group * np.arange(n_groups), side="right") - 1
    return np.array(work, np.int32), desc, np.maximum(first, 0).astype(np.int32), dup


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128", "si512", "water1002"])
def test_device_built_block_descriptors_equal_the_host_built_ones(name, host):
    """adopt_neighbour_list builds the work list, the BlockDesc records, the span-pipeline index and the
    hasDuplicates flag with kernels (geo_block_count / geo_describe / geo_group_index); here the same
    kernels run on the host against a restatement of what set_geometry builds from host arrays."""
    w, sk, geo, model, selfegy = case(name)
    cutoff = sk.cutoff()
    cell, rcell = cells_of(w, cutoff)
    di = ctypes.c_int
    for n_threads, seed in ((61, 1), (1, 2), (4099, 3)):
        g = host_list(host, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff,
                      n_threads=n_threads, seed=seed)
        h_start = np.zeros(g["nAtom"] + 1, np.int32)
        n_blocks = host.host_geo_count(_p(h_start, di), n_threads, ctypes.c_uint(seed))
        assert n_blocks == int(g["nNeighbourSK"].sum()) + g["nAtom"]
        assert np.array_equal(h_start, np.concatenate([[0], np.cumsum(g["nNeighbourSK"] + 1)]))
        work = np.zeros(n_blocks, np.int32)
        desc = np.zeros((n_blocks, 8), np.int32)
        n_groups = (g["sparseSize"] + 31) // 32
        first = np.full(n_groups, -7, np.int32)
        dup = host.host_geo_describe(n_blocks, 32, _p(work, di), _p(desc, di), _p(first, di), n_threads,
                                     ctypes.c_uint(seed))
        r_work, r_desc, r_first, r_dup = _host_descriptors(g)
        assert np.array_equal(work, r_work)
        assert np.array_equal(desc, r_desc)
        assert np.array_equal(first, r_first)
        assert dup == int(r_dup)
    # small cells fold several images onto one dense block; clusters and large boxes do not
    if name in ("h2o", "ce_o", "water1002"):
        assert not dup
    if name in ("si2", "fec128"):
        assert dup

"""Parity of the sm_100a kernels against the oracle, through the C ABI (via dftbplus_b200.api).

Tolerances: sparse H0 / S / rho 1e-12 relative in the max norm with a
1e-14 elementwise floor; forces 1e-7 Ha/bohr; integer/index-only data movement bit exact."""
import numpy as np
import pytest
import scipy.linalg

from helpers import (oracle, case, fermi_fillings, n_electrons, assert_sparse_close, TOL_FORCE)

pytestmark = pytest.mark.gpu

CASES = ["h2o", "ce_o", "si2", "fec128", "si512", "water1002"]


@pytest.fixture(scope="module")
def hp_cache():
    from dftbplus_b200.api import HotPath
    cache = {}

    def get(name, rank=0, nranks=1, method=2):
        key = (name, rank, nranks, method)
        if key not in cache:
            w, sk, geo, model, selfegy = case(name, method)
            hp = HotPath(0, rank, nranks)
            sk.upload(hp)
            hp.set_geometry(geo)
            cache[key] = hp
        return cache[key]
    yield get
    for hp in cache.values():
        hp.close()


@pytest.mark.parametrize("name", CASES)
def test_k1_build_h0_s(name, hp_cache):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    h_ref, s_ref = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    h, s = hp.buildH0S(sk.selfegy)
    assert_sparse_close(h, h_ref, what="H0 " + name)
    assert_sparse_close(s, s_ref, what="S " + name)
    assert np.array_equal(hp.buildH0(sk.selfegy), h) and np.array_equal(hp.buildS(), s)   # fused == separate
    assert hp.num_pairs == int(geo["nNeighbourSK"].sum())


def test_k1_tail_region_and_cutoff(hp_cache):
    """Pairs beyond the last grid point use the 5th-order tail; beyond rMax the blocks are zero."""
    w, sk, geo, model, selfegy = case("fec128")
    hp = hp_cache("fec128")
    s, s_ref = hp.buildS(), oracle.build_s(model, geo)
    r = np.sqrt(geo["neighDist2"])
    tail = [(i, k) for i in range(geo["nAtom"]) for k in range(1, geo["nNeighbourSK"][i] + 1) if r[i, k] > 10.0]
    assert len(tail) > 10
    for i, k in tail[:200]:
        n = geo["nOrbAtom"][i] * geo["nOrbAtom"][geo["img2CentCell"][geo["iNeighbour"][i, k] - 1] - 1]
        blk, ref = s[geo["iPair"][i, k]:geo["iPair"][i, k] + n], s_ref[geo["iPair"][i, k]:geo["iPair"][i, k] + n]
        assert np.abs(blk - ref).max() <= 1e-12 * np.abs(s_ref).max()
        assert np.allclose(blk, ref, rtol=1e-9, atol=1e-16)        # small tail values keep relative accuracy


@pytest.mark.parametrize("name", ["scan", "h2o", "ce_o", "fec128"])
@pytest.mark.parametrize("method", [1, 2])
def test_k1_interpolation_methods(name, method, hp_cache):
    """OldSKInterpolation = Yes (skEqGridOld, slakoeqgrid.F90:248-317) and the default method on a
    distance scan that visits every branch: polynomial, cubic spline, tail, beyond the cutoff."""
    w, sk, geo, model, selfegy = case(name, method)
    hp = hp_cache(name, method=method)
    assert abs(hp.sk_cutoff() - model.cutoff()) < 1e-12 and abs(sk.cutoff() - model.cutoff()) < 1e-12
    h, s = hp.buildH0S(sk.selfegy)
    h_ref, s_ref = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    assert_sparse_close(h, h_ref, what="H0 %s method %d" % (name, method))
    assert_sparse_close(s, s_ref, what="S %s method %d" % (name, method))
    if name == "scan":
        # block by block, relative to the block's own size: the tail values are tiny
        r = np.sqrt(geo["neighDist2"])
        seen = set()
        for k in range(1, geo["nNeighbourSK"][0] + 1):
            j = geo["iNeighbour"][0, k] - 1
            n = geo["nOrbAtom"][0] * geo["nOrbAtom"][geo["img2CentCell"][j] - 1]
            a = geo["iPair"][0, k]
            for got, ref in ((h[a:a + n], h_ref[a:a + n]), (s[a:a + n], s_ref[a:a + n])):
                assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max() + 1e-18, (method, r[0, k])
            seen.add(bool(np.abs(s_ref[a:a + n]).max() > 0))
        assert seen == {True} or seen == {True, False}
        assert geo["nNeighbourSK"][0] >= (14 if method == 1 else 18)


def test_k1_partial_ranks_sum_to_full(hp_cache):
    w, sk, geo, model, selfegy = case("fec128")
    full = hp_cache("fec128").buildS()
    parts = [hp_cache("fec128", r, 3).buildS() for r in range(3)]
    assert np.array_equal(sum(parts), full)
    for r in range(3):
        ref = oracle.build_s(model, geo, rank=r, nranks=3)
        assert np.array_equal(parts[r] != 0, ref != 0)             # same atom chunk (getChunkRanges)


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128", "si512"])
def test_k2_unpack_pack_real(name, hp_cache):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    x = np.random.default_rng(5).normal(size=geo["sparseSize"])
    sq, sq_ref = hp.unpackHS(x), oracle.unpack_hs_real(geo, x)
    if name in ("si2", "fec128"):
        assert np.allclose(sq, sq_ref, rtol=0, atol=1e-13)           # colliding images: sums reorder
    else:
        assert np.array_equal(sq, sq_ref)                            # pure data movement
    d = np.random.default_rng(6).normal(size=(geo["nOrb"], geo["nOrb"]))
    d = np.asfortranarray(d)
    acc0 = np.random.default_rng(7).normal(size=geo["sparseSize"])
    got = hp.packHS(acc0.copy(), d)
    ref = oracle.pack_hs_real(geo, acc0.copy(), d)
    assert np.array_equal(got, ref)                                   # gather + one add: bit exact


@pytest.mark.parametrize("name", ["si2", "si512", "fec128"])
def test_k2_unpack_pack_kpoints(name, hp_cache):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    x = np.random.default_rng(8).normal(size=geo["sparseSize"])
    for kp in ([0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.125, -0.375, 0.25]):
        sq, sq_ref = hp.unpackHS(x, kpoint=kp), oracle.unpack_hs_cplx(geo, x, kp)
        assert np.abs(sq - sq_ref).max() <= 1e-12 * np.abs(sq_ref).max()
        n = geo["nOrb"]
        rng = np.random.default_rng(9)
        d = np.asfortranarray(rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
        acc0 = rng.normal(size=geo["sparseSize"])
        got = hp.packHS(acc0.copy(), d, kpoint=kp, kweight=0.375)
        ref = oracle.pack_hs_cplx(geo, acc0.copy(), d, kp, 0.375)
        assert_sparse_close(got, ref, what="pack k " + name)


def _eig_case(name):
    w, sk, geo, model, selfegy = case(name)
    H = oracle.unpack_hs_real(geo, oracle.build_h0(model, geo, selfegy))
    S = oracle.unpack_hs_real(geo, oracle.build_s(model, geo))
    eig, C = scipy.linalg.eigh(H, S, lower=True)
    return w, sk, geo, model, selfegy, H, S, eig, np.asfortranarray(C), fermi_fillings(eig, n_electrons(w, sk))


@pytest.mark.parametrize("name", ["h2o", "ce_o", "fec128"])
def test_k3_density_matrix_real(name, hp_cache):
    w, sk, geo, model, selfegy, H, S, eig, C, f = _eig_case(name)
    hp = hp_cache(name)
    c0 = C.copy()
    dm = hp.getDensityMatrix(C, f)
    ref = oracle.density_matrix(C, f)
    assert np.array_equal(C, c0)                                     # eigvecs restored / untouched
    assert np.all(np.triu(dm, 1) == 0.0)                             # only the lower triangle is filled
    assert np.abs(dm - ref).max() <= 1e-12 * np.abs(ref).max()
    edm = hp.getEDensityMatrix(C, f, eig)
    assert np.abs(edm - oracle.energy_density_matrix(C, f, eig)).max() <= 1e-12 * np.abs(edm).max()
    # weights of either sign and exact zeros (Methfessel-Paxton-like): one pass, no sqrt
    rng = np.random.default_rng(10)
    f2 = rng.normal(size=len(f)); f2[3] = 0.0; f2[-5:] = 0.0
    dm2 = hp.getDensityMatrix(C, f2)
    assert np.abs(dm2 - oracle.weighted_density_matrix(C, f2)).max() <= 1e-12 * np.abs(dm2).max()
    assert np.abs(dm2 - oracle.density_matrix(C, f2)).max() <= 1e-11 * np.abs(dm2).max()


@pytest.mark.parametrize("n,n_occ", [(1500, 700), (2048, 1024), (1027, 1027)])
def test_k3_large_tiles(n, n_occ, hp_cache):
    """128x128 DMMA tile kernel: aligned (vector cp.async) and odd-ld paths, ragged edges."""
    hp = hp_cache("h2o")
    rng = np.random.default_rng(n)
    C = np.asfortranarray(rng.normal(size=(n, n)) / np.sqrt(n))
    f = np.zeros(n); f[:n_occ] = rng.uniform(0.5, 2.0, n_occ)
    dm = hp.getDensityMatrix(C, f)
    ref = oracle.weighted_density_matrix(C, f)
    assert np.all(np.triu(dm, 1) == 0.0)
    assert np.abs(dm - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("variant", [0, 20, 30, 31, 34, 35, 36])
@pytest.mark.parametrize("n,n_occ", [(1502, 701), (66, 5), (640, 640), (4230, 77)])
def test_k3_tma_feed_variants(variant, n, n_occ, hp_cache, monkeypatch):
    """TMA-fed K3 (cp.async.bulk + mbarrier pipeline, producer warp) in every tile configuration against
    the oracle and against the cp.async kernel (variant 20): ragged row / level edges, a level count
    that is not a multiple of the k-tile, diagonal tiles (one panel), more tiles than CTAs."""
    monkeypatch.setenv("DFTBP_B200_K3_VARIANT", str(variant))
    hp = hp_cache("h2o")
    rng = np.random.default_rng(n + variant)
    C = np.asfortranarray(rng.normal(size=(n, n)) / np.sqrt(n))
    f = np.zeros(n); f[:n_occ] = rng.normal(size=n_occ)
    f[min(2, n_occ - 1)] = 0.0
    dm = hp.getDensityMatrix(C, f)
    ref = oracle.weighted_density_matrix(C, f)
    assert np.all(np.triu(dm, 1) == 0.0)
    assert np.abs(dm - ref).max() <= 1e-12 * np.abs(ref).max()
    # a column panel of the same matrix (multi-GPU split) through the same kernel
    part = hp.getDensityMatrix(C, f, cols=(n // 3 + 1, 2 * n // 3))
    assert np.abs(part[:, n // 3:2 * n // 3] - np.tril(ref)[:, n // 3:2 * n // 3]).max() <= 1e-12 * np.abs(ref).max()


def test_k3_density_matrix_complex(hp_cache):
    w, sk, geo, model, selfegy = case("si2")
    hp = hp_cache("si2")
    h0, s = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    for kp in w["kpoints"]:
        Hk, Sk = oracle.unpack_hs_cplx(geo, h0, kp), oracle.unpack_hs_cplx(geo, s, kp)
        eig, C = scipy.linalg.eigh(Hk, Sk, lower=True)
        C = np.asfortranarray(C)
        f = fermi_fillings(eig, n_electrons(w, sk))
        dm = hp.getDensityMatrix(C, f)
        ref = oracle.density_matrix(C, f)
        assert np.abs(dm - ref).max() <= 1e-12 * np.abs(ref).max()
        assert np.all(np.triu(dm, 1) == 0.0) and np.all(np.diag(dm).imag == 0.0)
    rng = np.random.default_rng(12)
    n = 300
    C = np.asfortranarray((rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n))
    f = rng.normal(size=n); f[-7:] = 0.0
    dm = hp.getDensityMatrix(C, f)
    assert np.abs(dm - oracle.weighted_density_matrix(C, f)).max() <= 1e-12 * np.abs(dm).max()


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128", "water1002"])
def test_k4_mulliken(name, hp_cache):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    rng = np.random.default_rng(13)
    s = oracle.build_s(model, geo)
    rho = rng.normal(size=geo["sparseSize"])
    q0 = rng.normal(size=(geo["nAtom"], model.mOrb))
    got = hp.mulliken(s, rho, qq=q0.copy())
    ref = oracle.mulliken_per_orbital(model, geo, s, rho, qq=q0.copy())
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    qb0 = rng.normal(size=(geo["nAtom"], model.mOrb, model.mOrb))
    gotb = hp.mulliken(s, rho, qq=qb0.copy(), block=True)
    refb = oracle.mulliken_per_block(model, geo, s, rho, qq=qb0.copy())
    assert np.abs(gotb - refb).max() <= 1e-12 * max(1.0, np.abs(refb).max())


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128"])
def test_k5_derivative_shift(name, hp_cache):
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    rng = np.random.default_rng(14)
    s = oracle.build_s(model, geo)
    # density-like sparse inputs: decay like S so the magnitudes are physical
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    ref = oracle.derivative_shift(model, geo, dm[0], edm)
    got = hp.derivative_shift(dm[0], edm)
    assert np.abs(got - ref).max() < TOL_FORCE, np.abs(got - ref).max()
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    for n_spin in (1, 2):
        sh = 0.1 * rng.normal(size=(n_spin, geo["nAtom"], model.mOrb, model.mOrb))
        sh = 0.5 * (sh + sh.transpose(0, 1, 3, 2))
        ref = oracle.derivative_shift(model, geo, dm[:n_spin], edm, shift=sh)
        got = hp.derivative_shift(dm[:n_spin], edm, shift=sh)
        assert np.abs(got - ref).max() < TOL_FORCE, (n_spin, np.abs(got - ref).max())
        assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name,method", [("h2o", 2), ("ce_o", 2), ("fec128", 2), ("scan", 1), ("fec128", 1)])
def test_k5_richardson_and_old_interpolation(name, method, hp_cache):
    """Differentiation = Richardson (nonscc.F90:488-600) and forces with OldSKInterpolation."""
    w, sk, geo, model, selfegy = case(name, method)
    hp = hp_cache(name, method=method)
    rng = np.random.default_rng(15)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    sh = 0.1 * rng.normal(size=(2, geo["nAtom"], model.mOrb, model.mOrb))
    sh = 0.5 * (sh + sh.transpose(0, 1, 3, 2))
    for diff_type in (1, 2):
        ref = oracle.derivative_shift(model, geo, dm[0], edm, diff_type=diff_type)
        got = hp.derivative_shift(dm[0], edm, diff_type=diff_type)
        assert np.abs(got - ref).max() < TOL_FORCE, (diff_type, np.abs(got - ref).max())
        assert np.abs(got - ref).max() < 1e-8 * max(1.0, np.abs(ref).max())
        ref = oracle.derivative_shift(model, geo, dm, edm, shift=sh, diff_type=diff_type)
        got = hp.derivative_shift(dm, edm, shift=sh, diff_type=diff_type)
        assert np.abs(got - ref).max() < TOL_FORCE, (diff_type, np.abs(got - ref).max())
        assert np.abs(got - ref).max() < 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128"])
def test_k6_add_shift(name, hp_cache):
    """addShift_atom / addShift_block (dftb/shift.F90:41-228): diagonal and full block shifts,
    1 and 2 spin channels, accumulate and isInputZero forms, partial (per-rank) sums."""
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    rng = np.random.default_rng(16)
    s = oracle.build_s(model, geo)
    h0 = oracle.build_h0(model, geo, selfegy)
    n_atom, mo = geo["nAtom"], model.mOrb
    for n_spin in (1, 2):
        ham = np.stack([h0] + [0.3 * h0] * (n_spin - 1))
        v_atom = 0.1 * rng.normal(size=(n_spin, n_atom))
        v_diag = np.zeros((n_spin, n_atom, mo, mo))
        idx = np.arange(mo)
        v_diag[:, :, idx, idx] = 0.1 * rng.normal(size=(n_spin, n_atom, mo))
        v_full = 0.1 * rng.normal(size=(n_spin, n_atom, mo, mo))
        v_full = 0.5 * (v_full + v_full.transpose(0, 1, 3, 2))
        for shift in (v_atom, v_diag, v_full):
            for zero in (False, True):
                ref = oracle.add_shift(model, geo, ham, s, shift, is_input_zero=zero)
                got = hp.addShift(ham.copy(), s, shift, is_input_zero=zero)
                assert_sparse_close(got, ref, what="addShift %s nSpin %d ndim %d zero %s" % (name, n_spin, shift.ndim, zero))
    if name == "fec128":
        ham = np.stack([h0, h0])
        parts = sum(hp_cache(name, r, 3).addShift(ham.copy(), s, v_full, is_input_zero=True) for r in range(3))
        assert_sparse_close(parts, oracle.add_shift(model, geo, ham, s, v_full, is_input_zero=True), what="addShift ranks")


@pytest.mark.parametrize("name", ["si2", "fec128", "si512"])
def test_k5_stress(name, hp_cache):
    """getNonSCCStress / getBlockStress (dftb/stress.F90:70-326): same pair loop as derivative_shift,
    including the pairs between an atom and its own periodic image (si2: most pairs are)."""
    w, sk, geo, model, selfegy = case(name)
    hp = hp_cache(name)
    rng = np.random.default_rng(17)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    vol = abs(np.linalg.det(w["latvec"]))
    sh = 0.1 * rng.normal(size=(2, geo["nAtom"], model.mOrb, model.mOrb))
    sh = 0.5 * (sh + sh.transpose(0, 1, 3, 2))
    shd = np.zeros_like(sh)
    idx = np.arange(model.mOrb)
    shd[:, :, idx, idx] = sh[:, :, idx, idx]
    cases = [(dm[0], None, 1), (dm, shd, 1), (dm, sh, 1)]
    if name != "si512":
        cases.append((dm[:1], sh[:1], 2))                 # Richardson derivator
    for d, shift, diff_type in cases:
        ref = oracle.stress(model, geo, d, edm, vol, shift=shift, diff_type=diff_type)
        got = hp.stress(d, edm, vol, shift=shift, diff_type=diff_type)
        assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), (diff_type, np.abs(got - ref).max())
    # forces are unchanged by the stress plumbing (self-image pairs stay excluded there)
    ref = oracle.derivative_shift(model, geo, dm, edm, shift=sh)
    assert np.abs(hp.derivative_shift(dm, edm, shift=sh) - ref).max() < TOL_FORCE
    if name == "fec128":
        parts = sum(hp_cache(name, r, 3).stress(dm, edm, vol, shift=sh) for r in range(3))
        assert np.abs(parts - oracle.stress(model, geo, dm, edm, vol, shift=sh)).max() < 1e-9


@pytest.mark.parametrize("name,n_panel", [("fec128", 3), ("si512", 4), ("water1002", 8)])
def test_k3_column_panels_sum_to_full(name, n_panel, hp_cache):
    """One density matrix split over n_panel GPUs: column panels of the lower triangle + pack of the
    panel's atoms; the ranks' sparse arrays add up to the single-GPU result bit for bit."""
    from dftbplus_b200 import schedule
    w, sk, geo, model, selfegy, H, S, eig, C, f = _eig_case(name)
    hp = hp_cache(name)
    ss = geo["sparseSize"]
    full = hp.packHS(np.zeros(ss), hp.getDensityMatrix(C, f))
    efull = hp.packHS(np.zeros(ss), hp.getEDensityMatrix(C, f, eig))
    parts, eparts = np.zeros(ss), np.zeros(ss)
    for (a0, a1, c0, c1) in schedule.dm_panels(geo["iAtomStart"], n_panel):
        if a0 > a1:
            continue
        part = hp.packHS(np.zeros(ss), hp.getDensityMatrix(C, f, cols=(c0, c1)), atoms=(a0, a1))
        assert np.count_nonzero(part) <= np.count_nonzero(full)
        parts += part
        eparts += hp.packHS(np.zeros(ss), hp.getEDensityMatrix(C, f, eig, cols=(c0, c1)), atoms=(a0, a1))
    assert np.array_equal(parts, full)
    assert np.array_equal(eparts, efull)
    assert_sparse_close(full, oracle.pack_hs_real(geo, np.zeros(ss), oracle.density_matrix(C, f)), what="rho")


@pytest.mark.parametrize("name", ["water1002", "fec128"])
def test_md_like_coordinate_updates(name):
    """BASELINE config 5 pattern: the atoms move every step while the neighbour list stays; the host
    calls update_coords (the device re-sorts its pair schedule) and rebuilds H0 / S and the forces.
    Every step is compared with the oracle on the moved geometry; the images move with their atoms."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case(name)
    rng = np.random.default_rng(23)
    hp = HotPath(0)
    try:
        sk.upload(hp)
        hp.set_geometry(geo)
        s0 = oracle.build_s(model, geo)
        dm = s0 * rng.normal(size=geo["sparseSize"])
        edm = s0 * rng.normal(size=geo["sparseSize"])
        coords = geo["coords"].copy()
        for step in range(4):
            coords = coords + 0.02 * rng.normal(size=(geo["nAtom"], 3))[geo["img2CentCell"] - 1]
            g = dict(geo)
            g["coords"] = np.ascontiguousarray(coords)
            hp.update_coords(g["coords"])
            h, s = hp.buildH0S(sk.selfegy)
            assert_sparse_close(h, oracle.build_h0(model, g, selfegy), what="H0 step %d" % step)
            assert_sparse_close(s, oracle.build_s(model, g), what="S step %d" % step)
            d = hp.derivative_shift(dm, edm)
            assert np.abs(d - oracle.derivative_shift(model, g, dm, edm)).max() < TOL_FORCE
        assert not np.array_equal(s, s0)
    finally:
        hp.close()


def test_many_coordinate_updates_refresh_and_resort():
    """update_coords on an unchanged neighbour list refreshes the records of the pair schedule instead of
    sorting again; the distance order drifts with the atoms until the library sorts again (every 16
    refreshes).  40 steps of a random walk that ends 0.4 bohr away from the start, checked
    against the oracle along the way and with SCC forces of a share of the pairs at the end."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("fec128")
    rng = np.random.default_rng(29)
    hp = HotPath(0)
    try:
        sk.upload(hp)
        hp.set_geometry(geo)
        s0 = oracle.build_s(model, geo)
        dm = s0 * rng.normal(size=(2, geo["sparseSize"]))
        edm = s0 * rng.normal(size=geo["sparseSize"])
        mo, nA = hp.mOrb, geo["nAtom"]
        shift = np.zeros((2, nA, mo, mo))
        shift[:, :, np.arange(mo), np.arange(mo)] = 0.05 * rng.normal(size=(2, nA, mo))
        coords = geo["coords"].copy()
        drift = 0.01 * rng.normal(size=(nA, 3))
        for step in range(40):
            coords = coords + (drift + 0.004 * rng.normal(size=(nA, 3)))[geo["img2CentCell"] - 1]
            g = dict(geo)
            g["coords"] = np.ascontiguousarray(coords)
            hp.update_coords(g["coords"])
            h, s = hp.buildH0S(sk.selfegy)
            if step % 8 == 7 or step >= 37:
                assert_sparse_close(h, oracle.build_h0(model, g, selfegy), what="H0 step %d" % step)
                assert_sparse_close(s, oracle.build_s(model, g), what="S step %d" % step)
                d = hp.derivative_shift(dm, edm, shift=shift)
                assert np.abs(d - oracle.derivative_shift(model, g, dm, edm, shift=shift)).max() < TOL_FORCE, step
        acc = np.zeros_like(d)
        for part in range(3):
            hp.set_pair_partition(part, 3)
            hp.update_coords(g["coords"])
            acc += hp.derivative_shift(dm, edm, shift=shift)
        hp.set_pair_partition(0, 1)
        assert np.abs(acc - oracle.derivative_shift(model, g, dm, edm, shift=shift)).max() < TOL_FORCE
    finally:
        hp.close()


def test_hegv_cusolver(hp_cache):
    w, sk, geo, model, selfegy, H, S, eig, C, f = _eig_case("ce_o")
    hp = hp_cache("ce_o")
    Hc, Sc = H.copy(order="F"), S.copy(order="F")
    ev = hp.diagDenseMtx(Hc, Sc)
    assert np.abs(ev - eig).max() < 1e-10
    Hf, Sf = np.tril(H) + np.tril(H, -1).T, np.tril(S) + np.tril(S, -1).T
    assert np.abs(Hf @ Hc - Sf @ Hc * ev[None, :]).max() < 1e-9       # H C = S C e
    assert np.abs(Hc.T @ Sf @ Hc - np.eye(len(ev))).max() < 1e-9


def test_hegv_large_n_path(hp_cache, monkeypatch):
    """The 64-bit composite path (Xpotrf -> two trsm -> Xsyevd -> trsm) that takes over where the 32-bit
    workspace size of cusolverDn?sygvd overflows (n = 31744), forced here at small n; real and complex."""
    monkeypatch.setenv("DFTBP_B200_HEGV_LARGE", "1")
    w, sk, geo, model, selfegy, H, S, eig, C, f = _eig_case("fec128")
    hp = hp_cache("fec128")
    Hc, Sc = np.tril(H).copy(order="F"), np.tril(S).copy(order="F")     # only the lower triangles are referenced
    ev = hp.diagDenseMtx(Hc, Sc)
    assert np.abs(ev - eig).max() < 1e-10
    Hs, Ss = np.tril(H) + np.tril(H, -1).T, np.tril(S) + np.tril(S, -1).T
    assert np.abs(Hs @ Hc - Ss @ Hc * ev[None, :]).max() < 1e-9
    assert np.abs(Hc.T @ Ss @ Hc - np.eye(len(ev))).max() < 1e-9
    w, sk, geo, model, selfegy = case("si2")
    h0, s = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    kp = w["kpoints"][3]
    Hk, Sk = oracle.unpack_hs_cplx(geo, h0, kp), oracle.unpack_hs_cplx(geo, s, kp)
    Hf = np.tril(Hk) + np.tril(Hk, -1).conj().T
    Sf = np.tril(Sk) + np.tril(Sk, -1).conj().T
    eig = scipy.linalg.eigh(Hf, Sf, eigvals_only=True)
    Hc, Sc = np.tril(Hk).copy(order="F"), np.tril(Sk).copy(order="F")
    ev = hp.diagDenseMtx(Hc, Sc)
    assert np.abs(ev - eig).max() < 1e-10
    assert np.abs(Hf @ Hc - Sf @ Hc * ev[None, :]).max() < 1e-9
    ev_n = hp.diagDenseMtx(np.tril(Hk).copy(order="F"), np.tril(Sk).copy(order="F"), jobz="N")
    assert np.abs(ev_n - eig).max() < 1e-10


def test_device_resident_scc_step(hp_cache):
    """The whole chain on device pointers (torch tensors), no host copies between kernels:
    build -> unpack -> hegvd -> DM -> pack -> Mulliken -> E-DM -> forces; against the oracle chain."""
    import torch
    w, sk, geo, model, selfegy, H, S, eig, C, f = _eig_case("fec128")
    from dftbplus_b200.api import HotPath
    hp = HotPath(0)
    hp.use_torch_stream()          # torch tensors as device pointers: share torch's stream
    sk.upload(hp)
    hp.set_geometry(geo)
    dev = torch.device("cuda:0")
    n, ss = geo["nOrb"], geo["sparseSize"]
    ham = torch.zeros(ss, dtype=torch.float64, device=dev)
    over = torch.zeros(ss, dtype=torch.float64, device=dev)
    from dftbplus_b200.api import new_dense
    Hd, Sd = new_dense(n), new_dense(n)
    hp.buildH0S(sk.selfegy, ham, over)
    hp.unpackHS(ham, square=Hd)
    hp.unpackHS(over, square=Sd)
    ev = hp.diagDenseMtx(Hd, Sd)
    hp.synchronize()
    ev_h = ev.cpu().numpy()
    assert np.abs(ev_h - eig).max() < 1e-9
    fill = fermi_fillings(ev_h, n_electrons(w, sk))
    fd = torch.from_numpy(fill).to(dev)
    dm = hp.getDensityMatrix(Hd, fd)
    rho = torch.zeros(ss, dtype=torch.float64, device=dev)
    hp.packHS(rho, dm)
    qq = torch.zeros((geo["nAtom"], model.mOrb), dtype=torch.float64, device=dev)
    hp.mulliken(over, rho, qq=qq)
    edm = hp.getEDensityMatrix(Hd, fd, ev)
    erho = torch.zeros(ss, dtype=torch.float64, device=dev)
    hp.packHS(erho, edm)
    deriv = torch.zeros((geo["nAtom"], 3), dtype=torch.float64, device=dev)
    hp.derivative_shift(rho, erho, deriv=deriv)
    hp.synchronize()
    torch.cuda.synchronize()
    # oracle chain from the same eigenvectors (eigenvectors are only defined up to sign/rotation in
    # degenerate subspaces, but rho, q and the forces are invariants)
    rho_ref = oracle.pack_hs_real(geo, np.zeros(ss), oracle.density_matrix(C, fermi_fillings(eig, n_electrons(w, sk))))
    assert_sparse_close(rho.cpu().numpy(), rho_ref, rtol=1e-9, what="rho chain")
    q_ref = oracle.mulliken_per_orbital(model, geo, oracle.build_s(model, geo), rho_ref)
    assert np.abs(qq.cpu().numpy() - q_ref).max() < 1e-9
    assert abs(qq.sum().item() - n_electrons(w, sk)) < 1e-8
    erho_ref = oracle.pack_hs_real(geo, np.zeros(ss), oracle.energy_density_matrix(C, fermi_fillings(eig, n_electrons(w, sk)), eig))
    d_ref = oracle.derivative_shift(model, geo, rho_ref, erho_ref)
    assert np.abs(deriv.cpu().numpy() - d_ref).max() < TOL_FORCE
    assert np.abs(deriv.cpu().numpy().sum(axis=0)).max() < 1e-8


def test_full_size_properties_fec4096():
    """BASELINE size (4096 atoms, spd): direct oracle parity for the pair kernels plus the
    size-independent invariants: pack(unpack(x)) == x, sum of forces == 0, idempotent rebuild."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("fec4096")
    hp = HotPath(0)
    sk.upload(hp)
    hp.set_geometry(geo)
    h, s = hp.buildH0S(sk.selfegy)
    assert_sparse_close(h, oracle.build_h0(model, geo, selfegy), what="H0 fec4096")
    assert_sparse_close(s, oracle.build_s(model, geo), what="S fec4096")
    h2, s2 = hp.buildH0S(sk.selfegy)
    assert np.array_equal(h, h2) and np.array_equal(s, s2)
    rng = np.random.default_rng(15)
    rho = s * rng.normal(size=geo["sparseSize"])
    q = hp.mulliken(s, rho)
    assert np.abs(q - oracle.mulliken_per_orbital(model, geo, s, rho)).max() < 1e-11
    edm = s * rng.normal(size=geo["sparseSize"])
    d = hp.derivative_shift(rho, edm)
    assert np.abs(d.sum(axis=0)).max() < 1e-8
    idx = oracle.indices_with_workload(64, 5, 1, geo["nAtom"], geo["nNeighbourSK"])   # a 1/64 sample of atoms
    hp5 = HotPath(0, 5, 64)
    sk.upload(hp5)
    hp5.set_geometry(geo)
    d5 = hp5.derivative_shift(rho, edm)
    ref5 = oracle.derivative_shift(model, geo, rho, edm, iter_indices=idx)
    assert np.abs(d5 - ref5).max() < TOL_FORCE
    hp5.close()
    # device-resident round trip at full size (31744 x 31744 dense, 8 GB, stays on the GPU)
    import torch
    from dftbplus_b200.api import new_dense
    x = rng.normal(size=geo["sparseSize"])
    for a in range(geo["nAtom"]):
        n = geo["nOrbAtom"][a]
        blk = x[geo["iPair"][a, 0]:geo["iPair"][a, 0] + n * n].reshape(n, n)
        blk[:] = 0.5 * (blk + blk.T)
    hp.use_torch_stream()
    xd = torch.from_numpy(x).cuda()
    sq = new_dense(geo["nOrb"])
    hp.unpackHS(xd, square=sq)
    back = torch.zeros_like(xd)
    hp.packHS(back, sq)
    hp.synchronize()
    assert torch.equal(back, xd)
    del sq
    hp.close()

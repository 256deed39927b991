"""Parity at the sizes of the BASELINE configs (round-1 review: "config-size parity holes") and of the
finer force decomposition across GPUs, through the C ABI against the oracle.

Sorted after test_gpu_parity.py on purpose (the small cases fail first and faster)."""
import numpy as np
import pytest

from helpers import (oracle, case, assert_sparse_close, TOL_FORCE)

pytestmark = pytest.mark.gpu


def _diag_shift(rng, n_spin, n_atom, m_orb, scale=0.1):
    sh = np.zeros((n_spin, n_atom, m_orb, m_orb))
    idx = np.arange(m_orb)
    sh[:, :, idx, idx] = scale * rng.normal(size=(n_spin, n_atom, m_orb))
    return sh


def _full_shift(rng, n_spin, n_atom, m_orb, scale=0.1):
    sh = scale * rng.normal(size=(n_spin, n_atom, m_orb, m_orb))
    return 0.5 * (sh + sh.transpose(0, 1, 3, 2))


def test_fec4096_scc_forces_full_size():
    """BASELINE config 4 at full size, the call bench.py times: derivative_shift in the SCC form with two
    spin channels, diagonal (collinear-spin SCC) and full (DFTB+U-like) shift blocks, every atom
    against the oracle (forces.F90:548-669)."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("fec4096")
    hp = HotPath(0)
    sk.upload(hp)
    hp.set_geometry(geo)
    oracle.set_num_threads()
    rng = np.random.default_rng(41)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    for sh in (_diag_shift(rng, 2, geo["nAtom"], model.mOrb), _full_shift(rng, 2, geo["nAtom"], model.mOrb)):
        ref = oracle.derivative_shift(model, geo, dm, edm, shift=sh)
        got = hp.derivative_shift(dm, edm, shift=sh)
        err = np.abs(got - ref).max()
        assert err < TOL_FORCE and err < 1e-9 * max(1.0, np.abs(ref).max()), err
    # the same call dealt out over 8 GPUs chunk-wise (dftbp_b200_set_pair_partition): the parts add up
    sh = _diag_shift(rng, 2, geo["nAtom"], model.mOrb)
    full = hp.derivative_shift(dm, edm, shift=sh)
    parts = np.zeros_like(full)
    for part in range(8):
        hp.set_pair_partition(part, 8)
        parts += hp.derivative_shift(dm, edm, shift=sh)
    hp.set_pair_partition(0, 1)
    assert np.abs(parts - full).max() < 1e-11 * max(1.0, np.abs(full).max())
    hp.close()


@pytest.mark.parametrize("name", ["fec128", "si2", "ce_o"])
def test_pair_partition_sums_to_full(name):
    """dftbp_b200_set_pair_partition: forces (non-SCC, diagonal and full shifts, Richardson) and the stress
    tensor of the parts add up to the unpartitioned result and to the oracle."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case(name)
    hp = HotPath(0)
    sk.upload(hp)
    hp.set_geometry(geo)
    rng = np.random.default_rng(43)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    shd, shf = _diag_shift(rng, 2, geo["nAtom"], model.mOrb), _full_shift(rng, 2, geo["nAtom"], model.mOrb)
    calls = [dict(dm=dm[0], shift=None, diff_type=1), dict(dm=dm, shift=shd, diff_type=1),
             dict(dm=dm, shift=shf, diff_type=1), dict(dm=dm[:1], shift=shd[:1], diff_type=2)]
    for c in calls:
        ref = oracle.derivative_shift(model, geo, c["dm"], edm, shift=c["shift"], diff_type=c["diff_type"])
        for n_parts in (2, 5):
            acc = np.zeros_like(ref)
            for part in range(n_parts):
                hp.set_pair_partition(part, n_parts)
                acc += hp.derivative_shift(c["dm"], edm, shift=c["shift"], diff_type=c["diff_type"])
            tol = 1e-8 * max(1.0, np.abs(ref).max()) if c["diff_type"] == 2 else TOL_FORCE
            assert np.abs(acc - ref).max() < tol, (c["diff_type"], n_parts, np.abs(acc - ref).max())
    if w["latvec"] is not None:
        vol = abs(np.linalg.det(w["latvec"]))
        ref = oracle.stress(model, geo, dm, edm, vol, shift=shd)
        acc = np.zeros((3, 3))
        for part in range(3):
            hp.set_pair_partition(part, 3)
            acc += hp.stress(dm, edm, vol, shift=shd)
        assert np.abs(acc - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    hp.set_pair_partition(0, 1)
    hp.close()


def test_pair_partition_across_contexts():
    """The shares of dftbp_b200_set_pair_partition taken by DIFFERENT contexts (one per GPU in a multi-GPU
    run), each with its own device-built pair schedule, rebuilt several times: the schedule's order
    inside a distance bin is not deterministic, the shares must be disjoint and complete anyway."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("fec128")
    rng = np.random.default_rng(47)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    shd = _diag_shift(rng, 2, geo["nAtom"], model.mOrb)
    ref = oracle.derivative_shift(model, geo, dm, edm, shift=shd)
    n_parts = 4
    ctxs = []
    for part in range(n_parts):
        hp = HotPath(0)
        sk.upload(hp)
        hp.set_geometry(geo)
        hp.set_pair_partition(part, n_parts)
        ctxs.append(hp)
    try:
        for rep in range(3):
            acc = np.zeros_like(ref)
            for hp in ctxs:
                hp.update_coords(geo["coords"])           # a new schedule, as in every geometry step
                hp.buildH0S(sk.selfegy)                   # (the full build shares the context)
                acc += hp.derivative_shift(dm, edm, shift=shd)
            assert np.abs(acc - ref).max() < TOL_FORCE, (rep, np.abs(acc - ref).max())
    finally:
        for hp in ctxs:
            hp.close()


def test_water2046_pair_kernels():
    """BASELINE config 5 geometry (682 H2O, 2046 atoms): H0 / S, Mulliken and SCC forces vs the oracle."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("water2046")
    hp = HotPath(0)
    sk.upload(hp)
    hp.set_geometry(geo)
    oracle.set_num_threads()
    h, s = hp.buildH0S(sk.selfegy)
    assert_sparse_close(h, oracle.build_h0(model, geo, selfegy), what="H0 water2046")
    assert_sparse_close(s, oracle.build_s(model, geo), what="S water2046")
    rng = np.random.default_rng(44)
    rho = s * rng.normal(size=geo["sparseSize"])
    edm = s * rng.normal(size=geo["sparseSize"])
    q = hp.mulliken(s, rho)
    assert np.abs(q - oracle.mulliken_per_orbital(model, geo, s, rho)).max() < 1e-11
    sh = _diag_shift(rng, 1, geo["nAtom"], model.mOrb)
    for shift in (None, sh):
        ref = oracle.derivative_shift(model, geo, rho, edm, shift=shift)
        got = hp.derivative_shift(rho, edm, shift=shift)
        assert np.abs(got - ref).max() < TOL_FORCE
    hp.close()


def test_si512_kpoint_chain_all_32_kpoints():
    """BASELINE config 3: every k-point of the 4x4x4 Monkhorst-Pack mesh through unpack -> density matrix
    (complex, weighted) -> pack with the k-point weight, accumulated into one sparse rho like
    getDensityFromCplxEigvecs (main.F90:4289-4327); against the oracle."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case("si512")
    hp = HotPath(0)
    sk.upload(hp)
    hp.set_geometry(geo)
    n = geo["nOrb"]
    kpts, wts = w["kpoints"], w["kweights"]
    assert len(kpts) == 32 and abs(wts.sum() - 1.0) < 1e-12
    rng = np.random.default_rng(45)
    x = oracle.build_s(model, geo)
    n_lev = 96
    fill = np.zeros(n)
    fill[:n_lev] = 2.0 * rng.uniform(0.2, 1.0, n_lev)
    rho, rho_ref = np.zeros(geo["sparseSize"]), np.zeros(geo["sparseSize"])
    for ik, (kp, wk) in enumerate(zip(kpts, wts)):
        if ik % 8 == 0:     # the Bloch sum itself at a quarter of the mesh (it is the same routine for all)
            sq, sq_ref = hp.unpackHS(x, kpoint=kp), oracle.unpack_hs_cplx(geo, x, kp)
            assert np.abs(sq - sq_ref).max() <= 1e-12 * np.abs(sq_ref).max()
        C = np.zeros((n, n), order="F", dtype=np.complex128)
        C[:, :n_lev] = (rng.normal(size=(n, n_lev)) + 1j * rng.normal(size=(n, n_lev))) / np.sqrt(n)
        dmk = hp.getDensityMatrix(C, fill)
        ref = oracle.density_matrix(C, fill)
        assert np.abs(dmk - ref).max() <= 1e-12 * np.abs(ref).max()
        hp.packHS(rho, dmk, kpoint=kp, kweight=wk)
        oracle.pack_hs_cplx(geo, rho_ref, ref, kp, wk)
    assert_sparse_close(rho, rho_ref, what="rho over 32 k-points")
    hp.close()


def test_fec4096_density_matrix_sampled_entries():
    """K3 at n = 31744 (BASELINE config 4): 10^4 random entries of the lower triangle against
    dm[i, j] = sum_k f_k C[i, k] C[j, k] evaluated independently (torch, FP64)."""
    import torch
    from dftbplus_b200.api import HotPath, new_dense
    w, sk, geo, model, selfegy = case("fec4096")
    hp = HotPath(0)
    hp.use_torch_stream()
    n = geo["nOrb"]
    n_occ = 14336
    g = torch.Generator(device="cuda").manual_seed(5)
    C = new_dense(n)
    C.copy_(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g).mul_(n ** -0.5))
    fill = torch.zeros(n, dtype=torch.float64, device="cuda")
    fill[:n_occ] = 2.0 * torch.rand(n_occ, dtype=torch.float64, device="cuda", generator=g)
    dm = new_dense(n)
    hp.getDensityMatrix(C, fill, dm=dm)
    hp.synchronize()
    rng = np.random.default_rng(46)
    i = rng.integers(0, n, 10000)
    j = rng.integers(0, n, 10000)
    i, j = np.maximum(i, j), np.minimum(i, j)
    it, jt = torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda()
    Cm, dmm = C, dm     # new_dense: column-major storage, indexed [row, col]
    ref = (Cm[it, :n_occ] * fill[:n_occ] * Cm[jt, :n_occ]).sum(dim=1)
    got = dmm[it, jt]
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1e-12 * max(scale, 1.0), float((got - ref).abs().max())
    # strict upper triangle untouched / zero like the reference's 'L' storage
    up = dmm[jt[i != j], it[i != j]]
    assert float(up.abs().max()) == 0.0
    del C, dm
    hp.close()


# ---- species that the .skf format cannot describe: repeated and unordered shells ---------------------
def _table(rng, n_grid, n_int, dist):
    r = dist * np.arange(1, n_grid + 1)
    amp = rng.uniform(0.1, 0.5, n_int) * rng.choice([-1.0, 1.0], n_int)
    a, b = rng.uniform(0.2, 0.6, n_int), rng.uniform(0.8, 1.4, n_int)
    return np.asfortranarray(amp[None, :] * (1.0 + a[None, :] * r[:, None]) * np.exp(-b[None, :] * r[:, None]))


@pytest.mark.parametrize("shells", [([0, 0, 1], [0, 1, 2, 2]), ([1, 0], [2, 1, 0, 1]), ([0, 1, 2], [0, 0]), ([3, 0], [0, 1, 2])])
def test_non_canonical_shell_sets(shells):
    """Advisor finding of round 1: species with repeated or unordered shells (s s* p, s p d d*, ...) go through
    the generic paths of the pair kernels -- staging rows and write-out sized from the real block, not from
    (lmax + 1)^2.  Tables come in through set_sk_table on both sides (no .skf file can hold such a species)."""
    from dftbplus_b200.api import HotPath
    from dftbplus_b200 import geometry
    rng = np.random.default_rng(47)
    dist, n_grid = 0.02, 300
    model = oracle.Model(list(shells))
    hp = HotPath(0)
    hp.set_orbitals(list(shells))
    for i in (1, 2):
        for j in (1, 2):
            n_int = model.n_integrals(i, j)
            for which in (0, 1):
                tab = _table(rng, n_grid, n_int, dist)
                model.set_table(which, i, j, dist, tab)
                hp.set_sk_table(which, i, j, dist, tab)
    cutoff = model.cutoff()
    assert abs(cutoff - hp.sk_cutoff()) < 1e-12
    n_atom = 40
    coord0 = rng.uniform(0.0, 11.0, size=(n_atom, 3))
    species0 = rng.integers(1, 3, n_atom).astype(np.int32)
    geo = geometry.build_geometry(coord0, species0, model.nOrbSpecies, np.eye(3) * 11.0, cutoff)
    hp.set_geometry(geo)
    selfegy = [list(rng.normal(size=len(s))) for s in shells]
    h, s = hp.buildH0S(selfegy)
    assert_sparse_close(h, oracle.build_h0(model, geo, selfegy), what="H0 " + str(shells))
    assert_sparse_close(s, oracle.build_s(model, geo), what="S " + str(shells))
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    q = hp.mulliken(s, dm[0])
    assert np.abs(q - oracle.mulliken_per_orbital(model, geo, s, dm[0])).max() < 1e-11
    for sh in (None, _diag_shift(rng, 2, n_atom, model.mOrb), _full_shift(rng, 2, n_atom, model.mOrb)):
        d = dm[0] if sh is None else dm
        ref = oracle.derivative_shift(model, geo, d, edm, shift=sh)
        got = hp.derivative_shift(d, edm, shift=sh)
        assert np.abs(got - ref).max() < TOL_FORCE, (shells, np.abs(got - ref).max())
    sq = hp.unpackHS(s)
    assert np.array_equal(sq, oracle.unpack_hs_real(geo, s))
    hp.close()

"""ctypes front end of the CPU parity oracle (oracle/dftb_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Array conventions (shared with the CUDA library's Python host mirror, so the same numpy
buffers can be handed to both): memory is exactly the reference's Fortran layout; numpy sees
it C-ordered with the dimensions reversed, e.g. Fortran coords(3, nAllAtom) == numpy
coords[nAllAtom, 3], iNeighbour(0:mNeigh, nAtom) == iNeighbour[nAtom, mNeigh+1],
qq(mOrb, nAtom) == qq[nAtom, mOrb], shift(mOrb, mOrb, nAtom, nSpin) == shift[nSpin, nAtom,
mOrb, mOrb] (last index = Fortran first).  Dense matrices are numpy order='F' arrays indexed
[row, col].  Atom / species values stay 1-based, iPair stays 0-based.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "dftb_oracle.c")
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")

# Release flags of the reference build (sys/gnu.cmake:27) + OpenMP; no FMA contraction.
CFLAGS = ["-O2", "-funroll-all-loops", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared"]

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lib = None


def build(force=False):
    """Compile the C restatement (gcc only; a few seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc"] + CFLAGS + ["-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_model_create.restype = ctypes.c_void_p
        _lib.orc_model_cutoff.restype = ctypes.c_double
        _lib.orc_sk_cutoff.restype = ctypes.c_double
        _lib.orc_init()
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n=None):
    """OpenMP threads of the restatement; None = all host cores (the scheduler affinity mask)."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def sk_eval(tab, dist, r, method=2):
    """SlakoEqGrid_getSKIntegrals for one table; tab is [nGrid, nInteg] order='F'."""
    tab = np.asfortranarray(tab, dtype=np.float64)
    out = np.zeros(tab.shape[1])
    lib().orc_sk_eval(tab.shape[0], tab.shape[1], ctypes.c_double(dist), _d(tab), method,
                      ctypes.c_double(r), _d(out))
    return out


def chunk_ranges(group_size, rank, first, last):
    a, b = ctypes.c_int(), ctypes.c_int()
    lib().orc_chunk_ranges(group_size, rank, first, last, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def indices_with_workload(group_size, rank, first, last, workload):
    workload = np.ascontiguousarray(workload, dtype=np.int32)
    out = np.zeros(last - first + 1, dtype=np.int32)
    n = lib().orc_indices_with_workload(group_size, rank, first, last, _i(workload), _i(out))
    return out[:n].copy()


class Model:
    """Flattened TOrbitals + TSlakoCont (H and S)."""

    def __init__(self, ang_shells):
        """ang_shells: per species, the list of shell angular momenta, e.g. [[0,1],[0]]."""
        self.ang_shells = [list(a) for a in ang_shells]
        self.nSpecies = len(ang_shells)
        self.mShell = max(len(a) for a in ang_shells)
        self.nShell = np.array([len(a) for a in ang_shells], dtype=np.int32)
        self.angShell = np.zeros((self.nSpecies, self.mShell), dtype=np.int32)
        for i, a in enumerate(ang_shells):
            self.angShell[i, :len(a)] = a
        self.nOrbSpecies = np.array([sum(2 * l + 1 for l in a) for a in ang_shells], dtype=np.int32)
        self.mOrb = int(self.nOrbSpecies.max())
        self.tables = {}
        self._h = ctypes.c_void_p(lib().orc_model_create(self.nSpecies, self.mShell, _i(self.nShell),
                                                         _i(self.angShell)))

    def n_integrals(self, sp1, sp2):
        return sum(min(l1, l2) + 1 for l1 in self.ang_shells[sp1 - 1] for l2 in self.ang_shells[sp2 - 1])

    def set_table(self, which, sp1, sp2, dist, tab, method=2):
        tab = np.asfortranarray(tab, dtype=np.float64)
        assert tab.shape[1] == self.n_integrals(sp1, sp2)
        self.tables[(which, sp1, sp2)] = (dist, tab, method)
        lib().orc_model_set_table(self._h, which, sp1, sp2, tab.shape[0], tab.shape[1],
                                  ctypes.c_double(dist), _d(tab), method)

    def cutoff(self):
        return lib().orc_model_cutoff(self._h)

    def __del__(self):
        try:
            lib().orc_model_destroy(self._h)
        except Exception:
            pass


def rotate_h0(model, sk, direction, sp1, sp2):
    """rotateH0 for one species pair; returns hh[nOrb2, nOrb1] (rows = species 2)."""
    sk = np.ascontiguousarray(sk, dtype=np.float64)
    hh = np.zeros((model.mOrb, model.mOrb), order="F")
    lib().orc_rotate_h0(model._h, _d(hh), _d(sk), ctypes.c_double(direction[0]),
                        ctypes.c_double(direction[1]), ctypes.c_double(direction[2]), int(sp1), int(sp2))
    return hh[:model.nOrbSpecies[sp2 - 1], :model.nOrbSpecies[sp1 - 1]].copy()


def cell_translations(latvec, cutoff):
    """latvec[i] = i-th lattice vector (numpy [3,3], row i == Fortran latVec(:, i))."""
    latvec = np.ascontiguousarray(latvec, dtype=np.float64)
    # recVec2p(:, i) = i-th reciprocal vector without the 2 pi (initprogram.F90: invLatVec^T)
    rec = np.ascontiguousarray(np.linalg.inv(latvec).T)
    n = lib().orc_cell_translations(_d(latvec), _d(rec), ctypes.c_double(cutoff), None, None)
    cell = np.zeros((n, 3))
    rcell = np.zeros((n, 3))
    lib().orc_cell_translations(_d(latvec), _d(rec), ctypes.c_double(cutoff), _d(cell), _d(rcell))
    return cell, rcell


def neighbour_geometry(model, coord0, species0, latvec=None, cutoff=None):
    """handleCoordinateChange restated: cell translations, updateNeighbourList,
    getNrOfNeighboursForAll, getSparseDescriptor (main.F90:2413-2577).  Returns a dict with
    every array of SURVEY.md section 8.0."""
    L = lib()
    coord0 = np.ascontiguousarray(coord0, dtype=np.float64)
    species0 = np.ascontiguousarray(species0, dtype=np.int32)
    nAtom = coord0.shape[0]
    if cutoff is None:
        cutoff = model.cutoff()
    if latvec is None:
        cell = np.zeros((1, 3))
        rcell = np.zeros((1, 3))
    else:
        cell, rcell = cell_translations(latvec, cutoff)
    maxAll, maxNeigh = max(4 * nAtom, 64), 64
    while True:
        coord = np.zeros((maxAll, 3))
        img = np.zeros(maxAll, dtype=np.int32)
        icv = np.zeros(maxAll, dtype=np.int32)
        nN = np.zeros(nAtom, dtype=np.int32)
        iN = np.zeros((nAtom, maxNeigh + 1), dtype=np.int32)
        d2 = np.zeros((nAtom, maxNeigh + 1))
        nAll, mN = ctypes.c_int(), ctypes.c_int()
        perr = np.zeros(2, dtype=np.int32)
        rc = L.orc_update_neighbour_list(nAtom, _d(coord0), ctypes.c_double(cutoff), cell.shape[0],
                                         _d(rcell), maxAll, maxNeigh, _d(coord), _i(img), _i(icv),
                                         _i(nN), _i(iN), _d(d2), ctypes.byref(nAll),
                                         ctypes.byref(mN), _i(perr), 0)
        if rc == -1:
            raise RuntimeError("Atoms %d and %d too close together" % (perr[0], perr[1]))
        if rc == 0:
            break
        maxAll = max(maxAll, nAll.value)
        maxNeigh = max(maxNeigh, mN.value)
    nAll, mN = nAll.value, mN.value
    nSK = np.zeros(nAtom, dtype=np.int32)
    L.orc_nr_of_neighbours(nAtom, _i(nN), _d(d2), maxNeigh, ctypes.c_double(cutoff), _i(nSK))
    mNeigh = int(nSK.max()) if nAtom else 0
    iNeighbour = np.ascontiguousarray(iN[:, :mNeigh + 1])
    coords = np.ascontiguousarray(coord[:nAll])
    img2CentCell = np.ascontiguousarray(img[:nAll])
    iCellVec = np.ascontiguousarray(icv[:nAll])
    species = np.ascontiguousarray(species0[img2CentCell - 1])
    nOrbAtom = np.ascontiguousarray(model.nOrbSpecies[species0 - 1])
    iPair = np.zeros((nAtom, mNeigh + 1), dtype=np.int32)
    sparseSize = L.orc_sparse_descriptor(nAtom, _i(iNeighbour), mNeigh, _i(nSK), _i(img2CentCell),
                                         _i(nOrbAtom), _i(iPair))
    iAtomStart = np.ones(nAtom + 1, dtype=np.int32)
    iAtomStart[1:] = 1 + np.cumsum(nOrbAtom)
    return dict(nAtom=nAtom, nAllAtom=nAll, mNeigh=mNeigh, coords=coords, species=species,
                species0=species0, img2CentCell=img2CentCell, iCellVec=iCellVec, cellVec=cell,
                rCellVec=rcell, nNeighbourSK=nSK, iNeighbour=iNeighbour, iPair=iPair,
                iAtomStart=iAtomStart, sparseSize=int(sparseSize), nOrbAtom=nOrbAtom,
                nOrb=int(iAtomStart[-1] - 1), cutoff=cutoff, latvec=latvec,
                neighDist2=np.ascontiguousarray(d2[:, :mNeigh + 1]))


def fold_coords_to_cell(coord0, latvec):
    """foldCoordsToCell (pbc3d): coord0[nAtom,3], latvec[i] = i-th lattice vector.  Returns a copy."""
    c = np.array(coord0, dtype=np.float64, order="C", copy=True)
    lat = np.ascontiguousarray(latvec, dtype=np.float64)
    lib().orc_fold_coords_to_cell(c.shape[0], _d(c), _d(lat))
    return c


def neighbour_list_full(n_orb_species, coord0, species0, rcell, cutoff, sk_cutoff, symmetric=False):
    """updateNeighbourList(cutoff) + getNrOfNeighboursForAll(sk_cutoff) + getSparseDescriptor with
    the arrays exactly as the reference holds them (rows of iNeighbour / neighDist2 / iPair are
    0:mNeigh with mNeigh = maxval(nNeighbour) of the LIST cutoff).  rcell[nCellVec,3], zero vector
    first.  symmetric: the symmetric list (every image within reach of ANY central atom, all pairs in both
    directions; symNeighbourList of the hybrid functionals, main.F90:2587).  Raises
    RuntimeError("Atoms i and j too close together") like the reference."""
    L = lib()
    coord0 = np.ascontiguousarray(coord0, dtype=np.float64)
    species0 = np.ascontiguousarray(species0, dtype=np.int32)
    rcell = np.ascontiguousarray(rcell, dtype=np.float64)
    n_orb_species = np.asarray(n_orb_species, dtype=np.int32)
    nAtom = coord0.shape[0]
    maxAll, maxNeigh = max(4 * nAtom, 64), 64
    while True:
        coord = np.zeros((maxAll, 3))
        img = np.zeros(maxAll, dtype=np.int32)
        icv = np.zeros(maxAll, dtype=np.int32)
        nN = np.zeros(nAtom, dtype=np.int32)
        iN = np.zeros((nAtom, maxNeigh + 1), dtype=np.int32)
        d2 = np.zeros((nAtom, maxNeigh + 1))
        nAll, mN = ctypes.c_int(), ctypes.c_int()
        perr = np.zeros(2, dtype=np.int32)
        rc = L.orc_update_neighbour_list(nAtom, _d(coord0), ctypes.c_double(cutoff), rcell.shape[0],
                                         _d(rcell), maxAll, maxNeigh, _d(coord), _i(img), _i(icv),
                                         _i(nN), _i(iN), _d(d2), ctypes.byref(nAll),
                                         ctypes.byref(mN), _i(perr), 1 if symmetric else 0)
        if rc == -1:
            raise RuntimeError("Atoms %d and %d too close together" % (perr[0], perr[1]))
        if rc == 0 and maxNeigh == mN.value:
            break
        maxAll = max(maxAll, nAll.value)
        maxNeigh = mN.value
    nAll, mN = nAll.value, mN.value
    nSK = np.zeros(nAtom, dtype=np.int32)
    L.orc_nr_of_neighbours(nAtom, _i(nN), _d(d2), mN, ctypes.c_double(sk_cutoff), _i(nSK))
    img2CentCell = np.ascontiguousarray(img[:nAll])
    nOrbAtom = np.ascontiguousarray(n_orb_species[species0 - 1])
    iPair = np.zeros((nAtom, mN + 1), dtype=np.int32)
    sparseSize = L.orc_sparse_descriptor(nAtom, _i(iN), mN, _i(nSK), _i(img2CentCell), _i(nOrbAtom), _i(iPair))
    iAtomStart = np.ones(nAtom + 1, dtype=np.int32)
    iAtomStart[1:] = 1 + np.cumsum(nOrbAtom)
    return dict(nAtom=nAtom, nAllAtom=nAll, mNeigh=mN, coords=np.ascontiguousarray(coord[:nAll]),
                species=np.ascontiguousarray(species0[img2CentCell - 1]), img2CentCell=img2CentCell,
                iCellVec=np.ascontiguousarray(icv[:nAll]), nNeighbour=nN, nNeighbourSK=nSK, iNeighbour=iN,
                neighDist2=d2, iPair=iPair, iAtomStart=iAtomStart, sparseSize=int(sparseSize),
                nOrb=int(iAtomStart[-1] - 1))


def _build(model, geo, selfegy, rank, nranks):
    out = np.zeros(geo["sparseSize"])
    se = None
    if selfegy is not None:
        se = np.zeros((model.nSpecies, model.mShell))
        for i, row in enumerate(selfegy):
            se[i, :len(row)] = row
    lib().orc_build_sparse(model._h, _d(se), geo["nAtom"], _d(geo["coords"]), _i(geo["nNeighbourSK"]),
                           _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["species"]), _i(geo["iPair"]),
                           geo["sparseSize"], _d(out), rank, nranks)
    return out


def build_h0(model, geo, selfegy, rank=0, nranks=1):
    """buildH0 (nonscc.F90:66-124); with nranks > 1 returns this rank's zero-padded chunk."""
    return _build(model, geo, selfegy, rank, nranks)


def build_s(model, geo, rank=0, nranks=1):
    """buildS (nonscc.F90:128-181)."""
    return _build(model, geo, None, rank, nranks)


def unpack_hs_real(geo, orig):
    n = geo["nOrb"]
    sq = np.zeros((n, n), order="F")
    lib().orc_unpack_hs_real(_d(sq), n, _d(np.ascontiguousarray(orig)), geo["nAtom"], _i(geo["iNeighbour"]),
                             geo["mNeigh"], _i(geo["nNeighbourSK"]), _i(geo["iAtomStart"]),
                             _i(geo["iPair"]), _i(geo["img2CentCell"]))
    return sq


def unpack_hs_cplx(geo, orig, kpoint):
    n = geo["nOrb"]
    sq = np.zeros((n, n), order="F", dtype=np.complex128)
    kp = np.ascontiguousarray(kpoint, dtype=np.float64)
    lib().orc_unpack_hs_cplx(sq.ctypes.data_as(_dp), n, _d(np.ascontiguousarray(orig)), _d(kp), geo["nAtom"],
                             _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["nNeighbourSK"]),
                             _i(geo["iCellVec"]), _d(geo["cellVec"]), _i(geo["iAtomStart"]),
                             _i(geo["iPair"]), _i(geo["img2CentCell"]))
    return sq


def pack_hs_real(geo, primitive, square):
    """packHS_real: accumulates into primitive (in place) and returns it."""
    sq = np.asfortranarray(square, dtype=np.float64)
    lib().orc_pack_hs_real(_d(primitive), _d(sq), geo["nOrb"], geo["nAtom"], _i(geo["iNeighbour"]),
                           geo["mNeigh"], _i(geo["nNeighbourSK"]), _i(geo["iAtomStart"]),
                           _i(geo["iPair"]), _i(geo["img2CentCell"]))
    return primitive


def pack_hs_cplx(geo, primitive, square, kpoint, kweight):
    sq = np.asfortranarray(square, dtype=np.complex128)
    kp = np.ascontiguousarray(kpoint, dtype=np.float64)
    lib().orc_pack_hs_cplx(_d(primitive), sq.ctypes.data_as(_dp), geo["nOrb"], _d(kp),
                           ctypes.c_double(kweight), geo["nAtom"], _i(geo["iNeighbour"]), geo["mNeigh"],
                           _i(geo["nNeighbourSK"]), _i(geo["iCellVec"]), _d(geo["cellVec"]),
                           _i(geo["iAtomStart"]), _i(geo["iPair"]), _i(geo["img2CentCell"]))
    return primitive


_EPS = np.finfo(np.float64).eps


def _n_levels(filling):
    n = len(filling)
    for ii in range(n, 0, -1):
        if abs(filling[ii - 1]) >= _EPS:
            return ii
    return 1


def _herk_lower(c, alpha=1.0):
    """herk(dm, a) = BLAS xSYRK/xHERK 'L','N' (math/blasroutines.F90:1430-1514): only the lower
    triangle is referenced/written; the strict upper triangle keeps the zeros of dm(:,:)=0."""
    return np.tril(alpha * (c @ c.conj().T))


def density_matrix(eigvecs, filling):
    """fullDensityMatrix_real / _cmplx (densitymatrix.F90:530-658), both branches."""
    n_lev = _n_levels(filling)
    f = np.asarray(filling[:n_lev], dtype=np.float64)
    c = np.asarray(eigvecs)[:, :n_lev]
    shift = f.min()
    if shift > _EPS:
        return _herk_lower(c * np.sqrt(f)[None, :])
    dm = _herk_lower(c)
    shift = shift - 0.1  # arbitraryConstant, densitymatrix.F90:103
    return _herk_lower(c * np.sqrt(f - shift)[None, :]) + shift * dm


def energy_density_matrix(eigvecs, filling, eigen):
    """fullEnergyDensityMatrix_real / _cmplx (densitymatrix.F90:664-806)."""
    n_lev = _n_levels(filling)
    fp = np.asarray(filling[:n_lev]) * np.asarray(eigen[:n_lev])
    c = np.asarray(eigvecs)[:, :n_lev]
    mn, mx = fp.min(), fp.max()
    if ((mn < 0.0) == (mx < 0.0)) and abs(mn) > _EPS and abs(mx) > _EPS:
        return _herk_lower(c * np.sqrt(np.abs(fp))[None, :], alpha=np.sign(mx) if mx != 0 else 1.0)
    dm = _herk_lower(c)
    shift = mn - 0.1
    return _herk_lower(c * np.sqrt(fp - shift)[None, :]) + shift * dm


def weighted_density_matrix(eigvecs, weights):
    """The algebraic content of both routines above: lower(C diag(w) C^H) over the first nLevels
    columns, nLevels = last |w| >= eps."""
    n_lev = _n_levels(weights)
    c = np.asarray(eigvecs)[:, :n_lev]
    return np.tril((c * np.asarray(weights[:n_lev])[None, :]) @ c.conj().T)


def add_shift(model, geo, ham, over, shift, is_input_zero=False, iter_indices=None):
    """addShift (dftb/shift.F90:41-228): ham[nSpin, sparseSize] += 1/2 (S V_i + V_j S).
    shift[nSpin, nAtom] -> addShift_atom, shift[nSpin, nAtom, mOrb, mOrb] -> addShift_block.
    Returns the updated ham (a copy)."""
    ham = np.array(np.atleast_2d(ham), dtype=np.float64, order="C")
    n_spin = ham.shape[0]
    over = np.ascontiguousarray(over, dtype=np.float64)
    shift = np.ascontiguousarray(shift, dtype=np.float64)
    it = None if iter_indices is None else np.ascontiguousarray(iter_indices, dtype=np.int32)
    n_it = 0 if it is None else len(it)
    if shift.ndim == 2:
        lib().orc_add_shift_atom(_d(ham), _d(over), geo["sparseSize"], n_spin, geo["nAtom"], _i(geo["nOrbAtom"]),
                                 _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["nNeighbourSK"]),
                                 _i(geo["img2CentCell"]), _i(geo["iPair"]), _d(shift), int(is_input_zero),
                                 _i(it), n_it)
    else:
        lib().orc_add_shift_block(_d(ham), _d(over), geo["sparseSize"], n_spin, geo["nAtom"], model.mOrb,
                                  _i(geo["nOrbAtom"]), _i(geo["iNeighbour"]), geo["mNeigh"],
                                  _i(geo["nNeighbourSK"]), _i(geo["img2CentCell"]), _i(geo["iPair"]), _d(shift),
                                  int(is_input_zero), _i(it), n_it)
    return ham


def mulliken_per_orbital(model, geo, over, rho, qq=None, iter_indices=None):
    if qq is None:
        qq = np.zeros((geo["nAtom"], model.mOrb))
    it = None if iter_indices is None else np.ascontiguousarray(iter_indices, dtype=np.int32)
    lib().orc_mulliken_per_orbital(_d(qq), model.mOrb, _d(np.ascontiguousarray(over)),
                                   _d(np.ascontiguousarray(rho)), geo["nAtom"], _i(geo["nOrbAtom"]),
                                   _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["nNeighbourSK"]),
                                   _i(geo["img2CentCell"]), _i(geo["iPair"]), _i(it),
                                   0 if it is None else len(it))
    return qq


def mulliken_per_atom(model, geo, over, rho, qq=None, iter_indices=None):
    """mullikenPerAtom, populations.F90:68-112: qq(:) = qq + sum(qPerOrbital, dim=1) (orbitals summed
    in index order)."""
    q_orb = mulliken_per_orbital(model, geo, over, rho, iter_indices=iter_indices)
    s = np.zeros(geo["nAtom"])
    for o in range(q_orb.shape[1]):
        s = s + q_orb[:, o]
    return s if qq is None else qq + s


def onsite_population(geo, rho):
    """getOnsitePopulation, populations.F90:1195-1222: trace of the on-site block of every atom."""
    qq = np.zeros(geo["nAtom"])
    for i in range(geo["nAtom"]):
        n = int(geo["iAtomStart"][i + 1] - geo["iAtomStart"][i])
        o0 = int(geo["iPair"][i, 0])
        s = 0.0
        for k in range(n):
            s = s + rho[o0 + k * (n + 1)]
        qq[i] = s
    return qq


def mulliken_per_block(model, geo, over, rho, qq=None, iter_indices=None):
    if qq is None:
        qq = np.zeros((geo["nAtom"], model.mOrb, model.mOrb))
    it = None if iter_indices is None else np.ascontiguousarray(iter_indices, dtype=np.int32)
    lib().orc_mulliken_per_block(_d(qq), model.mOrb, _d(np.ascontiguousarray(over)),
                                 _d(np.ascontiguousarray(rho)), geo["nAtom"], _i(geo["nOrbAtom"]),
                                 _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["nNeighbourSK"]),
                                 _i(geo["img2CentCell"]), _i(geo["iPair"]), _i(it),
                                 0 if it is None else len(it))
    return qq


DELTA_X_DIFF = float(np.finfo(np.float64).eps ** 0.25)  # nonscc.F90:43


def first_deriv_block(model, geo, which, atom_i, atom_j, diff_type=1, delta=DELTA_X_DIFF):
    """TNonSccDiff%getFirstDeriv for one diatomic block; returns deriv[3, mOrb(col), mOrb(row)]."""
    out = np.zeros((3, model.mOrb, model.mOrb))
    lib().orc_first_deriv_block(model._h, which, _d(out), _d(geo["coords"]), _i(geo["species"]),
                                atom_i, atom_j, diff_type, ctypes.c_double(delta))
    return out


def derivative_shift(model, geo, dm, edm, shift=None, diff_type=1, delta=DELTA_X_DIFF, iter_indices=None):
    """derivative_shift generic (forces.F90:27-38): dm[nSpin, sparseSize] (or 1-D),
    shift[nSpin, nAtom, mOrb, mOrb] or None.  Returns deriv[nAtom, 3]."""
    dm = np.ascontiguousarray(np.atleast_2d(dm), dtype=np.float64)
    nSpin = dm.shape[0]
    edm = np.ascontiguousarray(edm, dtype=np.float64)
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    it = None if iter_indices is None else np.ascontiguousarray(iter_indices, dtype=np.int32)
    deriv = np.zeros((geo["nAtom"], 3))
    lib().orc_derivative_shift(_d(deriv), model._h, _d(dm), nSpin, _d(edm), geo["sparseSize"], _d(sh),
                               geo["nAtom"], _d(geo["coords"]), _i(geo["species"]),
                               _i(geo["iNeighbour"]), geo["mNeigh"], _i(geo["nNeighbourSK"]),
                               _i(geo["img2CentCell"]), _i(geo["iPair"]), diff_type,
                               ctypes.c_double(delta), _i(it), 0 if it is None else len(it))
    return deriv


def stress(model, geo, dm, edm, cell_vol, shift=None, diff_type=1, delta=DELTA_X_DIFF, iter_indices=None):
    """getNonSCCStress / getBlockStress (dftb/stress.F90:70-326).  Returns st[3, 3] with
    st[ii, jj] = reference st(jj, ii) (numpy sees the column-major array transposed)."""
    dm = np.ascontiguousarray(np.atleast_2d(dm), dtype=np.float64)
    edm = np.ascontiguousarray(edm, dtype=np.float64)
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    it = None if iter_indices is None else np.ascontiguousarray(iter_indices, dtype=np.int32)
    st = np.zeros((3, 3))
    lib().orc_stress(_d(st), model._h, _d(dm), dm.shape[0], _d(edm), geo["sparseSize"], _d(sh), geo["nAtom"],
                     _d(geo["coords"]), _i(geo["species"]), _i(geo["iNeighbour"]), geo["mNeigh"],
                     _i(geo["nNeighbourSK"]), _i(geo["img2CentCell"]), _i(geo["iPair"]), diff_type,
                     ctypes.c_double(delta), ctypes.c_double(cell_vol), _i(it), 0 if it is None else len(it))
    return st

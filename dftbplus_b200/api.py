"""Host-side mirror of the reference's hot-path routines, on top of the C ABI.

The method names and argument meaning follow the Fortran module procedures they stand in for
(buildH0 / buildS: dftb/nonscc.F90; unpackHS / packHS: dftb/sparse2dense.F90; getDensityMatrix /
getEDensityMatrix: dftb/densitymatrix.F90; mulliken: dftb/populations.F90; derivative_shift:
dftb/forces.F90; diagDenseMtx: dftbplus/eigenvects.F90), so the parity tests read like calls
into the reference.  Arrays: numpy (host; copied by the library) or torch CUDA tensors (device;
used in place).  Layout conventions are those of include/dftbp_b200.h: memory is the Fortran
layout, numpy/torch see it C-ordered with reversed dimensions; dense matrices are column-major
(numpy order='F', torch tensors with strides (1, ld) -- see `new_dense`).
"""
import ctypes

import numpy as np

from . import _lib


class DftbB200Error(RuntimeError):
    """Non-zero status from libdftbp_b200 (the Fortran shim raises TStatus / calls error())."""

    def __init__(self, code, msg):
        super().__init__("libdftbp_b200 error %d: %s" % (code, msg))
        self.code = code


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return ctypes.c_void_p(x.data_ptr())
    return ctypes.c_void_p(x.ctypes.data)


def _ld(mat):
    """leading dimension of a column-major dense matrix (numpy F-order or torch strides (1, ld))."""
    if _is_torch(mat):
        assert mat.dim() == 2 and (mat.stride(0) == 1 or mat.shape[0] == 1), "dense matrix must be column-major"
        return max(int(mat.stride(1)), int(mat.shape[0])) if mat.shape[1] > 1 else int(mat.shape[0])
    assert mat.flags.f_contiguous, "dense matrix must be order='F'"
    return mat.shape[0]


def new_dense(n, dtype=None, device=None):
    """Column-major n x n torch matrix on the GPU, indexed [row, col]."""
    import torch
    return torch.zeros((n, n), dtype=dtype or torch.float64, device=device or "cuda").t()


def _f64(a):
    return a if _is_torch(a) else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return a if _is_torch(a) else np.ascontiguousarray(a, dtype=np.int32)


class HotPath:
    """One library context (= one TDftbPlus instance on one GPU)."""

    def __init__(self, device=0, rank=0, nranks=1):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.dftbp_b200_create(ctypes.byref(h), device, rank, nranks)
        self._h = h
        if rc != 0:
            msg = self._L.dftbp_b200_last_error(h).decode() if h else "create failed"
            if h:
                self._L.dftbp_b200_destroy(h)
            self._h = None
            raise DftbB200Error(rc, msg)
        self.device, self.rank, self.nranks = device, rank, nranks
        self.geo = None
        self.mOrb = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.dftbp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DftbB200Error(rc, self._L.dftbp_b200_last_error(self._h).decode())

    # ---- plumbing ------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._ck(self._L.dftbp_b200_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def use_torch_stream(self):
        """Enqueue on torch's current stream (needed whenever torch tensors are passed as device
        pointers: the producer/consumer kernels of those tensors live on that stream)."""
        import torch
        s = torch.cuda.current_stream().cuda_stream
        self.set_stream(s if s != 0 else 1)      # 0 = legacy default stream -> cudaStreamLegacy

    TIMERS = ("sparseH0S", "sparseToDense", "denseToSparse", "densityMatrix", "mulliken", "forceCalc",
              "diagonalization", "addShift", "neighbourList")

    def timer_enable(self, on=True):
        self._ck(self._L.dftbp_b200_timer_enable(self._h, 1 if on else 0))

    def timer_read(self):
        """{name: (total device ms, calls)} with the reference's timer names (environment.F90:96-134)."""
        out = {}
        for i, name in enumerate(self.TIMERS):
            ms, n = ctypes.c_double(), ctypes.c_longlong()
            self._ck(self._L.dftbp_b200_timer_read(self._h, i, ctypes.byref(ms), ctypes.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def synchronize(self):
        self._ck(self._L.dftbp_b200_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._L.dftbp_b200_kernel_launches(self._h))

    @property
    def num_pairs(self):
        return int(self._L.dftbp_b200_num_pairs(self._h))

    # ---- once per run ---------------------------------------------------------------------
    def set_orbitals(self, ang_shells):
        """TOrbitals: ang_shells[iSp] = list of shell angular momenta of species iSp+1."""
        nsp = len(ang_shells)
        msh = max(len(a) for a in ang_shells)
        n_shell = np.array([len(a) for a in ang_shells], dtype=np.int32)
        ang = np.zeros((nsp, msh), dtype=np.int32)
        for i, a in enumerate(ang_shells):
            ang[i, :len(a)] = a
        self._ck(self._L.dftbp_b200_set_orbitals(self._h, nsp, msh, _ptr(n_shell), _ptr(ang)))
        self.ang_shells = [list(a) for a in ang_shells]
        self.nOrbSpecies = np.array([sum(2 * l + 1 for l in a) for a in ang_shells], dtype=np.int32)
        self.mOrb = int(self.nOrbSpecies.max())
        self.mShell = msh

    def set_sk_table(self, which, sp1, sp2, dist, tab, method=2):
        """SlakoCont addTable: tab[nGrid, nInteg] (order='F'), which 0 = H, 1 = S."""
        tab = np.asfortranarray(tab, dtype=np.float64)
        self._ck(self._L.dftbp_b200_set_sk_table(self._h, which, sp1, sp2, tab.shape[0], tab.shape[1],
                                                 float(dist), _ptr(tab), method))

    def sk_cutoff(self):
        c = ctypes.c_double()
        self._ck(self._L.dftbp_b200_get_sk_cutoff(self._h, ctypes.byref(c)))
        return c.value

    # ---- once per geometry ------------------------------------------------------------------
    def set_geometry(self, geo):
        """geo: dict with the arrays of SURVEY.md section 8.0 (see dftbplus_b200.geometry)."""
        g = geo
        self._ck(self._L.dftbp_b200_set_geometry(
            self._h, g["nAtom"], g["nAllAtom"], g["mNeigh"], _ptr(_f64(g["coords"])), _ptr(_i32(g["species"])),
            _ptr(_i32(g["img2CentCell"])), _ptr(_i32(g["iCellVec"])), g["cellVec"].shape[0],
            _ptr(_f64(g["cellVec"])), _ptr(_i32(g["nNeighbourSK"])), _ptr(_i32(g["iNeighbour"])),
            _ptr(_i32(g["iPair"])), _ptr(_i32(g["iAtomStart"])), g["sparseSize"]))
        self.geo = g

    def getCellTranslations(self, latvec, cutoff):
        """getCellTranslations (dftb/periodic.F90:227-263).  latvec[i] = i-th lattice vector (rows; the
        memory is Fortran's latVec(:, i)).  Returns (cellVec[nCellVec, 3], rCellVec[nCellVec, 3])."""
        latvec = np.ascontiguousarray(latvec, dtype=np.float64)
        rec = np.ascontiguousarray(np.linalg.inv(latvec).T)   # recVec2p(:, i), initprogram.F90 (invLatVec^T)
        n = ctypes.c_int()
        rc = self._L.dftbp_b200_cell_translations(_ptr(latvec), _ptr(rec), float(cutoff), 0, ctypes.byref(n), None, None)
        if rc != 0:
            raise DftbB200Error(rc, "cell_translations: bad arguments")
        cell, rcell = np.zeros((n.value, 3)), np.zeros((n.value, 3))
        rc = self._L.dftbp_b200_cell_translations(_ptr(latvec), _ptr(rec), float(cutoff), n.value, ctypes.byref(n),
                                                  _ptr(cell), _ptr(rcell))
        if rc != 0:
            raise DftbB200Error(rc, "cell_translations: bad arguments")
        return cell, rcell

    def foldCoordsToCell(self, coord, latvec):
        """boundaryCond%foldCoordsToCell (pbc3d): coord[nAtom, 3] (numpy: a folded copy is returned;
        CUDA tensor: folded in place), latvec[i] = i-th lattice vector."""
        latvec = np.ascontiguousarray(latvec, dtype=np.float64)
        if not _is_torch(coord):
            coord = np.array(coord, dtype=np.float64, order="C", copy=True)
        self._ck(self._L.dftbp_b200_fold_coords_to_cell(self._h, int(coord.shape[0]), _ptr(coord), _ptr(latvec)))
        return coord

    def updateNeighbourListAndSpecies(self, coord0, species0, cutoff, latvec=None, sk_cutoff=None, adopt=True,
                                      fetch=True, symmetric=False):
        """updateNeighbourListAndSpecies + getNrOfNeighboursForAll + getSparseDescriptor on the device
        (handleCoordinateChange, dftbplus/main.F90:2562-2578).  coord0[nAtom, 3] (host array or CUDA
        tensor), species0[nAtom] 1-based, cutoff = list cutoff, sk_cutoff (default: cutoff).
        adopt: make the result the context's geometry (the caller's arrays are not involved; on one rank the
        work list and block descriptors are built by kernels, the host reads back nAtom + 2 ints); fetch: return the
        geometry dict of SURVEY.md section 8.0 (host copies), else only the sizes.  symmetric: the reference's
        symmetric list (for the host's use; never adopted)."""
        sk_cutoff = cutoff if sk_cutoff is None else sk_cutoff
        if latvec is None:
            cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
        else:
            cell, rcell = self.getCellTranslations(latvec, cutoff)
        coord0, species0 = _f64(coord0), _i32(species0)
        nAtom = int(coord0.shape[0])
        nAll, mN, sparse = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._ck(self._L.dftbp_b200_update_neighbour_list(
            self._h, nAtom, _ptr(coord0), _ptr(species0), cell.shape[0], _ptr(cell), _ptr(rcell), float(cutoff),
            float(sk_cutoff), 1 if symmetric else 0, ctypes.byref(nAll), ctypes.byref(mN), ctypes.byref(sparse)))
        g = dict(nAtom=nAtom, nAllAtom=nAll.value, mNeigh=mN.value, sparseSize=sparse.value, cellVec=cell,
                 rCellVec=rcell, cutoff=cutoff, latvec=latvec)
        if fetch:
            g.update(coords=np.zeros((g["nAllAtom"], 3)), species=np.zeros(g["nAllAtom"], np.int32),
                     img2CentCell=np.zeros(g["nAllAtom"], np.int32), iCellVec=np.zeros(g["nAllAtom"], np.int32),
                     nNeighbour=np.zeros(nAtom, np.int32), iNeighbour=np.zeros((nAtom, g["mNeigh"] + 1), np.int32),
                     neighDist2=np.zeros((nAtom, g["mNeigh"] + 1)), nNeighbourSK=np.zeros(nAtom, np.int32),
                     iPair=np.zeros((nAtom, g["mNeigh"] + 1), np.int32), iAtomStart=np.zeros(nAtom + 1, np.int32))
            self._ck(self._L.dftbp_b200_get_neighbour_list(
                self._h, _ptr(g["coords"]), _ptr(g["species"]), _ptr(g["img2CentCell"]), _ptr(g["iCellVec"]),
                _ptr(g["nNeighbour"]), _ptr(g["iNeighbour"]), _ptr(g["neighDist2"]), _ptr(g["nNeighbourSK"]),
                _ptr(g["iPair"]), _ptr(g["iAtomStart"])))
            g["nOrb"] = int(g["iAtomStart"][-1] - 1)
            g["species0"] = np.asarray(g["species"][:nAtom])
            g["nOrbAtom"] = np.diff(g["iAtomStart"]).astype(np.int32)
        if not fetch:
            sp = species0.cpu().numpy() if _is_torch(species0) else species0
            g["nOrb"] = int(self.nOrbSpecies[np.asarray(sp) - 1].sum())
        self._last_list = g
        if adopt and not symmetric:
            self._ck(self._L.dftbp_b200_adopt_neighbour_list(self._h))
            self.geo = g        # without fetch: the sizes only (enough for the default output buffers)
        return g

    def getNrOfNeighboursForAll(self, cutoff, n_atom=None):
        """getNrOfNeighboursForAll (dftb/periodic.F90:996-1054) on the list of the last update."""
        out = np.zeros(self.geo["nAtom"] if n_atom is None else n_atom, dtype=np.int32)
        self._ck(self._L.dftbp_b200_nr_of_neighbours(self._h, float(cutoff), _ptr(out)))
        return out

    def update_coords(self, coords):
        self._ck(self._L.dftbp_b200_update_coords(self._h, _ptr(_f64(coords))))

    # ---- K1 ---------------------------------------------------------------------------------
    def _selfegy(self, selfegy):
        se = np.zeros((len(self.ang_shells), self.mShell))
        for i, row in enumerate(selfegy):
            se[i, :len(row)] = row
        return se

    def _sparse_out(self, out):
        return np.zeros(self.geo["sparseSize"]) if out is None else out

    def buildH0(self, selfegy, ham=None):
        ham = self._sparse_out(ham)
        se = self._selfegy(selfegy)
        self._ck(self._L.dftbp_b200_build_h0(self._h, _ptr(se), _ptr(ham)))
        return ham

    def buildS(self, over=None):
        over = self._sparse_out(over)
        self._ck(self._L.dftbp_b200_build_s(self._h, _ptr(over)))
        return over

    def buildH0S(self, selfegy, ham=None, over=None):
        ham, over = self._sparse_out(ham), self._sparse_out(over)
        se = self._selfegy(selfegy)
        self._ck(self._L.dftbp_b200_build_h0_s(self._h, _ptr(se), _ptr(ham), _ptr(over)))
        return ham, over

    # ---- K2 ---------------------------------------------------------------------------------
    def unpackHS(self, orig, kpoint=None, square=None):
        n = self.geo["nOrb"]
        if kpoint is None:
            if square is None:
                square = np.zeros((n, n), order="F")
            self._ck(self._L.dftbp_b200_unpack_hs_real(self._h, _ptr(_f64(orig)), _ptr(square), _ld(square)))
        else:
            if square is None:
                square = np.zeros((n, n), order="F", dtype=np.complex128)
            kp = (ctypes.c_double * 3)(*[float(x) for x in kpoint])
            self._ck(self._L.dftbp_b200_unpack_hs_cplx(self._h, _ptr(_f64(orig)), kp, _ptr(square), _ld(square)))
        return square

    def packHS(self, primitive, square, kpoint=None, kweight=1.0, atoms=None):
        """Accumulates into primitive (in place) like the reference.  atoms = (atomFirst, atomLast),
        1-based inclusive: only the blocks of those atoms (a column panel, schedule.dm_panels)."""
        a0, a1 = (1, self.geo["nAtom"]) if atoms is None else (int(atoms[0]), int(atoms[1]))
        if kpoint is None:
            self._ck(self._L.dftbp_b200_pack_hs_real_atoms(self._h, _ptr(primitive), _ptr(square), _ld(square), a0, a1))
        else:
            kp = (ctypes.c_double * 3)(*[float(x) for x in kpoint])
            self._ck(self._L.dftbp_b200_pack_hs_cplx_atoms(self._h, _ptr(primitive), _ptr(square), _ld(square), kp,
                                                           float(kweight), a0, a1))
        return primitive

    # ---- K3 ---------------------------------------------------------------------------------
    def _dm(self, eigvecs, filling, weights, dm, cols=None):
        cplx = (str(eigvecs.dtype).endswith("complex128"))
        n = eigvecs.shape[0]
        if dm is None:
            if _is_torch(eigvecs):
                dm = new_dense(n, dtype=eigvecs.dtype, device=eigvecs.device)
            else:
                dm = np.zeros((n, n), order="F", dtype=eigvecs.dtype)
        if cols is None:
            fn = self._L.dftbp_b200_density_matrix_cplx if cplx else self._L.dftbp_b200_density_matrix_real
            self._ck(fn(self._h, _ptr(dm), _ptr(eigvecs), n, _ld(eigvecs), _ptr(_f64(filling)),
                        _ptr(None if weights is None else _f64(weights))))
        else:
            fn = self._L.dftbp_b200_density_matrix_cplx_cols if cplx else self._L.dftbp_b200_density_matrix_real_cols
            self._ck(fn(self._h, _ptr(dm), _ptr(eigvecs), n, _ld(eigvecs), _ptr(_f64(filling)),
                        _ptr(None if weights is None else _f64(weights)), int(cols[0]), int(cols[1])))
        return dm

    def getDensityMatrix(self, eigvecs, filling, dm=None, cols=None):
        """TDensityMatrix%getDensityMatrix (full, real or complex by dtype).  cols = (colFirst, colLast),
        1-based inclusive: only that column panel of the lower triangle (schedule.dm_panels)."""
        return self._dm(eigvecs, filling, None, dm, cols)

    def getEDensityMatrix(self, eigvecs, filling, eigen, dm=None, cols=None):
        """TDensityMatrix%getEDensityMatrix: weights filling*eigen, nLevels from filling."""
        w = filling * eigen
        return self._dm(eigvecs, filling, w, dm, cols)

    # ---- K4 ---------------------------------------------------------------------------------
    def mulliken(self, over, rho, qq=None, block=False):
        nA = self.geo["nAtom"]
        if qq is None:
            qq = np.zeros((nA, self.mOrb, self.mOrb) if block else (nA, self.mOrb))
        fn = self._L.dftbp_b200_mulliken_block if block else self._L.dftbp_b200_mulliken_orbital
        self._ck(fn(self._h, _ptr(qq), _ptr(_f64(over)), _ptr(_f64(rho))))
        return qq

    def mullikenPerAtom(self, over, rho, qq=None):
        """mullikenPerAtom (dftb/populations.F90:68-112): qq[nAtom] accumulates."""
        if qq is None:
            qq = np.zeros(self.geo["nAtom"])
        self._ck(self._L.dftbp_b200_mulliken_atom(self._h, _ptr(qq), _ptr(_f64(over)), _ptr(_f64(rho))))
        return qq

    def getOnsitePopulation(self, rho, qq=None):
        """getOnsitePopulation (dftb/populations.F90:1195-1222): qq[nAtom] = trace of the on-site blocks."""
        if qq is None:
            qq = np.zeros(self.geo["nAtom"])
        self._ck(self._L.dftbp_b200_onsite_population(self._h, _ptr(_f64(rho)), _ptr(qq)))
        return qq

    # ---- K6 ---------------------------------------------------------------------------------
    def addShift(self, ham, over, shift, is_input_zero=False):
        """addShift (dftb/shift.F90:41-228), in place on ham[nSpin, sparseSize] (or 1-D for one spin):
        shift[nSpin, nAtom] -> addShift_atom, shift[nSpin, nAtom, mOrb, mOrb] -> addShift_block."""
        n_spin = ham.shape[0] if ham.ndim == 2 else 1
        block = shift.ndim == 4
        fn = self._L.dftbp_b200_add_shift_block if block else self._L.dftbp_b200_add_shift_atom
        self._ck(fn(self._h, _ptr(ham), _ptr(_f64(over)), _ptr(_f64(shift)), n_spin, 1 if is_input_zero else 0))
        return ham

    # ---- K5 ---------------------------------------------------------------------------------
    def set_pair_partition(self, part, n_parts):
        """Subsequent derivative_shift / stress calls take share `part` of n_parts of the pairs (a contiguous
        range of the neighbour list, dftbp_b200_set_pair_partition); the partial results add up to the
        full one, also when the shares are taken by different contexts / GPUs."""
        self._ck(self._L.dftbp_b200_set_pair_partition(self._h, int(part), int(n_parts)))

    def derivative_shift(self, dm, edm, shift=None, diff_type=1, delta=0.0, deriv=None):
        """dm[nSpin, sparseSize] (or 1-D), edm[sparseSize], shift[nSpin, nAtom, mOrb, mOrb] or None."""
        if not _is_torch(dm):
            dm = np.ascontiguousarray(np.atleast_2d(dm), dtype=np.float64)
        n_spin = dm.shape[0] if dm.ndim == 2 else 1
        if deriv is None:
            deriv = np.zeros((self.geo["nAtom"], 3))
        self._ck(self._L.dftbp_b200_derivative_shift(
            self._h, _ptr(deriv), _ptr(dm), n_spin, _ptr(_f64(edm)),
            _ptr(None if shift is None else _f64(shift)), diff_type, float(delta)))
        return deriv

    def stress(self, dm, edm, cell_vol, shift=None, diff_type=1, delta=0.0, st=None):
        """getNonSCCStress / getBlockStress (dftb/stress.F90:70-326); returns st[3, 3] with
        st[ii, jj] = reference st(jj, ii)."""
        if not _is_torch(dm):
            dm = np.ascontiguousarray(np.atleast_2d(dm), dtype=np.float64)
        n_spin = dm.shape[0] if dm.ndim == 2 else 1
        if st is None:
            st = np.zeros((3, 3))
        self._ck(self._L.dftbp_b200_stress(
            self._h, _ptr(st), _ptr(dm), n_spin, _ptr(_f64(edm)),
            _ptr(None if shift is None else _f64(shift)), diff_type, float(delta), float(cell_vol)))
        return st

    # ---- cuSOLVER ---------------------------------------------------------------------------
    def Efilling(self, eigenvals, n_electrons, kT, k_weight=(1.0,), filling=None):
        """Efilling (dftb/etemp.F90:59-191, Fermi distribution) on the device.  eigenvals [nSpin][nKpt][nLevels]
        in memory (= the Fortran array (nLevels, nKpt, nSpin)); lower-dimensional inputs are taken as one
        k-point / one spin channel.  One electron per level: callers of spin-unpolarised runs pass half the
        electrons and double the results (getFillingsAndBandEnergies, main.F90:4597-4648).
        Returns dict(filling, Ef, Ebs, TS, E0); with torch inputs the filling stays on the device."""
        tor = _is_torch(eigenvals)
        shape = tuple(eigenvals.shape)
        n_lev = int(shape[-1])
        n_k = int(shape[-2]) if len(shape) >= 2 else 1
        n_spin = int(shape[-3]) if len(shape) >= 3 else 1
        if tor:
            import torch
            eig = eigenvals.contiguous()
            if filling is None:
                filling = torch.empty_like(eig)
            kw = torch.as_tensor(np.asarray(k_weight, dtype=np.float64)).to(eig.device) if not _is_torch(k_weight) else k_weight
        else:
            eig = np.ascontiguousarray(eigenvals, dtype=np.float64)
            if filling is None:
                filling = np.empty_like(eig)
            kw = np.ascontiguousarray(k_weight, dtype=np.float64)
        small = np.zeros(7)        # Ebs(2), TS(2), E0(2), Ef: host outputs
        self._ck(self._L.dftbp_b200_efilling(
            self._h, _ptr(small[0:2]), _ptr(small[6:7]), _ptr(small[2:4]), _ptr(small[4:6]), _ptr(filling), _ptr(eig),
            n_lev, n_k, n_spin, float(n_electrons), float(kT), _ptr(kw), 0))
        return dict(filling=filling, Ef=float(small[6]), Ebs=small[0:n_spin].copy(), TS=small[2:2 + n_spin].copy(),
                    E0=small[4:4 + n_spin].copy())

    # ---- N4: Coulomb 1/R matrix and shifts -------------------------------------------------------
    def invRCluster(self, coord, invr=None):
        """invRCluster (dftb/coulomb.F90:623-733): lower triangle of 1 / |r_i - r_j|, column-major [n, n]."""
        n = int(coord.shape[0])
        if invr is None:
            invr = np.zeros((n, n), order="F")
        self._ck(self._L.dftbp_b200_invr_cluster(self._h, n, _ptr(_f64(coord)), _ptr(invr), _ld(invr)))
        return invr

    def invRPeriodic(self, coord, r_lat, g_lat, alpha, volume, r_cutoff, invr=None):
        """invRPeriodic (dftb/coulomb.F90:815-1024): the Ewald 1/R matrix.  r_lat [nRLat, 3] lattice translations
        (absolute, origin first, getCellTranslations order), g_lat [nG, 3] reciprocal points (absolute, no origin,
        no inversion pairs) -- what TCoulomb%updateLatVecs keeps as rLatPoints_ / gLatPoints_."""
        n = int(coord.shape[0])
        if invr is None:
            invr = np.zeros((n, n), order="F")
        self._ck(self._L.dftbp_b200_invr_periodic(self._h, n, _ptr(_f64(coord)), int(r_lat.shape[0]), _ptr(_f64(r_lat)),
                                                  int(g_lat.shape[0]), _ptr(_f64(g_lat)), float(alpha), float(volume),
                                                  float(r_cutoff), _ptr(invr), _ld(invr)))
        return invr

    def coulombShift(self, invr, dq, shift=None):
        """TCoulomb%updateShifts (dftb/coulomb.F90:537-586): shiftPerAtom = hemv(invRMat, deltaQAtom, 'L')."""
        n = int(dq.shape[0])
        if shift is None:
            if _is_torch(dq):
                import torch
                shift = torch.zeros_like(dq)
            else:
                shift = np.zeros(n)
        self._ck(self._L.dftbp_b200_coulomb_shift(self._h, n, _ptr(invr), _ld(invr), _ptr(_f64(dq)), _ptr(shift)))
        return shift

    def shortGammaUpdate(self, n_hubb_u, uniq_hubb_u, short_cutoffs, fetch=False):
        """TShortGamma%updateCoords (dftb/shortgamma.F90:177-196) on the device neighbour list of the last
        updateNeighbourListAndSpecies.  n_hubb_u [nSpecies], uniq_hubb_u [nSpecies, mU], short_cutoffs
        [nSpecies(sp1), nSpecies(sp2), mU(iU1), mU(iU2)] = the Fortran arrays (mU, nSpecies) / (iU2, iU1, iSp2, iSp1).
        fetch: return dict(nNeighU [nAtom, nSpecies, mU, mU], maxNeigh [nAtom], shortGamma [nAtom, mNeigh + 1, mU, mU])."""
        nh = _i32(n_hubb_u)
        uq = np.ascontiguousarray(uniq_hubb_u, dtype=np.float64)
        sc = np.ascontiguousarray(short_cutoffs, dtype=np.float64)
        n_sp, m_u = int(uq.shape[0]), int(uq.shape[1])
        self._ck(self._L.dftbp_b200_short_gamma_update(self._h, n_sp, m_u, _ptr(nh), _ptr(uq), _ptr(sc)))
        if not fetch:
            return None
        n, ld = self._last_list["nAtom"], self._last_list["mNeigh"] + 1
        out = dict(nNeighU=np.zeros((n, n_sp, m_u, m_u), dtype=np.int32), maxNeigh=np.zeros(n, dtype=np.int32),
                   shortGamma=np.zeros((n, ld, m_u, m_u)))
        self._ck(self._L.dftbp_b200_short_gamma_get(self._h, _ptr(out["nNeighU"]), _ptr(out["maxNeigh"]), _ptr(out["shortGamma"])))
        return out

    def shortGammaShifts(self, i_hubb_u, dq_shell, shift_shell=None):
        """TShortGamma%updateShifts (dftb/shortgamma.F90:220-244, buildShifts_ :860-933): dq_shell [nAtom, mShell]
        -> shiftShell [nAtom, mShell]; i_hubb_u [nSpecies, mShell] (1-based, 0 = no shell)."""
        dq = _f64(dq_shell)
        if shift_shell is None:
            shift_shell = np.zeros(tuple(dq.shape))
        self._ck(self._L.dftbp_b200_short_gamma_shifts(self._h, _ptr(_i32(i_hubb_u)), _ptr(dq), _ptr(shift_shell)))
        return shift_shell

    # ---- N4: Broyden mixer with device-resident vectors ------------------------------------------------
    def mixerInit(self, max_scc_iter=100, mix_param=0.2, omega0=0.01, min_weight=1.0, max_weight=1.0e5,
                  weight_fac=1.0e-2):
        """TBroydenMixerReal_init (mixer/broydenmixer.F90:110-139) with the parser's defaults."""
        self._ck(self._L.dftbp_b200_mixer_init(self._h, int(max_scc_iter), float(mix_param), float(omega0),
                                               float(min_weight), float(max_weight), float(weight_fac)))

    def mixerReset(self, n_elem):
        """TBroydenMixerReal_reset (:143-170): a new SCC cycle over vectors of n_elem charges."""
        self._ck(self._L.dftbp_b200_mixer_reset(self._h, int(n_elem)))

    def mixerMix(self, q_inp_result, q_diff):
        """TBroydenMixerReal_mix (:180-342): q_inp_result is overwritten with the mixed charges (in place)."""
        self._ck(self._L.dftbp_b200_mixer_mix(self._h, _ptr(q_inp_result), _ptr(_f64(q_diff)), int(q_inp_result.shape[0])))
        return q_inp_result

    def diagDenseMtx(self, H, S, eigen=None, jobz="V"):
        """hegvd: H, S column-major, overwritten (H <- eigenvectors); returns eigenvalues."""
        n = H.shape[0]
        cplx = str(H.dtype).endswith("complex128")
        if eigen is None:
            if _is_torch(H):
                import torch
                eigen = torch.zeros(n, dtype=torch.float64, device=H.device)
            else:
                eigen = np.zeros(n)
        fn = self._L.dftbp_b200_hegv_cplx if cplx else self._L.dftbp_b200_hegv_real
        self._ck(fn(self._h, _ptr(H), _ptr(S), _ptr(eigen), n, _ld(H), jobz.encode()[0:1]))
        return eigen

#!/usr/bin/env python3
"""Golden vectors of the neighbour list made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
updateNeighbourList (dftb/periodic.F90:446-694), with index_heap_sort = the machine translation of gen_ref_math.py.

TEST INFRASTRUCTURE ONLY.  Hand-mapped glue (not part of the algorithm): the array growth helpers reallocateArrays1 / 2
(copy and zero-pad), incrmntOfArray (common/memman.F90:70-78), allocateNeighbourArrays / fillNeighbourArrays in their
non-MPI branch (row 0 of every atom = the atom itself, rows 1.. = the sorted neighbours), the helical rotation (never
reached).  Outputs -> tests/golden/neigh_ref_golden.npz; tests/test_oracle_neigh_ref.py compares the oracle's
orc_update_neighbour_list with them."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402
import gen_ref_math  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/periodic.F90"
REF_S2D = "/root/reference/src/dftbp/dftb/sparse2dense.F90"


class Neigh:
    pass


def _resize(arr, shape):
    new = np.zeros(shape, dtype=arr.a.dtype, order="F")
    sl = tuple(slice(0, min(a, b)) for a, b in zip(arr.a.shape, shape))
    new[sl] = arr.a[sl]
    arr.a = new


def routine():
    mref = gen_ref_math.load(os.path.join(HERE, "_ref", "libmathref.so"))

    def realloc1(img, icv, coord, m):
        _resize(img, (m,)); _resize(icv, (m,)); _resize(coord, (3, m))

    def realloc2(inb, nd2, m):
        _resize(inb, (m, inb.a.shape[1])); _resize(nd2, (m, nd2.a.shape[1]))

    def heap_sort(indx, arr, tol):
        n = len(indx)
        if n == 0:
            return
        idx = np.zeros(n, dtype=np.int32)
        vals = np.ascontiguousarray(arr, dtype=np.float64)
        mref.mref_index_heap_sort(n, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                  vals.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), tol)
        indx[:] = idx

    def alloc_arrays(neigh, max_neigh, n_atom, env):
        neigh.iNeighbour = fx.FArr(np.zeros((max_neigh + 1, n_atom), dtype=np.int64, order="F"), (0, 1))
        neigh.neighDist2 = fx.FArr(np.zeros((max_neigh + 1, n_atom), order="F"), (0, 1))

    def fill_arrays(neigh, inb, nd2, start, end, max_neigh, n_atom):
        neigh.iNeighbour.a[1:, :] = inb.a[:max_neigh, :]
        neigh.neighDist2.a[1:, :] = nd2.a[:max_neigh, :]
        neigh.iNeighbour.a[0, :] = np.arange(1, n_atom + 1)
        neigh.neighDist2.a[0, :] = 0.0

    join = lambda fn: (lambda a: "%s(%s)" % (fn, ", ".join(a)))  # noqa: E731
    calls = {"reallocateArrays1": join("reallocateArrays1"), "reallocateArrays2": join("reallocateArrays2"),
             "index_heap_sort": join("index_heap_sort"), "allocateNeighbourArrays": join("allocateNeighbourArrays"),
             "fillNeighbourArrays": join("fillNeighbourArrays"), "rotate3": "pass"}
    env = dict(reallocateArrays1=realloc1, reallocateArrays2=realloc2, index_heap_sort=heap_sort,
               allocateNeighbourArrays=alloc_arrays, fillNeighbourArrays=fill_arrays,
               incrmntOfArray=lambda m: m + m // 2 + 1, minNeighDist2=1.0e-2 * 1.0e-2, tolSameDist2=1.0e-5 * 1.0e-5, zAxis=None)
    return fx.compile_routine(open(REF).read(), "subroutine", "updateNeighbourList", calls, extra_env=env,
                              extra_arrays=("nNeighbour", "iNeighbour", "neighDist2"))[0]


def run(f, coord0, rcell, cutoff, symmetric=False):
    """coord0 [nAtom, 3], rcell [nCellVec, 3] -> dict like oracle.neighbour_list_full (or error pair)."""
    n = coord0.shape[0]
    F = np.asfortranarray
    coord = fx.FArr(np.zeros((3, n), order="F"))
    img = fx.FArr(np.zeros(n, dtype=np.int64))
    icv = fx.FArr(np.zeros(n, dtype=np.int64))
    neigh = Neigh()
    neigh.nNeighbour = fx.FArr(np.zeros(n, dtype=np.int64))
    loc = f(coord, img, icv, neigh, 0, fx.FArr(F(np.asarray(coord0, dtype=np.float64).T)), float(cutoff),
            fx.FArr(F(np.asarray(rcell, dtype=np.float64).T)), None, None, symmetric, None)
    err = loc["pairError"].a
    if err[0] != 0:
        return dict(error=(int(err[0]), int(err[1])))
    nall = int(loc["nAllAtom"])
    return dict(nAllAtom=nall, mNeigh=int(neigh.iNeighbour.a.shape[0] - 1), coords=np.ascontiguousarray(coord.a.T[:nall]),
                img2CentCell=img.a[:nall].astype(np.int32), iCellVec=icv.a[:nall].astype(np.int32),
                nNeighbour=neigh.nNeighbour.a.astype(np.int32), iNeighbour=np.ascontiguousarray(neigh.iNeighbour.a.T).astype(np.int32),
                neighDist2=np.ascontiguousarray(neigh.neighDist2.a.T))


def sparse_descriptor(r, n_orb_atom):
    """getSparseDescriptor (dftb/sparse2dense.F90:3617-3662) executed from its source on the list r."""
    f = fx.compile_routine(open(REF_S2D).read(), "subroutine", "getSparseDescriptor", {}, extra_arrays=("nOrbAtom",))[0]
    n = len(n_orb_atom)
    orb = Neigh()
    orb.nOrbAtom = fx.FArr(np.asarray(n_orb_atom, dtype=np.int64))
    ipair = fx.FArr(np.zeros((1, n), dtype=np.int64, order="F"), (0, 1))       # too small on purpose: the routine reallocates
    loc = f(fx.wrap(f, "iNeighbour", np.asfortranarray(r["iNeighbour"].T.astype(np.int64))),
            fx.wrap(f, "nNeighbour", r["nNeighbour"].astype(np.int64)), fx.wrap(f, "img2CentCell", r["img2CentCell"].astype(np.int64)),
            orb, ipair, 0)
    return np.ascontiguousarray(loc["iPair"].a.T).astype(np.int32), int(loc["sparseSize"])


def cases():
    """name -> (coord0, rcell, cutoff, symmetric); inputs are regenerated by the test from the same code."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import case, oracle
    out = {}
    for name in ("h2o", "ce_o", "si2", "fec128"):
        w, sk, geo, model, selfegy = case(name)
        cutoff = sk.cutoff()
        rcell = np.zeros((1, 3)) if w["latvec"] is None else oracle.cell_translations(w["latvec"], cutoff)[1]
        out[name] = (np.asarray(w["coord0"], dtype=np.float64), rcell, cutoff, False)
    # a perfect lattice: every shell holds distances equal within tolSameDist2 (the order is the heap sort's)
    a = 5.42
    basis = np.array([[0, 0, 0], [.5, .5, .5]])
    cc = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)])
    c0 = (cc[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    rcell = oracle.cell_translations(np.diag([2 * a, 2 * a, 2 * a]), 9.5)[1]
    out["crystal"] = (c0, rcell, 9.5, False)
    out["crystal_symmetric"] = (c0, oracle.cell_translations(np.diag([2 * a, 2 * a, 2 * a]), 7.0)[1], 7.0, True)
    w, sk, geo, model, selfegy = case("fec128")
    c1 = np.array(w["coord0"], dtype=np.float64)
    c1[77] = c1[12] + [0.003, 0.0, -0.002]
    c1[90] = c1[31] + [0.0, 0.004, 0.0]
    out["too_close"] = (c1, oracle.cell_translations(w["latvec"], sk.cutoff())[1], sk.cutoff(), False)
    return out


def orbitals_of(n):
    """orbitals per atom of the descriptor check: alternating s p d / s p / s atoms"""
    return np.array([(9, 4, 1)[i % 3] for i in range(n)], dtype=np.int64)


def main():
    if not os.path.exists(REF):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    f = routine()
    g = {}
    for name, (c0, rcell, cutoff, sym) in cases().items():
        r = run(f, c0, rcell, cutoff, sym)
        if "error" not in r and not sym:
            r["iPair"], r["sparseSize"] = sparse_descriptor(r, orbitals_of(len(c0)))
        for k, v in r.items():
            g["%s__%s" % (name, k)] = np.asarray(v)
        print(name, {k: (v if np.ndim(v) == 0 or k == "error" else np.shape(v)) for k, v in r.items()})
    path = os.path.join(ROOT, "tests", "golden", "neigh_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()

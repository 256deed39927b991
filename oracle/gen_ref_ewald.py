#!/usr/bin/env python3
"""Golden vectors of the Coulomb 1/R matrix made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
invRPeriodicSerial, addNeighbourContribsIS, ewaldReciprocal, invRClusterSerial (dftb/coulomb.F90) together with the dynamic
neighbour list they iterate over -- TNeighIterator_init, TNeighIterator_getNextNeighbours (type/dynneighlist.F90) and
TLatPointIter_init, TLatPointIter_getNextPoint, getNextPoint, getRanges_ (type/latpointiter.F90); rTerm = the machine
translation of gen_ref_scc.py.  Type-bound calls are mapped to the routines they bind to; components the types initialise
by default (posExt = negExt = 0, tAll, tExcludeOrig, tExcludeInv) are set by new_object().  TEST INFRASTRUCTURE ONLY.
Outputs -> tests/golden/ewald_ref_golden.npz; tests/test_oracle_ewald_ref.py compares orc_invr_periodic / orc_invr_cluster."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402
import gen_ref_scc  # noqa: E402

REF_C = "/root/reference/src/dftbp/dftb/coulomb.F90"
REF_D = "/root/reference/src/dftbp/type/dynneighlist.F90"
REF_L = "/root/reference/src/dftbp/type/latpointiter.F90"


class Obj:
    pass


def new_object(kind):
    o = Obj()
    if kind == "latPointGen":                      # TLatPointIter with its default component values
        o.posExt, o.negExt, o.tAll, o.tExcludeOrig, o.tExcludeInv = 0, 0, True, True, False
        o.latVecs = fx.FArr(np.zeros((3, 3), order="F"))
        o.ranges = fx.FArr(np.zeros((2, 3), dtype=np.int64, order="F"))
        o.curPoint = fx.FArr(np.zeros(3, dtype=np.int64))
        o.cutoff2 = 0.0
    if kind == "neighIter":                        # TNeighIterator: its fixed-size components
        o.cellVec = fx.FArr(np.zeros(3))
        o.coordsAtom1 = fx.FArr(np.zeros(3))
        o.latVecs = fx.FArr(np.zeros((3, 3), order="F"))
        o.tFinished = False
    return o


def routines():
    scc = gen_ref_scc.load(os.path.join(HERE, "_ref", "libsccref.so"))
    csrc, dsrc, lsrc = open(REF_C).read(), open(REF_D).read(), open(REF_L).read()
    join = lambda fn: (lambda a: "%s(%s)" % (fn, ", ".join(a)))  # noqa: E731
    arrays = ("coords0", "latVecs", "invLatVecs", "ranges", "curPoint", "cellVec", "coordsAtom1")
    env = dict(new_object=new_object, rTerm=lambda r, a: scc.rTerm(r, a), iterChunkSize_=1000)
    lcalls = {"getRanges_": join("getRanges_"), "getNextPoint": join("getNextPoint")}
    for name in ("getRanges_", "getNextPoint"):
        env[name] = fx.compile_routine(lsrc, "subroutine", name, lcalls, extra_env=env, extra_arrays=arrays)[0]
    for name in ("TLatPointIter_init", "TLatPointIter_getNextPoint"):
        env[name] = fx.compile_routine(lsrc, "subroutine", name, lcalls, extra_env=env, extra_arrays=arrays)[0]
    dcalls = {"TLatPointIter_init": join("TLatPointIter_init"),
              "this%latPointGen%getNextPoint": lambda a: "this.tFinished = TLatPointIter_getNextPoint(this.latPointGen, %s, None)['tFinished']" % a[0]}
    for name in ("TNeighIterator_init", "TNeighIterator_getNextNeighbours"):
        env[name] = fx.compile_routine(dsrc, "subroutine", name, dcalls, extra_env=env, extra_arrays=arrays)[0]
    ccalls = {"TNeighIterator_init": lambda a: "%s = new_object('neighIter'); TNeighIterator_init(%s)" % (a[0], ", ".join(a)),
              "neighIter%getNextNeighbours": lambda a: "%s = TNeighIterator_getNextNeighbours(neighIter, %s)['nNeighbourSK']" % (a[0], ", ".join(a)),
              "addNeighbourContribsIS": join("addNeighbourContribsIS")}
    env["ewaldReciprocal"] = None
    rec = fx.compile_routine(csrc, "function", "ewaldReciprocal", ccalls, extra_env=env, extra_arrays=arrays)[0]
    env["ewaldReciprocal"] = lambda *a: rec(*a)["recSum"]
    env["addNeighbourContribsIS"] = fx.compile_routine(csrc, "subroutine", "addNeighbourContribsIS", ccalls, extra_env=env,
                                                       extra_arrays=arrays)[0]
    per = fx.compile_routine(csrc, "subroutine", "invRPeriodicSerial", ccalls, extra_env=env, extra_arrays=arrays)[0]
    clu = fx.compile_routine(csrc, "subroutine", "invRClusterSerial", ccalls, extra_env=env, extra_arrays=arrays)[0]
    return per, clu


def run_periodic(per, coord0, latvec, setup):
    """TDynNeighList as TCoulomb keeps it (updateLatVecs / updateCoords): cutoff = maxREwald, invLatVecs = recVecs / 2 pi"""
    F = np.asfortranarray
    n = coord0.shape[0]
    nl = Obj()
    nl.cutoff, nl.nAtom, nl.tPeriodic = float(setup["maxR"]), n, True
    nl.coords0 = fx.FArr(F(np.asarray(coord0, dtype=np.float64).T))
    nl.latVecs = fx.FArr(F(np.asarray(latvec, dtype=np.float64).T))             # latVecs(:, i) = i-th lattice vector
    nl.invLatVecs = fx.FArr(F(np.linalg.inv(np.asarray(latvec, dtype=np.float64))))   # (:, i) = i-th reciprocal vector / 2 pi
    inv_r = np.full((n, n), 7.0, order="F")
    per(fx.FArr(F(np.asarray(coord0, dtype=np.float64).T)), nl, fx.FArr(F(setup["gLat"].T)), float(setup["alpha"]),
        float(setup["volume"]), fx.FArr(inv_r))
    return inv_r


def cases():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import case
    out = {}
    a = 5.0
    out["nacl"] = (np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=np.float64) * a / 2, np.eye(3) * a)
    out["skew"] = (np.array([[0.1, 0.2, 0.3], [2.0, 1.1, 0.4], [3.5, 3.0, 2.2], [1.0, 4.0, 5.5]]),
                   np.array([[6.0, 0.4, 0.0], [0.3, 7.0, 0.5], [0.0, 0.6, 8.0]]))
    w = case("fec128")[0]
    out["fec16"] = (np.asarray(w["coord0"], dtype=np.float64)[:16], np.asarray(w["latvec"], dtype=np.float64))
    return out


def main():
    if not os.path.exists(REF_C):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    sys.path.insert(0, ROOT)
    from oracle import coulomb_model as cm
    per, clu = routines()
    g = {}
    for name, (c0, lat) in cases().items():
        s = cm.ewald_setup(lat)
        g[name + "__invr"] = run_periodic(per, c0, lat, s)
        n = c0.shape[0]
        inv_c = np.full((n, n), 7.0, order="F")
        clu(fx.FArr(np.asfortranarray(c0.T)), fx.FArr(inv_c))
        g[name + "__invr_cluster"] = inv_c
        print(name, "max |invR|", float(np.abs(g[name + "__invr"]).max()), "nG", len(s["gLat"]), "nRLat", len(s["rLat"]))
    path = os.path.join(ROOT, "tests", "golden", "ewald_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Golden vectors of the short-range gamma made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
updateNrOfNeighbours_, updateShortGammaValues_ and buildShifts_ (dftb/shortgamma.F90:779-936); expGamma = the machine
translation of gen_ref_scc.py, getNrOfNeighbours = the unit-test-pinned bisection of the oracle on the list's distances.
TEST INFRASTRUCTURE ONLY.  Outputs -> tests/golden/gamma_ref_golden.npz; tests/test_oracle_gamma_ref.py compares
orc_short_gamma_values / orc_short_gamma_shifts with them."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402
import gen_ref_scc  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/shortgamma.F90"


class Obj:
    pass


def routines(orc):
    scc = gen_ref_scc.load(os.path.join(HERE, "_ref", "libsccref.so"))
    orc.orc_search_asc_real_geq.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double]

    def nr_of_neighbours(neigh, cutoff, i_at):
        n = int(neigh.nNeighbour[i_at])
        d2 = np.ascontiguousarray(neigh.neighDist2[1:n, i_at]) if n > 0 else np.zeros(1)
        return int(orc.orc_search_asc_real_geq(n, d2.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cutoff * cutoff, 1.0e-10))

    env = dict(expGamma=lambda r, a, b: scc.expGamma(r, a, b), expGammaDamped=None, getNrOfNeighbours=nr_of_neighbours)
    calls = {"getIndicesWithWorkload": "pass", "mpifx_allreduceip": "pass",
             "h5Correction%scaleShortGamma": "pass"}
    arrays = ("nHubbU", "uniqHubbU", "iHubbU", "iNeighbour", "nShell", "isDamped", "nNeighbour", "neighDist2")
    src = open(REF).read()
    cnt = fx.compile_routine(src, "subroutine", "updateNrOfNeighbours_", calls, extra_env=env, extra_arrays=arrays)[0]
    val = fx.compile_routine(src, "subroutine", "updateShortGammaValues_", calls, extra_env=env, extra_arrays=arrays)[0]
    shf = fx.compile_routine(src, "subroutine", "buildShifts_", calls, extra_env=env, extra_arrays=arrays)[0]
    return cnt, val, shf


def run(R, nl, hub, cuts, n_shell, dq_shell):
    cnt, val, shf = R
    F = np.asfortranarray
    n, m_u, n_sp, m_sh = nl["nAtom"], hub["mU"], len(hub["nHubbU"]), hub["mShell"]
    hubb = Obj()
    hubb.nHubbU = fx.FArr(hub["nHubbU"].astype(np.int64))
    hubb.uniqHubbU = fx.FArr(F(hub["uniqHubbU"].T))                         # (mU, nSpecies)
    hubb.iHubbU = fx.FArr(F(hub["iHubbU"].T.astype(np.int64)))              # (mShell, nSpecies)
    hubb.mHubbU = m_u
    neigh = Obj()
    neigh.nNeighbour = fx.FArr(nl["nNeighbour"].astype(np.int64))
    neigh.iNeighbour = fx.FArr(F(nl["iNeighbour"].T.astype(np.int64)), (0, 1))
    neigh.neighDist2 = fx.FArr(F(nl["neighDist2"].T), (0, 1))
    species = fx.FArr(nl["species"].astype(np.int64))
    coords = fx.FArr(F(nl["coords"].T))
    short_cut = fx.FArr(F(np.transpose(cuts, (3, 2, 1, 0))))                 # [sp1, sp2, u1, u2] -> (iU2, iU1, iSp2, iSp1)
    n_neigh = fx.FArr(np.zeros((m_u, m_u, n_sp, n), dtype=np.int64, order="F"))
    cnt(short_cut, species, hubb, neigh, n_neigh)
    damping = Obj()
    damping.isDamped = fx.FArr(np.zeros(n_sp, dtype=bool))
    damping.exponent = 0.0
    gam = fx.FArr(np.zeros((m_u, m_u, 1, n), order="F"), (1, 1, 0, 1))       # too small: the routine reallocates it
    loc = val(coords, species, neigh, n_neigh, hubb, damping, None, gam)
    gam = loc["shortGamma"]
    orb = Obj()
    orb.nShell = fx.FArr(np.asarray(n_shell, dtype=np.int64))
    shift = fx.FArr(np.full((m_sh, n), 7.0, order="F"))
    shf(None, orb, species, neigh.iNeighbour, fx.FArr(nl["img2CentCell"].astype(np.int64)), hubb, n_neigh, gam,
        fx.FArr(F(np.asarray(dq_shell, dtype=np.float64).T)), shift)
    return dict(nNeighU=np.ascontiguousarray(np.transpose(n_neigh.a, (3, 2, 1, 0))).astype(np.int32),
                shortGamma=np.ascontiguousarray(np.transpose(gam.a, (3, 2, 1, 0))), shiftShell=np.ascontiguousarray(shift.a.T))


def cases():
    """name -> (neighbour list, unique-Hubbard set, cutoffs, nShell, dq_shell)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import case, oracle
    from oracle import coulomb_model as cm
    out = {}
    for name, resolved in (("h2o", True), ("ce_o", True), ("fec128", False), ("fec128r", True)):
        w, sk, geo, model, selfegy = case(name.rstrip("r"))
        n_shell = [len(a) for a in sk.ang_shells]
        base = [[0.31, 0.27, 0.27, 0.22], [0.42, 0.36, 0.30, 0.25], [0.5, 0.45, 0.4, 0.35]] if resolved else [[0.31] * 4, [0.42] * 4, [0.5] * 4]
        hub = cm.unique_hubbard([base[i % 3][:n] for i, n in enumerate(n_shell)], n_shell)
        cuts = cm.short_gamma_cutoffs(hub)
        cutoff = max(sk.cutoff(), float(cuts.max()))
        rcell = np.zeros((1, 3)) if w["latvec"] is None else oracle.cell_translations(w["latvec"], cutoff)[1]
        nl = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, sk.cutoff())
        rng = np.random.default_rng(17 + nl["nAtom"])
        dq = rng.normal(size=(nl["nAtom"], hub["mShell"])) * 0.1
        for a in range(nl["nAtom"]):
            dq[a, n_shell[w["species0"][a] - 1]:] = 0.0
        out[name] = (nl, hub, cuts, n_shell, dq)
    return out


def sample(v, n=6000):
    flat = np.ravel(v)
    return flat[np.random.default_rng(7).integers(0, flat.size, size=n)]


def main():
    if not os.path.exists(REF):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    sys.path.insert(0, ROOT)
    from oracle import oracle as o
    R = routines(o.lib())
    g = {}
    for name, (nl, hub, cuts, n_shell, dq) in cases().items():
        r = run(R, nl, hub, cuts, n_shell, dq)
        for k, v in r.items():
            if k == "shortGamma" and v.size > 20000:
                g["%s__%s_shape" % (name, k)] = np.array(v.shape)
                v = sample(v)              # the large cases: 6000 entries at fixed positions
            g["%s__%s" % (name, k)] = v
        print(name, {k: v.shape for k, v in r.items()}, float(np.abs(r["shiftShell"]).max()))
    path = os.path.join(ROOT, "tests", "golden", "gamma_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()

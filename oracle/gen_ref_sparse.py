#!/usr/bin/env python3
"""Golden vectors of the sparse-layout routines made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py).

TEST INFRASTRUCTURE ONLY.  Routines, read where they lie under /root/reference:
    unpackHS_real, unpackHS_cmplx_kpts, packHS_real, packHS_cmplx_kpts      dftb/sparse2dense.F90
    mullikenPerOrbital, mullikenPerBlock, mullikenPerAtom, getOnsitePopulation  dftb/populations.F90
    addShift_atom, addShift_block                                           dftb/shift.F90
    buildH0, buildS, buildDiatomicBlocks, getFirstDerivFiniteDiff           dftb/nonscc.F90
    derivativeNonSccEuclidian, derivative_blockEuclidean                    dftb/forces.F90
    getNonSCCStress, getBlockStress                                         dftb/stress.F90
    fullDensityMatrix_real / _cmplx, fullEnergyDensityMatrix_real / _cmplx  dftb/densitymatrix.F90 (herk = the BLAS call
                                                                            C := alpha A A^T + beta C, lower triangle)
    rotateH0                                                                dftb/sk.F90 (its ss ... ff callees are the
                                                                            machine translations of gen_ref_sk.py; the SK
                                                                            integrals come from the oracle's interpolation,
                                                                            itself pinned by gen_ref_math.py)
Inputs: the small workloads of the parity tests (H2O, Ce/O cluster with f shells, the 2-atom Si cell whose images
fold onto the same dense block, bcc Fe/C 128) on the oracle's geometry; outputs -> tests/golden/sparse_ref_golden.npz.
tests/test_oracle_sparse_ref.py compares the oracle's restatements with them (and, where /root/reference exists,
re-runs the reference source)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402

REF = "/root/reference/src/dftbp/dftb/"
CASES = ("h2o", "ce_o", "si2", "fec128")
SAMPLED = ("fec128",)          # large case: 4000 entries of every output (fixed positions) instead of the whole array


def sample(v, n=4000):
    flat = np.ravel(v, order="F")
    idx = np.random.default_rng(99).integers(0, flat.size, size=min(n, flat.size))
    return flat[idx]

CALLS = {"distributeRangeWithWorkload": "{args[4]} = FArr(np.arange({args[1]}, ({args[2]}) + 1))",
         "assembleChunks": "pass"}


class Orb:
    pass


def routines():
    s2d = open(REF + "sparse2dense.F90").read()
    pop = open(REF + "populations.F90").read()
    shf = open(REF + "shift.F90").read()
    r = {}
    for name in ("unpackHS_real", "unpackHS_cmplx_kpts", "packHS_real", "packHS_cmplx_kpts"):
        r[name] = fx.compile_routine(s2d, "subroutine", name, CALLS)[0]
    for name in ("mullikenPerOrbital", "mullikenPerBlock"):
        r[name] = fx.compile_routine(pop, "subroutine", name, CALLS, extra_arrays=("nOrbAtom", "nOrbSpecies"))[0]
    pcalls = dict(CALLS)
    pcalls["mullikenPerOrbital"] = lambda a: "mullikenPerOrbital(%s)" % ", ".join(a)
    for name in ("mullikenPerAtom", "getOnsitePopulation"):
        r[name] = fx.compile_routine(pop, "subroutine", name, pcalls, extra_env=dict(mullikenPerOrbital=r["mullikenPerOrbital"]),
                                     extra_arrays=("nOrbAtom", "nOrbSpecies"))[0]
    for name in ("addShift_atom", "addShift_block"):
        r[name] = fx.compile_routine(shf, "subroutine", name, CALLS, extra_arrays=("nOrbAtom", "nOrbSpecies"))[0]
    # buildH0 / buildS: the reference's loops, rotation driver and shell-pair routines; integrals from the oracle
    import ctypes
    sk_src = open(REF + "sk.F90").read()
    non = open(REF + "nonscc.F90").read()
    skref = ctypes.CDLL(os.path.join(HERE, "_ref", "libskref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    skref.skref_block.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp]
    ang = {"s": 0, "p": 1, "d": 2, "f": 3}

    def ref_pair(which, tmpH, *a):
        sk = np.ascontiguousarray(a[-1], dtype=np.float64)
        ll, mm, nn = (a[0], a[1], a[2]) if len(a) == 4 else (0.0, 0.0, 1.0)
        hh = np.zeros(49)
        skref.skref_block(ang[which[0]], ang[which[1]], ll, mm, nn, sk.ctypes.data_as(dp), hh.ctypes.data_as(dp))
        tmpH.a[:, :] = hh.reshape((7, 7), order="F")

    pairs = {n: (lambda a, n=n: "ref_pair(%r, %s)" % (n, ", ".join(a))) for n in ("ss", "sp", "sd", "sf", "pp", "pd", "pf", "dd", "df", "ff")}
    orb_arrays = ("nShell", "angShell", "nOrbSpecies", "nOrbAtom", "iShellOrb")
    rot = fx.compile_routine(sk_src, "subroutine", "rotateH0", pairs, extra_env=dict(mAngRot_=3, ref_pair=ref_pair),
                             extra_arrays=orb_arrays)[0]

    def get_sk_integrals(cont, inter, dist, sp1, sp2):
        d, tab, method = cont[(int(sp1), int(sp2))]
        out = np.zeros(tab.shape[1])
        _lib().orc_sk_eval(tab.shape[0], tab.shape[1], ctypes.c_double(d), tab.ctypes.data_as(dp), int(method),
                           ctypes.c_double(dist), out.ctypes.data_as(dp))
        inter.a[:tab.shape[1]] = out

    calls = dict(CALLS)
    calls["distributeRangeInChunks"] = "{args[3]}, {args[4]} = {args[1]}, {args[2]}"
    calls["getSKIntegrals"] = lambda a: "getSKIntegrals(%s)" % ", ".join(a)
    calls["rotateH0"] = lambda a: "rotateH0(%s)" % ", ".join(a)
    calls["buildDiatomicBlocks"] = lambda a: "buildDiatomicBlocks(%s)" % ", ".join(a)
    env = dict(getSKIntegrals=get_sk_integrals, rotateH0=rot, getMIntegrals=lambda cont: max(t[1].shape[1] for t in cont.values()))
    bdb = fx.compile_routine(non, "subroutine", "buildDiatomicBlocks", calls, extra_env=env, extra_arrays=orb_arrays)[0]
    env["buildDiatomicBlocks"] = bdb
    r["buildDiatomicBlocks"] = bdb
    for name in ("buildH0", "buildS"):
        r[name] = fx.compile_routine(non, "subroutine", name, calls, extra_env=env, extra_arrays=orb_arrays)[0]
    # forces: the finite-difference derivator and the two contraction routines
    fdd = fx.compile_routine(non, "subroutine", "getFirstDerivFiniteDiff", calls, extra_env=env, extra_arrays=orb_arrays)[0]
    env["getFirstDerivFiniteDiff"] = fdd
    frc = open(REF + "forces.F90").read()
    calls["derivator%getFirstDeriv"] = lambda a: "getFirstDerivFiniteDiff(%s, derivator)" % ", ".join(a)
    for name in ("derivativeNonSccEuclidian", "derivative_blockEuclidean"):
        r[name] = fx.compile_routine(frc, "subroutine", name, calls, extra_env=env, extra_arrays=orb_arrays)[0]
    # the Richardson derivator with the non-SCC force contraction
    rich = fx.compile_routine(non, "subroutine", "getFirstDerivRichardson", calls, extra_env=env, extra_arrays=orb_arrays)[0]
    env["getFirstDerivRichardson"] = rich
    calls_r = dict(calls)
    calls_r["derivator%getFirstDeriv"] = lambda a: "getFirstDerivRichardson(%s)" % ", ".join(a)
    r["derivativeNonSccEuclidian_richardson"] = fx.compile_routine(frc, "subroutine", "derivativeNonSccEuclidian", calls_r,
                                                                   extra_env=env, extra_arrays=orb_arrays)[0]
    strs = open(REF + "stress.F90").read()
    for name in ("getNonSCCStress", "getBlockStress"):
        r[name] = fx.compile_routine(strs, "subroutine", name, calls, extra_env=env, extra_arrays=orb_arrays)[0]

    def herk(c, a, alpha=1.0, beta=0.0):
        a = a.a if isinstance(a, fx.FArr) else a
        c.a[:, :] = np.tril(alpha * (a @ a.conj().T) + beta * c.a) + np.triu(c.a, 1)

    dms = open(REF + "densitymatrix.F90").read()
    dcalls = {"herk": lambda a: "herk(%s)" % ", ".join(a)}
    for name in ("fullDensityMatrix_real", "fullEnergyDensityMatrix_real", "fullDensityMatrix_cmplx", "fullEnergyDensityMatrix_cmplx"):
        r[name] = fx.compile_routine(dms, "subroutine", name, dcalls, extra_env=dict(herk=herk, arbitraryConstant=0.1))[0]
    return r


def _lib():
    sys.path.insert(0, ROOT)
    from oracle import oracle as o
    return o.lib()


def inputs(name):
    """Deterministic inputs of one case (shared with the test): geometry arrays in Fortran dimension order + random
    sparse / dense data."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import case
    w, sk, geo, model, selfegy = case(name)
    rng = np.random.default_rng(len(name) * 1000 + geo["nAtom"])
    n, ss, nA, mOrb = geo["nOrb"], geo["sparseSize"], geo["nAtom"], int(max(sk.n_orb_species))
    d = dict(geo=geo, mOrb=mOrb, nOrbSpecies=np.array(sk.n_orb_species, dtype=np.int64), model=model, selfegy=selfegy)
    d["orig"] = rng.normal(size=ss)
    d["over"] = rng.normal(size=ss)
    d["rho"] = rng.normal(size=ss)
    sq = rng.normal(size=(n, n))
    d["square"] = np.asfortranarray(sq)
    d["square_c"] = np.asfortranarray(sq + 1j * rng.normal(size=(n, n)))
    d["kpoint"] = np.array([0.13, -0.27, 0.41])
    d["kweight"] = 0.375
    d["shift_atom"] = rng.normal(size=(2, nA))                      # [spin, atom] = Fortran (nAtom, nSpin)
    d["shift_block"] = rng.normal(size=(2, nA, mOrb, mOrb))         # Fortran (mOrb, mOrb, nAtom, nSpin)
    d["ham0"] = rng.normal(size=(2, ss))
    return d


def run_case(R, d):
    geo = d["geo"]
    nA, n, ss, mOrb = geo["nAtom"], geo["nOrb"], geo["sparseSize"], d["mOrb"]
    F = np.asfortranarray
    iN = F(geo["iNeighbour"].T.astype(np.int64))
    iP = F(geo["iPair"].T.astype(np.int64))
    nN = geo["nNeighbourSK"].astype(np.int64)
    start = geo["iAtomStart"].astype(np.int64)
    img = geo["img2CentCell"].astype(np.int64)
    icv = geo["iCellVec"].astype(np.int64)
    cellvec = F(np.asarray(geo["cellVec"], dtype=np.float64).T)
    species = geo["species"].astype(np.int64)
    orb = Orb()
    orb.mOrb = mOrb
    orb.nOrbSpecies = fx.FArr(d["nOrbSpecies"])
    orb.nOrbAtom = fx.FArr(np.diff(start))
    out = {}

    def W(fn, nm, a):
        return fx.wrap(fn, nm, a)

    f = R["unpackHS_real"]
    sq = np.full((n, n), 7.0, order="F")
    f(W(f, "square", sq), W(f, "orig", d["orig"]), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "iAtomStart", start),
      W(f, "iSparseStart", iP), W(f, "img2CentCell", img))
    out["unpack_real"] = sq
    f = R["unpackHS_cmplx_kpts"]
    sqc = np.full((n, n), 7.0 + 1j, order="F", dtype=np.complex128)
    f(W(f, "square", sqc), W(f, "orig", d["orig"]), W(f, "kPoint", d["kpoint"]), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN),
      W(f, "iCellVec", icv), W(f, "cellVec", cellvec), W(f, "iAtomStart", start), W(f, "iSparseStart", iP), W(f, "img2CentCell", img))
    out["unpack_cplx"] = sqc
    f = R["packHS_real"]
    prim = d["rho"].copy()
    f(W(f, "primitive", prim), W(f, "square", d["square"].copy(order="F")), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), mOrb,
      W(f, "iAtomStart", start), W(f, "iSparseStart", iP), W(f, "img2CentCell", img))
    out["pack_real"] = prim
    f = R["packHS_cmplx_kpts"]
    prim = d["rho"].copy()
    f(W(f, "primitive", prim), W(f, "square", d["square_c"].copy(order="F")), W(f, "kPoint", d["kpoint"]), d["kweight"],
      W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), mOrb, W(f, "iCellVec", icv), W(f, "cellVec", cellvec),
      W(f, "iAtomStart", start), W(f, "iSparseStart", iP), W(f, "img2CentCell", img))
    out["pack_cplx"] = prim
    f = R["mullikenPerOrbital"]
    qq = np.zeros((mOrb, nA), order="F")
    f(None, W(f, "qq", qq), W(f, "over", d["over"]), W(f, "rho", d["rho"]), orb, W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN),
      W(f, "img2CentCell", img), W(f, "iPair", iP))
    out["mulliken_orbital"] = qq
    f = R["mullikenPerBlock"]
    qb = np.zeros((mOrb, mOrb, nA), order="F")
    f(None, W(f, "qq", qb), W(f, "over", d["over"]), W(f, "rho", d["rho"]), orb, W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN),
      W(f, "img2CentCell", img), W(f, "iPair", iP))
    out["mulliken_block"] = qb
    f = R["mullikenPerAtom"]
    qa = np.linspace(-1.0, 1.0, nA)                                # accumulates into qq
    f(None, W(f, "qq", qa), W(f, "over", d["over"]), W(f, "rho", d["rho"]), orb, W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN),
      W(f, "img2CentCell", img), W(f, "iPair", iP))
    out["mulliken_atom"] = qa
    f = R["getOnsitePopulation"]
    qo = np.full(nA, 7.0)
    f(W(f, "rho", d["rho"]), orb, W(f, "iPair", iP), W(f, "qq", qo))
    out["onsite_population"] = qo
    # buildH0 / buildS on the reference's TOrbitals view of the model
    model = d["model"]
    nSp, mShell = model.nSpecies, model.mShell
    orb.nShell = fx.FArr(model.nShell.astype(np.int64))
    orb.angShell = fx.FArr(F(model.angShell.T.astype(np.int64)))            # (mShell, nSpecies)
    ishell = np.zeros((mOrb, nSp), dtype=np.int64, order="F")                 # iShellOrb(mOrb, nSpecies)
    for sp in range(nSp):
        k = 0
        for ish, l in enumerate(model.ang_shells[sp]):
            ishell[k:k + 2 * l + 1, sp] = ish + 1
            k += 2 * l + 1
    orb.iShellOrb = fx.FArr(ishell)
    selfegy = np.zeros((mShell, nSp), order="F")
    for sp in range(nSp):
        selfegy[:len(d["selfegy"][sp]), sp] = d["selfegy"][sp]
    coords = F(np.asarray(geo["coords"], dtype=np.float64).T)                  # (3, nAllAtom)
    contH = {(a, b): model.tables[(0, a, b)] for a in range(1, nSp + 1) for b in range(1, nSp + 1)}
    contS = {(a, b): model.tables[(1, a, b)] for a in range(1, nSp + 1) for b in range(1, nSp + 1)}
    f = R["buildH0"]
    ham = np.full(ss, 7.0)
    f(None, W(f, "ham", ham), contH, W(f, "selfegy", selfegy), W(f, "coords", coords), W(f, "nNeighbourSK", nN), W(f, "iNeighbours", iN),
      W(f, "species", species), W(f, "iPair", iP), orb)
    out["build_h0"] = ham
    f = R["buildS"]
    over = np.full(ss, 7.0)
    f(None, W(f, "over", over), contS, W(f, "coords", coords), W(f, "nNeighbourSK", nN), W(f, "iNeighbours", iN), W(f, "species", species),
      W(f, "iPair", iP), orb)
    out["build_s"] = over
    # forces (derivative_shift): `derivator` is passed as the finite-difference step deltaXDiff = epsilon**0.25 (nonscc.F90:43)
    delta = float(np.finfo(np.float64).eps) ** 0.25
    dm2 = F(np.stack([d["rho"], 0.3 * d["over"]], axis=1))                   # DM(sparseSize, nSpin = 2)
    f = R["derivativeNonSccEuclidian"]
    der = np.full((3, nA), 7.0, order="F")
    f(None, W(f, "deriv", der), delta, W(f, "DM", d["rho"]), W(f, "EDM", d["over"]), contH, contS, W(f, "coords", coords),
      W(f, "species", species), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "img2CentCell", img), W(f, "iPair", iP), orb)
    out["deriv_nonscc"] = der
    if geo["nAtom"] <= 8:          # the Richardson tables are slow in the interpreter: the small cases only
        f = R["derivativeNonSccEuclidian_richardson"]
        der = np.full((3, nA), 7.0, order="F")
        f(None, W(f, "deriv", der), None, W(f, "DM", d["rho"]), W(f, "EDM", d["over"]), contH, contS, W(f, "coords", coords),
          W(f, "species", species), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "img2CentCell", img), W(f, "iPair", iP), orb)
        out["deriv_nonscc_richardson"] = der
    f = R["derivative_blockEuclidean"]
    der = np.full((3, nA), 7.0, order="F")
    shb = F(np.transpose(d["shift_block"], (3, 2, 1, 0)).copy())
    f(None, W(f, "deriv", der), delta, W(f, "DM", dm2), W(f, "EDM", d["orig"]), contH, contS, W(f, "coords", coords),
      W(f, "species", species), W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "img2CentCell", img), W(f, "iPair", iP), orb,
      W(f, "shift", shb))
    out["deriv_block"] = der
    # stress (periodic cases; the cluster cases run it too: a volume of 1000 bohr^3 is as good as any)
    vol = 1000.0
    f = R["getNonSCCStress"]
    st = np.full((3, 3), 7.0, order="F")
    f(None, W(f, "st", st), delta, W(f, "DM", d["rho"]), W(f, "EDM", d["over"]), contH, contS, W(f, "coords", coords), W(f, "species", species),
      W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "img2CentCell", img), W(f, "iPair", iP), orb, vol)
    out["stress_nonscc"] = st
    f = R["getBlockStress"]
    st = np.full((3, 3), 7.0, order="F")
    f(None, W(f, "st", st), delta, W(f, "DM", dm2), W(f, "EDM", d["orig"]), contH, contS, W(f, "coords", coords), W(f, "species", species),
      W(f, "iNeighbour", iN), W(f, "nNeighbourSK", nN), W(f, "img2CentCell", img), W(f, "iPair", iP), orb, W(f, "shift", shb), vol)
    out["stress_block"] = st
    # density matrices: both branches of both routines (positive fillings / zero and negative weights)
    nd = min(n, 40)
    rngd = np.random.default_rng(5 + nd)
    C = F(rngd.normal(size=(nd, nd)))
    eig = np.sort(rngd.normal(size=nd))
    for tag, fill in (("pos", np.concatenate([np.full(nd // 2, 2.0), rngd.uniform(0.1, 1.9, 2), np.zeros(nd - nd // 2 - 2)])),
                      ("mixed", np.concatenate([rngd.normal(size=nd // 2), np.zeros(nd - nd // 2)]))):
        for key, fname in (("dm", "fullDensityMatrix_real"), ("edm", "fullEnergyDensityMatrix_real")):
            f = R[fname]
            dm = np.full((nd, nd), 7.0, order="F")
            cc = C.copy(order="F")
            if key == "dm":
                f(W(f, "dm", dm), W(f, "eigenvecs", cc), W(f, "filling", fill))
            else:
                f(W(f, "dm", dm), W(f, "eigenvecs", cc), W(f, "filling", fill), W(f, "eigen", eig if tag == "mixed" else -np.abs(eig) - 0.1))
            out["%s_%s" % (key, tag)] = dm
            out["%s_%s_cback" % (key, tag)] = cc       # the eigenvectors are scaled and scaled back
        for key, fname in (("dmc", "fullDensityMatrix_cmplx"), ("edmc", "fullEnergyDensityMatrix_cmplx")):
            f = R[fname]
            dmc = np.full((nd, nd), 7.0 + 1j, order="F", dtype=np.complex128)
            ccz = F(C + 1j * C[::-1, ::-1])
            if key == "dmc":
                f(W(f, "dm", dmc), W(f, "eigenvecs", ccz), W(f, "filling", fill))
            else:
                f(W(f, "dm", dmc), W(f, "eigenvecs", ccz), W(f, "filling", fill), W(f, "eigen", eig if tag == "mixed" else -np.abs(eig) - 0.1))
            out["%s_%s" % (key, tag)] = dmc
        out["dmin_%s_fill" % tag] = fill
    out["dmin_C"], out["dmin_eig"] = C, eig
    for zero in (True, False):
        f = R["addShift_atom"]
        ham = F(d["ham0"].T.copy())                                    # (sparseSize, nSpin)
        f(None, W(f, "ham", ham), W(f, "over", d["over"]), W(f, "nNeighbour", nN), W(f, "iNeighbour", iN), W(f, "species", species), orb,
          W(f, "iPair", iP), nA, W(f, "img2CentCell", img), W(f, "shift", F(d["shift_atom"].T.copy())), zero)
        out["add_shift_atom_%d" % zero] = ham
        f = R["addShift_block"]
        ham = F(d["ham0"].T.copy())
        shb = F(np.transpose(d["shift_block"], (3, 2, 1, 0)).copy())   # [spin, atom, b, a] -> Fortran (a, b, atom, spin)
        f(None, W(f, "ham", ham), W(f, "over", d["over"]), W(f, "nNeighbour", nN), W(f, "iNeighbour", iN), W(f, "species", species), orb,
          W(f, "iPair", iP), nA, W(f, "img2CentCell", img), W(f, "shift", shb), zero)
        out["add_shift_block_%d" % zero] = ham
    return out


def main():
    if not os.path.exists(REF + "sparse2dense.F90"):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    R = routines()
    g = {}
    for name in CASES:
        out = run_case(R, inputs(name))
        for k, v in out.items():
            g["%s__%s" % (name, k)] = sample(v) if name in SAMPLED else v
        print(name, "done:", ", ".join(sorted(out)))
    path = os.path.join(ROOT, "tests", "golden", "sparse_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()

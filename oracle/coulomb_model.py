"""Ewald set-up of the SCC Coulomb term restated (TEST INFRASTRUCTURE, like everything under oracle/):
TCoulomb%updateLatVecs (dftb/coulomb.F90:330-372): getOptimalAlphaEwald (:1824-1908), getMaxGEwald
(:1911-1975), getMaxREwald (:1978-2038), diffRecReal / gTerm / rTerm (:2424-2489), getLatticePoints with
onlyInside / reduceByInversion / withoutOrigin (dftb/periodic.F90:277-330, type/latpointiter.F90:112-228),
nSearchIter = 30 (common/accuracy.F90:114), EwaldTolerance default 1e-9 (dftbplus/parser.F90:3364).
The matrices themselves are in dftb_oracle.c (orc_invr_periodic, orc_invr_cluster, orc_symv_lower).
Parity unpinned against the Fortran binary (it cannot be built here); pinned to physics through the
Madelung constants of NaCl and CsCl (tests/test_coulomb_oracle.py)."""
import ctypes
import math

import numpy as np

from . import oracle as _o

N_SEARCH_ITER = 30
TOL_EWALD = 1.0e-9
HUGE = float(np.finfo(np.float64).max)


def g_term(gg, alpha, vol):
    return 4.0 * math.pi * (math.exp(-0.25 * gg ** 2 / alpha ** 2) / (vol * gg ** 2))


def r_term(rr, alpha):
    return math.erfc(alpha * rr) / rr


def diff_rec_real(alpha, min_g, min_r, vol):
    return ((g_term(4.0 * min_g, alpha, vol) - g_term(5.0 * min_g, alpha, vol))
            - (r_term(2.0 * min_r, alpha) - r_term(3.0 * min_r, alpha)))


def optimal_alpha_ewald(latvec, recvec, vol, tol=TOL_EWALD):
    """latvec[i], recvec[i] = i-th (reciprocal, with 2 pi) lattice vector."""
    min_g = math.sqrt(min(float(np.sum(v ** 2)) for v in recvec))
    min_r = math.sqrt(min(float(np.sum(v ** 2)) for v in latvec))
    alpha_init = 1.0e-8
    alpha = alpha_init
    diff = diff_rec_real(alpha, min_g, min_r, vol)
    while diff < -tol and alpha <= HUGE:
        alpha = 2.0 * alpha
        diff = diff_rec_real(alpha, min_g, min_r, vol)
    if alpha > HUGE or alpha == alpha_init:
        raise RuntimeError("Failure in determining optimal alpha for Ewaldsum.")
    alpha_left = 0.5 * alpha
    while diff < tol and alpha <= HUGE:
        alpha = 2.0 * alpha
        diff = diff_rec_real(alpha, min_g, min_r, vol)
    if alpha > HUGE:
        raise RuntimeError("Failure in determining optimal alpha for Ewaldsum.")
    alpha_right = alpha
    alpha = (alpha_left + alpha_right) / 2.0
    it = 0
    diff = diff_rec_real(alpha, min_g, min_r, vol)
    while abs(diff) > tol and it <= N_SEARCH_ITER:
        if diff < 0:
            alpha_left = alpha
        else:
            alpha_right = alpha
        alpha = (alpha_left + alpha_right) / 2.0
        diff = diff_rec_real(alpha, min_g, min_r, vol)
        it += 1
    if it > N_SEARCH_ITER:
        raise RuntimeError("Failure in determining optimal alpha for Ewaldsum.")
    return alpha


def _max_x(term, min_value):
    xx = 1.0e-8
    yy = term(xx)
    while yy > min_value and xx <= HUGE:
        xx = 2.0 * xx
        yy = term(xx)
    if xx > HUGE or xx == 1.0e-8:
        raise RuntimeError("Failure in getMax*Ewald")
    x_left, x_right = 0.5 * xx, xx
    y_left, y_right = term(x_left), yy
    it = 1
    while y_left - y_right > min_value and it <= N_SEARCH_ITER:
        xx = 0.5 * (x_left + x_right)
        yy = term(xx)
        if yy >= min_value:
            x_left, y_left = xx, yy
        else:
            x_right, y_right = xx, yy
        it += 1
    if it > N_SEARCH_ITER:
        raise RuntimeError("Failure in getMax*Ewald")
    return xx


def max_g_ewald(alpha, vol, min_value=TOL_EWALD):
    return _max_x(lambda g: g_term(g, alpha, vol), min_value)


def max_r_ewald(alpha, min_value=TOL_EWALD):
    return _max_x(lambda r: r_term(r, alpha), min_value)


def lattice_points(latvec, inv2p, dist, only_inside=True, reduce_by_inversion=True):
    """getLatticePoints(..., onlyInside, reduceByInversion, withoutOrigin = .true.) with pos / negExtension 0:
    integer points (rows) in iterator order.  latvec[i] = i-th vector of the lattice the points belong to,
    inv2p[i] = i-th column of the matrix getRanges_ takes (the reciprocal of that lattice without 2 pi)."""
    rng = [int(math.floor(dist * math.sqrt(float(np.sum(inv2p[i] ** 2))))) for i in range(3)]
    lo = [-r for r in rng]
    hi = list(rng)
    pts = []
    cur = list(lo)

    def nxt(p):
        p[2] += 1
        if p[2] > hi[2]:
            p[2] = lo[2]
            p[1] += 1
            if p[1] > hi[1]:
                p[1] = lo[1]
                p[0] += 1

    while True:
        if cur[0] > hi[0]:
            break
        if reduce_by_inversion:
            if cur[0] < 0:
                cur[0] = 0
            if cur[1] < 0 and cur[0] == 0:
                cur[1] = 0
            if cur[2] < 0 and cur[0] == 0 and cur[1] == 0:
                cur[2] = 0
        if cur == [0, 0, 0]:
            nxt(cur)
            continue
        p = np.array(cur, dtype=np.float64)
        rr = p[0] * latvec[0] + p[1] * latvec[1] + p[2] * latvec[2]
        if (not only_inside) or float(np.sum(rr ** 2)) <= dist * dist:
            pts.append(p)
        nxt(cur)
    return np.array(pts, dtype=np.float64).reshape(-1, 3)


def ewald_setup(latvec, tol=TOL_EWALD, alpha=None):
    """TCoulomb%updateLatVecs: alpha, the real-space cutoff and lattice translations (origin first, absolute
    coordinates), the reciprocal points gLatPoints (absolute), the cell volume."""
    latvec = np.ascontiguousarray(latvec, dtype=np.float64)
    inv = np.linalg.inv(latvec).T            # inv[i] = i-th reciprocal vector without 2 pi
    rec = 2.0 * math.pi * inv
    vol = abs(float(np.linalg.det(latvec)))
    if alpha is None:
        alpha = optimal_alpha_ewald(latvec, rec, vol, tol)
    max_r = max_r_ewald(alpha, tol)
    max_g = max_g_ewald(alpha, vol, tol)
    g_int = lattice_points(rec, latvec / (2.0 * math.pi), max_g)
    g_abs = g_int @ rec                      # matmul(recVecs, gLatPoints)
    _cell, r_abs = _o.cell_translations(latvec, max_r)
    return dict(alpha=alpha, volume=vol, maxR=max_r, maxG=max_g, gLat=np.ascontiguousarray(g_abs),
                rLat=np.ascontiguousarray(r_abs))


def invr_cluster(coord):
    coord = np.ascontiguousarray(coord, dtype=np.float64)
    n = coord.shape[0]
    a = np.zeros((n, n), order="F")
    _o.lib().orc_invr_cluster(n, _o._d(coord), _o._d(a), n)
    return a


def invr_periodic(coord, setup):
    coord = np.ascontiguousarray(coord, dtype=np.float64)
    n = coord.shape[0]
    a = np.zeros((n, n), order="F")
    _o.lib().orc_invr_periodic(n, _o._d(coord), setup["rLat"].shape[0], _o._d(setup["rLat"]), setup["gLat"].shape[0],
                               _o._d(setup["gLat"]), ctypes.c_double(setup["alpha"]), ctypes.c_double(setup["volume"]),
                               ctypes.c_double(setup["maxR"]), _o._d(a), n)
    return a


def coulomb_shift(invr, dq):
    dq = np.ascontiguousarray(dq, dtype=np.float64)
    n = dq.shape[0]
    y = np.zeros(n)
    _o.lib().orc_symv_lower(n, _o._d(invr), invr.shape[0], _o._d(dq), _o._d(y))
    return y


# ---- short-range gamma (dftb/shortgamma.F90, dftb/shortgammafuncs.F90, dftb/uniquehubbard.F90) ----------------
MIN_HUB_DIFF = 0.3125e-5


def unique_hubbard(hubb_u, n_shell):
    """TUniqueHubbard_init (dftb/uniquehubbard.F90:58-100): hubb_u[sp][shell] -> nHubbU[sp], uniqHubbU[sp][mU],
    iHubbU[sp][mShell] (1-based, 0 = no such shell), mU."""
    n_sp = len(n_shell)
    m_shell = max(n_shell)
    n_u = [1] * n_sp
    uniq = np.zeros((n_sp, m_shell))
    i_u = np.zeros((n_sp, m_shell), dtype=np.int32)
    for sp in range(n_sp):
        uniq[sp, 0] = hubb_u[sp][0]
        i_u[sp, 0] = 1
        for il in range(1, n_shell[sp]):
            for iu in range(n_u[sp]):
                if abs(hubb_u[sp][il] - uniq[sp, iu]) < MIN_HUB_DIFF:
                    i_u[sp, il] = iu + 1
                    break
            else:
                uniq[sp, n_u[sp]] = hubb_u[sp][il]
                n_u[sp] += 1
                i_u[sp, il] = n_u[sp]
    m_u = max(n_u)
    return dict(nHubbU=np.array(n_u, dtype=np.int32), uniqHubbU=np.ascontiguousarray(uniq[:, :m_u]), iHubbU=i_u, mU=m_u,
                mShell=m_shell, nShell=np.array(n_shell, dtype=np.int32))


def _gamma_lib():
    lib = _o.lib()
    lib.orc_exp_gamma.restype = ctypes.c_double
    lib.orc_exp_gamma.argtypes = [ctypes.c_double] * 3
    lib.orc_exp_gamma_cutoff.restype = ctypes.c_double
    lib.orc_exp_gamma_cutoff.argtypes = [ctypes.c_double] * 2
    return lib


def exp_gamma(rab, ua, ub):
    return _gamma_lib().orc_exp_gamma(rab, ua, ub)


def short_gamma_cutoffs(hub):
    """setupShortGammaCutoffs_: shortCutoffs[sp1][sp2][iU1][iU2] = Fortran (iU2, iU1, iSp2, iSp1)."""
    lib = _gamma_lib()
    n_sp, m_u = len(hub["nHubbU"]), hub["mU"]
    c = np.zeros((n_sp, n_sp, m_u, m_u))
    for s1 in range(n_sp):
        for s2 in range(s1, n_sp):
            for u1 in range(hub["nHubbU"][s1]):
                for u2 in range(hub["nHubbU"][s2]):
                    v = lib.orc_exp_gamma_cutoff(hub["uniqHubbU"][s2, u2], hub["uniqHubbU"][s1, u1])
                    c[s1, s2, u1, u2] = v
                    c[s2, s1, u2, u1] = v
    return c


def short_gamma_values(nl, hub, cutoffs):
    """updateCoords of TShortGamma on the full neighbour list nl (oracle.neighbour_list_full)."""
    n, m_n, m_u, n_sp = nl["nAtom"], nl["mNeigh"], hub["mU"], len(hub["nHubbU"])
    n_neigh_u = np.zeros((n, n_sp, m_u, m_u), dtype=np.int32)
    max_neigh = np.zeros(n, dtype=np.int32)
    gam = np.zeros((n, m_n + 1, m_u, m_u))
    _o.lib().orc_short_gamma_values(n, m_n, m_u, n_sp, _o._i(hub["nHubbU"]), _o._d(hub["uniqHubbU"]),
                                    _o._d(np.ascontiguousarray(cutoffs)), _o._i(nl["species"]), _o._d(nl["coords"]),
                                    _o._i(nl["nNeighbour"]), _o._i(nl["iNeighbour"]), _o._d(nl["neighDist2"]),
                                    _o._i(n_neigh_u), _o._i(max_neigh), _o._d(gam))
    return dict(nNeighU=n_neigh_u, maxNeigh=max_neigh, shortGamma=gam)


def short_gamma_shifts(nl, hub, vals, dq_shell):
    n, m_n, m_u, m_sh = nl["nAtom"], nl["mNeigh"], hub["mU"], hub["mShell"]
    dq_shell = np.ascontiguousarray(dq_shell, dtype=np.float64)
    out = np.zeros((n, m_sh))
    _o.lib().orc_short_gamma_shifts(n, m_n, m_u, m_sh, len(hub["nHubbU"]), _o._i(hub["nShell"]),
                                    _o._i(np.ascontiguousarray(hub["iHubbU"])), _o._i(nl["species"]), _o._i(nl["img2CentCell"]),
                                    _o._i(nl["iNeighbour"]), _o._i(vals["maxNeigh"]), _o._d(vals["shortGamma"]),
                                    _o._d(dq_shell), _o._d(out))
    return out

"""Host-side set-up of the SCC potential, mirroring the reference routines whose RESULTS the device routines
take as arguments (the Fortran host keeps calling its own; this mirror serves bench.py and the harness):

  ewald_setup          TCoulomb%updateLatVecs, dftb/coulomb.F90:330-372 (getOptimalAlphaEwald :1824-1908,
                       getMaxGEwald :1911-1975, getMaxREwald :1978-2038, getLatticePoints with onlyInside /
                       reduceByInversion / withoutOrigin, dftb/periodic.F90:277-330, getCellTranslations)
  unique_hubbard       TUniqueHubbard_init, dftb/uniquehubbard.F90:58-100
  short_gamma_cutoffs  setupShortGammaCutoffs_, dftb/shortgamma.F90:753-776 (expGammaCutoff / expGamma,
                       dftb/shortgammafuncs.F90:36-148)

Independent of oracle/coulomb_model.py (tests compare the two)."""
import math

import numpy as np

TOL_EWALD = 1.0e-9          # EwaldTolerance default, dftbplus/parser.F90:3364
N_SEARCH_ITER = 30          # common/accuracy.F90:114
MIN_HUB_DIFF = 0.3125e-5
TOL_SAME_DIST = 1.0e-5
MIN_SHORT_GAMMA = TOL_SHORT_GAMMA = 1.0e-10
_HUGE = float(np.finfo(np.float64).max)


def _g_term(g, alpha, vol):
    return 4.0 * math.pi * (math.exp(-0.25 * g * g / (alpha * alpha)) / (vol * g * g))


def _r_term(r, alpha):
    return math.erfc(alpha * r) / r


def _bracket_and_bisect(term, min_value, what):
    x = 1.0e-8
    y = term(x)
    while y > min_value and x <= _HUGE:
        x *= 2.0
        y = term(x)
    if x > _HUGE or x == 1.0e-8:
        raise RuntimeError("Failure in " + what)
    lo, hi, y_lo, y_hi = 0.5 * x, x, term(0.5 * x), y
    it = 1
    while y_lo - y_hi > min_value and it <= N_SEARCH_ITER:
        x = 0.5 * (lo + hi)
        y = term(x)
        if y >= min_value:
            lo, y_lo = x, y
        else:
            hi, y_hi = x, y
        it += 1
    if it > N_SEARCH_ITER:
        raise RuntimeError("Failure in " + what)
    return x


def optimal_alpha(latvec, recvec, vol, tol=TOL_EWALD):
    min_g = math.sqrt(min(float(v @ v) for v in recvec))
    min_r = math.sqrt(min(float(v @ v) for v in latvec))

    def diff(a):
        return ((_g_term(4.0 * min_g, a, vol) - _g_term(5.0 * min_g, a, vol))
                - (_r_term(2.0 * min_r, a) - _r_term(3.0 * min_r, a)))

    a = 1.0e-8
    d = diff(a)
    while d < -tol and a <= _HUGE:
        a *= 2.0
        d = diff(a)
    if a > _HUGE or a == 1.0e-8:
        raise RuntimeError("Failure in determining optimal alpha for Ewaldsum.")
    left = 0.5 * a
    while d < tol and a <= _HUGE:
        a *= 2.0
        d = diff(a)
    right = a
    a = (left + right) / 2.0
    d = diff(a)
    it = 0
    while abs(d) > tol and it <= N_SEARCH_ITER:
        if d < 0:
            left = a
        else:
            right = a
        a = (left + right) / 2.0
        d = diff(a)
        it += 1
    if it > N_SEARCH_ITER:
        raise RuntimeError("Failure in determining optimal alpha for Ewaldsum.")
    return a


def _points(vecs, inv2p, dist, ext, half):
    """Lattice points of TLatPointIter: ranges from getRanges_ (+ ext on both sides), third index fastest.
    half: onlyInside + reduceByInversion + no origin (the reciprocal points); else all points, origin first."""
    rng = [int(math.floor(dist * math.sqrt(float(inv2p[i] @ inv2p[i])))) + ext for i in range(3)]
    out = []
    for a in range(0 if half else -rng[0], rng[0] + 1):
        for b in range(-rng[1], rng[1] + 1):
            if half and a == 0 and b < 0:
                continue
            for c in range(-rng[2], rng[2] + 1):
                if (a, b, c) == (0, 0, 0) or (half and a == 0 and b == 0 and c < 0):
                    continue
                p = a * vecs[0] + b * vecs[1] + c * vecs[2]
                if half and float(p @ p) > dist * dist:
                    continue
                out.append(p)
    pts = np.array(out, dtype=np.float64).reshape(-1, 3)
    return pts if half else np.vstack([np.zeros((1, 3)), pts])


def ewald_setup(latvec, tol=TOL_EWALD, alpha=None):
    """latvec[i] = i-th lattice vector.  Returns dict(alpha, volume, maxR, maxG, rLat [nRLat, 3] origin first,
    gLat [nG, 3]) -- the arguments of HotPath.invRPeriodic."""
    latvec = np.ascontiguousarray(latvec, dtype=np.float64)
    inv = np.linalg.inv(latvec).T
    rec = 2.0 * math.pi * inv
    vol = abs(float(np.linalg.det(latvec)))
    if alpha is None:
        alpha = optimal_alpha(latvec, rec, vol, tol)
    max_r = _bracket_and_bisect(lambda r: _r_term(r, alpha), tol, "getMaxREwald")
    max_g = _bracket_and_bisect(lambda g: _g_term(g, alpha, vol), tol, "getMaxGEwald")
    return dict(alpha=alpha, volume=vol, maxR=max_r, maxG=max_g,
                rLat=np.ascontiguousarray(_points(latvec, inv, max_r, 1, False)),
                gLat=np.ascontiguousarray(_points(rec, latvec / (2.0 * math.pi), max_g, 0, True)))


def unique_hubbard(hubb_u, n_shell):
    """hubb_u[sp][shell] -> dict(nHubbU [nSp], uniqHubbU [nSp, mU], iHubbU [nSp, mShell] 1-based, mU, mShell)."""
    n_sp, m_shell = len(n_shell), max(n_shell)
    n_u = np.ones(n_sp, dtype=np.int32)
    uniq = np.zeros((n_sp, m_shell))
    i_u = np.zeros((n_sp, m_shell), dtype=np.int32)
    for sp in range(n_sp):
        uniq[sp, 0], i_u[sp, 0] = hubb_u[sp][0], 1
        for sh in range(1, n_shell[sp]):
            same = [k for k in range(n_u[sp]) if abs(hubb_u[sp][sh] - uniq[sp, k]) < MIN_HUB_DIFF]
            if same:
                i_u[sp, sh] = same[0] + 1
            else:
                uniq[sp, n_u[sp]] = hubb_u[sp][sh]
                n_u[sp] += 1
                i_u[sp, sh] = n_u[sp]
    m_u = int(n_u.max())
    return dict(nHubbU=n_u, uniqHubbU=np.ascontiguousarray(uniq[:, :m_u]), iHubbU=i_u, mU=m_u, mShell=m_shell,
                nShell=np.array(n_shell, dtype=np.int32))


def exp_gamma(rab, ua, ub):
    ta, tb = 3.2 * ua, 3.2 * ub
    if rab < TOL_SAME_DIST:
        if abs(ua - ub) < MIN_HUB_DIFF:
            return -0.5 * (ua + ub)
        return -0.5 * ((ta * tb) / (ta + tb) + (ta * tb) ** 2 / (ta + tb) ** 3)
    if abs(ua - ub) < MIN_HUB_DIFF:
        tm = 0.5 * (ta + tb)
        return math.exp(-tm * rab) * (1.0 / rab + 0.6875 * tm + 0.1875 * rab * tm ** 2
                                      + 0.02083333333333333333 * rab ** 2 * tm ** 3)

    def sub(t1, t2):
        return math.exp(-t1 * rab) * ((0.5 * t2 ** 4 * t1 / (t1 ** 2 - t2 ** 2) ** 2)
                                      - (t2 ** 6 - 3.0 * t2 ** 4 * t1 ** 2) / (rab * (t1 ** 2 - t2 ** 2) ** 3))
    return sub(ta, tb) + sub(tb, ta)


def exp_gamma_cutoff(u2, u1, min_value=MIN_SHORT_GAMMA):
    rab = 1.0
    while exp_gamma(rab, u2, u1) > min_value:
        rab *= 2.0
    if rab < 2.0:
        raise RuntimeError("Failure in short-range gamma cut-off: requested tolerance too large")
    lo_cut, hi_cut = rab + 0.1, 0.5 * rab - 0.1          # minCutOff, maxCutOff of the reference
    max_gamma, min_gamma = exp_gamma(hi_cut, u2, u1), exp_gamma(lo_cut, u2, u1)
    lower = min_gamma
    cut = hi_cut + 0.1
    gamma = exp_gamma(cut, u2, u1)
    while gamma - lower > TOL_SHORT_GAMMA:
        hi_cut = 0.5 * (cut + lo_cut)
        if (max_gamma >= min_gamma) == (min_value >= exp_gamma(hi_cut, u2, u1)):
            lo_cut = hi_cut
            lower = exp_gamma(lo_cut, u2, u1)
        else:
            cut = hi_cut
            gamma = exp_gamma(cut, u2, u1)
    return lo_cut


def short_gamma_cutoffs(hub):
    """[sp1, sp2, iU1, iU2] = shortCutoffs(iU2, iU1, iSp2, iSp1) of the reference."""
    n_sp, m_u = len(hub["nHubbU"]), hub["mU"]
    c = np.zeros((n_sp, n_sp, m_u, m_u))
    for s1 in range(n_sp):
        for s2 in range(s1, n_sp):
            for a in range(hub["nHubbU"][s1]):
                for b in range(hub["nHubbU"][s2]):
                    c[s1, s2, a, b] = c[s2, s1, b, a] = exp_gamma_cutoff(hub["uniqHubbU"][s2, b], hub["uniqHubbU"][s1, a])
    return c

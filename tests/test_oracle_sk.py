"""Oracle pinning: the SK interpolation and rotation restatements against the reference's own
arithmetic (machine translation of sk.F90 -> tests/golden/sk_rot_golden.npz, and oracle/_ref when
the reference tree is present)."""
import ctypes
import os

import numpy as np
import pytest

from helpers import oracle, ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "sk_rot_golden.npz"))


@pytest.mark.parametrize("la,lb", [(a, b) for a in range(4) for b in range(a, 4)])
def test_rotation_matches_reference_closed_forms(la, lb):
    """rotateH0 for single-shell species (la on atom 1, lb on atom 2) and the swapped pair."""
    m = oracle.Model([[la], [lb]])
    dirs, sk, hh = GOLD["dirs"], GOLD["sk_%d%d" % (la, lb)], GOLD["hh_%d%d" % (la, lb)]
    for k in range(len(dirs)):
        ref = hh[k].T[:2 * lb + 1, :2 * la + 1]                 # tmpH(1:nOrb2, 1:nOrb1)
        got = oracle.rotate_h0(m, sk[k], dirs[k], 1, 2)
        assert np.abs(got - ref).max() < 1e-14
        if la != lb:                                             # sk.F90:128-133
            got2 = oracle.rotate_h0(m, sk[k], dirs[k], 2, 1)
            assert np.abs(got2 - (-1.0) ** (la + lb) * ref.T).max() < 1e-14


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libskref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_rotation_against_translated_reference_random():
    """Multi-shell species pair (spd / spdf) assembled the way rotateH0 does, fresh random inputs."""
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libskref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    ref.skref_block.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 3 + [dp, dp]
    shells = [[0, 1, 2], [0, 1, 2, 3], [0, 1]]
    m = oracle.Model(shells)
    rng = np.random.default_rng(3)
    for trial in range(30):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        sp1, sp2 = (int(v) for v in rng.integers(1, 4, size=2))
        sk = rng.normal(size=m.n_integrals(sp1, sp2))
        got = oracle.rotate_h0(m, sk, d, sp1, sp2)
        want = np.zeros_like(got)
        ind, col = 0, 0
        for l1 in shells[sp1 - 1]:
            row = 0
            for l2 in shells[sp2 - 1]:
                tmp = np.zeros(49)
                lo, hi = min(l1, l2), max(l1, l2)
                ref.skref_block(lo, hi, d[0], d[1], d[2], sk[ind:ind + lo + 1].ctypes.data_as(dp),
                                tmp.ctypes.data_as(dp))
                t = tmp.reshape(7, 7).T                           # t[row, col] = tmpH(row, col)
                if l1 <= l2:
                    blk = t[:2 * l2 + 1, :2 * l1 + 1]
                else:
                    blk = (-1.0) ** (l1 + l2) * t[:2 * l1 + 1, :2 * l2 + 1].T
                want[row:row + 2 * l2 + 1, col:col + 2 * l1 + 1] = blk
                ind += lo + 1
                row += 2 * l2 + 1
            col += 2 * l1 + 1
        assert np.abs(got - want).max() < 2e-14


def _table(n_grid=60, dist=0.05, n_int=3, seed=0):
    rng = np.random.default_rng(seed)
    r = dist * np.arange(1, n_grid + 1)
    coef = rng.normal(size=(n_int, 6))
    tab = np.stack([np.polyval(c, r) for c in coef], axis=1)
    return np.asfortranarray(tab), coef, dist


def test_interpolation_reproduces_polynomials():
    """8-point Neville is exact for degree <= 7 data (slakoeqgrid.F90:211-222)."""
    tab, coef, dist = _table()
    for r in [0.051, 0.17, 0.5, 1.2345, 2.0, 2.87, 2.999]:
        got = oracle.sk_eval(tab, dist, r)
        want = np.array([np.polyval(c, r) for c in coef])
        assert np.allclose(got, want, rtol=1e-11, atol=1e-11)
    # on a grid point the table value comes back
    assert np.allclose(oracle.sk_eval(tab, dist, 7 * dist), tab[6], rtol=1e-13)


def test_interpolation_tail_and_cutoff():
    """Beyond the grid: 5th-order polynomial to zero over distFudge = 1 bohr, then exactly zero
    (slakoeqgrid.F90:201-206,224-240); continuous at the last grid point."""
    n_grid, dist = 60, 0.05
    r = dist * np.arange(1, n_grid + 1)
    tab = np.asfortranarray(np.stack([np.exp(-r), (1 + r) * np.exp(-1.3 * r)], axis=1))
    r_last = n_grid * dist
    inside = oracle.sk_eval(tab, dist, r_last - 1e-9)
    tail0 = oracle.sk_eval(tab, dist, r_last + 1e-9)
    assert np.allclose(inside, tail0, rtol=1e-6)
    assert np.all(oracle.sk_eval(tab, dist, r_last + 1.0) == 0.0)
    assert np.all(oracle.sk_eval(tab, dist, r_last + 5.0) == 0.0)
    mid = oracle.sk_eval(tab, dist, r_last + 0.5)
    assert np.all(np.abs(mid) < np.abs(tab[-1])) and np.all(mid != 0.0)
    assert abs(oracle.lib().orc_sk_cutoff(n_grid, ctypes.c_double(dist), 2) - (r_last + 1.0)) < 1e-15


def test_old_interpolation_branches():
    """skEqGridOld (slakoeqgrid.F90:248-317): 3-point fit, free cubic spline, tail, zero."""
    tab, coef, dist = _table(n_int=2, seed=4)
    quad = np.asfortranarray(np.stack([1 + 2 * r + 3 * r * r for r in [dist * np.arange(1, 61)]], axis=1))
    got = oracle.sk_eval(quad, dist, 1.234, method=1)
    assert np.allclose(got, 1 + 2 * 1.234 + 3 * 1.234 ** 2, rtol=1e-12)
    r_last = 60 * dist
    a = oracle.sk_eval(quad, dist, r_last - 0.5 * dist, method=1)      # spline interval
    assert np.isfinite(a).all() and abs(a[0] - (1 + 2 * (r_last - 0.5 * dist) + 3 * (r_last - 0.5 * dist) ** 2)) < 1e-2
    assert np.all(oracle.sk_eval(quad, dist, r_last + 0.31, method=1) == 0.0)
    assert np.all(oracle.sk_eval(quad, dist, r_last + 0.1, method=1) != 0.0)


def test_old_interpolation_branches():
    """SlakoEqGrid_interOld_ (slakoeqgrid.F90:248-317): the 3-point polynomial reproduces a
    quadratic exactly, the free cubic spline of the last interval passes through the last two grid
    points, the 5th-order tail starts at the last grid value and ends at zero at (mInd - 1) * dist,
    and the interaction vanishes beyond; getCutoff adds distFudgeOld = 0.3."""
    dist, n = 0.02, 500
    r = dist * np.arange(1, n + 1)
    quad = np.asfortranarray(np.stack([0.3 - 0.2 * r + 0.05 * r * r, np.exp(-0.3 * r)], axis=1))
    for x in (0.031, 0.5, 3.333, 9.9, 9.9799):
        got = oracle.sk_eval(quad, dist, x, method=1)
        assert abs(got[0] - (0.3 - 0.2 * x + 0.05 * x * x)) < 1e-13
        assert abs(got[1] - np.exp(-0.3 * x)) < 1e-6
    assert np.allclose(oracle.sk_eval(quad, dist, 9.98, method=1), quad[-2], rtol=1e-13)
    assert np.allclose(oracle.sk_eval(quad, dist, 10.0 - 1e-12, method=1), quad[-1], rtol=1e-9)
    assert np.allclose(oracle.sk_eval(quad, dist, 10.0, method=1), quad[-1], rtol=1e-12)
    m_ind = n + int(np.floor(0.3 / dist))
    end = (m_ind - 1) * dist
    assert np.abs(oracle.sk_eval(quad, dist, end - 1e-9, method=1)).max() < 1e-20
    assert np.all(oracle.sk_eval(quad, dist, end + 1e-9, method=1) == 0.0)
    assert np.all(oracle.sk_eval(quad, dist, 10.29, method=1) == 0.0)
    # monotone decay through the tail for a monotone table
    tail = np.array([oracle.sk_eval(quad, dist, x, method=1)[1] for x in np.linspace(10.0, end, 15)])
    assert np.all(np.diff(tail) < 0) and tail[-1] >= 0
    assert abs(oracle.lib().orc_sk_cutoff(n, ctypes.c_double(dist), 1) - (n * dist + 0.3)) < 1e-14


def test_interpolation_against_exact_rational_arithmetic():
    """SlakoEqGrid_interNew_ / _interOld_ restated a third time in exact rational arithmetic
    (fractions.Fraction): the unique interpolating polynomial through the reference's window
    (rows iLast-nInter+1..iLast, iLast = max(min(nGrid, ind+nRight), nInter)), the free cubic spline
    and the 5th-order tail as closed forms.  The oracle (floating-point Neville in the reference's
    operation order) must agree to rounding -- independent of how either is coded."""
    from fractions import Fraction as Fr
    rng = np.random.default_rng(12)
    n_grid, dist = 40, 0.25
    tab = np.asfortranarray(rng.normal(size=(n_grid, 2)) * np.exp(-0.1 * np.arange(n_grid))[:, None])
    D = Fr(dist)

    def lagrange(rows, x, col):
        xs = [Fr(j + 1) * D for j in rows]                      # row j (0-based) sits at (j+1)*dist
        ys = [Fr(float(tab[j, col])) for j in rows]
        tot = Fr(0)
        for i, (xi, yi) in enumerate(zip(xs, ys)):
            term = yi
            for k, xk in enumerate(xs):
                if k != i:
                    term *= (x - xk) / (xi - xk)
            tot += term
        return tot

    def poly5(y0, y0p, y0pp, xx, dx):
        dx1, dx2 = y0p * dx, y0pp * dx * dx
        dd, ee, ff = 10 * y0 - 4 * dx1 + dx2 / 2, -15 * y0 + 7 * dx1 - dx2, 6 * y0 - 3 * dx1 + dx2 / 2
        xr = xx / dx
        return ((ff * xr + ee) * xr + dd) * xr ** 3

    def exact(r, col, method):
        x = Fr(r)
        ind = int(x // D)
        if method == 2:
            r_max = n_grid * D + 1
            if x >= r_max:
                return Fr(0)
            if ind < n_grid:
                i_last = max(min(n_grid, ind + 4), 8)
                return lagrange(range(i_last - 8, i_last), x, col)
            return None                                             # tail: FD derivatives, not a closed form
        m_ind = n_grid + int(Fr(3, 10) // D)
        y0, y1, y2 = (Fr(float(tab[n_grid - 3 + k, col])) for k in range(3))
        y1p, y1pp = (y2 - y0) / (2 * D), (y2 + y0 - 2 * y1) / (D * D)
        cub_d = (((y2 - y1) / D - y1p) / D - y1pp / 2) / D
        if ind < n_grid - 1:
            i_last = max(min(n_grid, ind + 2), 3)
            return lagrange(range(i_last - 3, i_last), x, col)
        if ind < n_grid:
            xx = x - (n_grid - 1) * D
            return ((cub_d * xx + y1pp / 2) * xx + y1p) * xx + y1
        if ind < m_ind - 1:
            yp = (3 * cub_d * D + y1pp) * D + y1p
            ypp = 6 * cub_d * D + y1pp
            return poly5(y2, yp, ypp, x - (m_ind - 1) * D, -(m_ind - n_grid - 1) * D)
        return Fr(0)

    checked = 0
    for method in (1, 2):
        for r in (0.26, 0.9, 1.13, 3.77, 8.01, 9.4, 9.74, 9.76, 9.99, 10.01, 10.2, 10.9, 11.01):
            got = oracle.sk_eval(tab, dist, r, method=method)
            for col in range(2):
                ref = exact(r, col, method)
                if ref is None:
                    continue
                scale = np.abs(tab[:, col]).max()
                assert abs(got[col] - float(ref)) <= 5e-13 * scale, (method, r, col, got[col], float(ref))
                checked += 1
    assert checked > 40

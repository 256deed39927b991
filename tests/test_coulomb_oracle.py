"""The oracle's restatement of the SCC Coulomb 1/R matrix (dftb/coulomb.F90: invRCluster, invRPeriodic with
its Ewald set-up, updateShifts) pinned to physics: Madelung constants, independence of the Ewald parameter,
translation invariance, the cluster matrix against numpy."""
import math

import numpy as np
import pytest

from helpers import oracle  # noqa: F401  (builds the C restatement)
from oracle import coulomb_model as cm


def _nacl(a=5.0):
    lat = np.eye(3) * a
    pos = []
    q = []
    for i in range(2):
        for j in range(2):
            for k in range(2):
                pos.append([i * a / 2, j * a / 2, k * a / 2])
                q.append(1.0 if (i + j + k) % 2 == 0 else -1.0)
    return lat, np.array(pos), np.array(q), a / 2


def test_ewald_parameters_follow_the_reference_search():
    lat = np.array([[9.0, 0.3, 0.0], [0.0, 11.0, 0.5], [0.2, 0.0, 13.0]])
    s = cm.ewald_setup(lat)
    inv = np.linalg.inv(lat).T
    rec = 2 * math.pi * inv
    min_g = math.sqrt(min(np.sum(rec ** 2, axis=1)))
    min_r = math.sqrt(min(np.sum(lat ** 2, axis=1)))
    assert abs(cm.diff_rec_real(s["alpha"], min_g, min_r, s["volume"])) <= cm.TOL_EWALD
    # the cutoffs: the reference's bisection stops once the terms at the two ends of the bracket differ by
    # less than the tolerance and returns the last midpoint -- the term there is within one tolerance of it
    assert abs(cm.r_term(s["maxR"], s["alpha"]) - cm.TOL_EWALD) <= cm.TOL_EWALD
    assert abs(cm.g_term(s["maxG"], s["alpha"], s["volume"]) - cm.TOL_EWALD) <= cm.TOL_EWALD
    # reciprocal points: inside the sphere, no origin, exactly one of every +-G pair
    g = s["gLat"]
    assert np.all(np.sum(g ** 2, axis=1) <= s["maxG"] ** 2 * (1 + 1e-12)) and np.all(np.sum(g ** 2, axis=1) > 0)
    keys = {tuple(np.round(v, 9)) for v in g}
    assert len(keys) == len(g) and all(tuple(np.round(-v, 9)) not in keys for v in g)
    # every point of the half space is there: count against a brute-force enumeration
    n = 0
    r = int(s["maxG"] / min_g) + 3
    for a in range(-r, r + 1):
        for b in range(-r, r + 1):
            for c in range(-r, r + 1):
                v = a * rec[0] + b * rec[1] + c * rec[2]
                if (a, b, c) != (0, 0, 0) and np.sum(v ** 2) <= s["maxG"] ** 2:
                    n += 1
    assert n == 2 * len(g)
    assert np.all(s["rLat"][0] == 0.0)


@pytest.mark.parametrize("structure,madelung", [("nacl", 1.7475645946331822), ("cscl", 1.7626747730709)])
def test_madelung_constants(structure, madelung):
    if structure == "nacl":
        lat, pos, q, d = _nacl()
    else:
        a = 4.0
        lat = np.eye(3) * a
        pos = np.array([[0, 0, 0], [a / 2, a / 2, a / 2]])
        q = np.array([1.0, -1.0])
        d = a * math.sqrt(3) / 2
    s = cm.ewald_setup(lat)
    invr = cm.invr_periodic(pos, s)
    assert np.all(np.triu(invr, 1) == 0.0)
    v = cm.coulomb_shift(invr, q)          # potential at every site
    # potential at a cation = -M / d, at an anion +M / d
    assert np.allclose(v * q, -madelung / d, rtol=0, atol=3e-7)
    # independent of alpha (other split between real and reciprocal space)
    s2 = cm.ewald_setup(lat, alpha=1.3 * s["alpha"])
    assert np.abs(cm.invr_periodic(pos, s2) - invr).max() < 3e-7
    # translation of all atoms
    assert np.abs(cm.invr_periodic(pos + [0.31, -0.2, 0.77], s) - invr).max() < 1e-12


def test_cluster_matrix_and_shift():
    rng = np.random.default_rng(4)
    x = rng.uniform(-5, 5, (17, 3))
    a = cm.invr_cluster(x)
    d = np.linalg.norm(x[:, None, :] - x[None, :, :], axis=2)
    ref = np.tril(1.0 / np.where(d > 0, d, np.inf), -1)
    assert np.abs(a - ref).max() < 1e-15
    q = rng.normal(size=17)
    full = ref + ref.T
    assert np.abs(cm.coulomb_shift(a, q) - full @ q).max() < 1e-13


def test_gamma_limits_of_the_oracle():
    """gamma = 1/R - expGamma -> the Hubbard U on site (equal U) and the closed form for different U;
    continuity across the minHubDiff switch."""
    for r in (1e-4, 1e-3):
        assert abs(1.0 / r - cm.exp_gamma(r, 0.3, 0.3) - 0.3) < 1e-5
    ta, tb = 3.2 * 0.3, 3.2 * 0.4
    lim = 0.5 * (ta * tb / (ta + tb) + (ta * tb) ** 2 / (ta + tb) ** 3)
    assert abs(1.0 / 1e-3 - cm.exp_gamma(1e-3, 0.3, 0.4) - lim) < 1e-5
    assert abs(cm.exp_gamma(0.0, 0.3, 0.4) + lim) < 1e-15 and cm.exp_gamma(0.0, 0.3, 0.3) == -0.3
    # the general formula joins the equal-U one (just above minHubDiff it cancels catastrophically -- in the
    # reference too: 1.4e-3 off at U2 - U1 = 4e-6 -- so the join is checked at a separation of 1e-3)
    assert abs(cm.exp_gamma(3.0, 0.3, 0.301) - 0.5 * (cm.exp_gamma(3.0, 0.3, 0.3) + cm.exp_gamma(3.0, 0.301, 0.301))) < 1e-6


def test_product_host_mirror_equals_the_oracle_restatement():
    """dftbplus_b200/coulomb.py (set-up used by bench.py and the harness' device arm) against oracle/coulomb_model.py:
    Ewald parameters, lattice points (same points in the same order), unique Hubbard values, gamma cutoffs."""
    from dftbplus_b200 import coulomb as pc
    for lat in (np.eye(3) * 9.7, np.array([[9.0, 0.3, 0.0], [0.0, 11.0, 0.5], [0.2, 0.0, 13.0]]),
                np.array([[0.0, 5.1, 5.1], [5.1, 0.0, 5.1], [5.1, 5.1, 0.0]])):
        a, b = pc.ewald_setup(lat), cm.ewald_setup(lat)
        for k in ("alpha", "volume", "maxR", "maxG"):
            assert a[k] == b[k], k
        assert a["rLat"].shape == b["rLat"].shape and a["gLat"].shape == b["gLat"].shape
        assert np.abs(a["rLat"] - b["rLat"]).max() < 1e-12 and np.abs(a["gLat"] - b["gLat"]).max() < 1e-14
    hub_u, n_shell = [[0.31, 0.27, 0.27 + 1e-6], [0.42, 0.36]], [3, 2]
    a, b = pc.unique_hubbard(hub_u, n_shell), cm.unique_hubbard(hub_u, n_shell)
    for k in ("nHubbU", "uniqHubbU", "iHubbU"):
        assert np.array_equal(a[k], b[k]), k
    ca, cb = pc.short_gamma_cutoffs(a), cm.short_gamma_cutoffs(b)
    assert np.abs(ca - cb).max() < 1e-9
    for r in (0.0, 0.7, 3.0, 11.0):
        for ua, ub in ((0.3, 0.3), (0.3, 0.41)):
            assert abs(pc.exp_gamma(r, ua, ub) - cm.exp_gamma(r, ua, ub)) <= 1e-13 * max(1.0, abs(cm.exp_gamma(r, ua, ub)))

"""Shared test inputs: workloads prepared once for both the oracle and the CUDA library."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (test infrastructure; never imported by the product)
from oracle import oldskdata  # noqa: E402
from dftbplus_b200 import workloads  # noqa: E402

# H0/S/rho: 1e-12 relative in the max norm plus an elementwise 1e-14 floor
RTOL_SPARSE = 1e-12
ATOL_FLOOR = 1e-14
TOL_FORCE = 1e-7      # Ha/bohr


def assert_sparse_close(got, ref, rtol=RTOL_SPARSE, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(got - ref).max() if ref.size else 0.0
    assert err <= rtol * scale + ATOL_FLOOR, "%s: max err %.3e vs scale %.3e" % (what, err, scale)


def extra_workloads():
    """Small cases beyond BASELINE.json: an f-shell molecule and a tiny periodic cell with many
    self images (every neighbour folds onto a handful of atoms)."""
    rng = np.random.default_rng(11)
    ce = dict(name="ce_o", names=["Ce", "O", "H"],
              coord0=np.array([[0, 0, 0], [3.4, 0.3, -0.2], [-0.5, 3.1, 0.9], [0.4, -0.6, 3.6],
                               [5.0, 1.2, 0.3], [2.2, 2.9, 2.4]]) + rng.normal(scale=0.05, size=(6, 3)),
              species0=np.array([1, 2, 1, 2, 3, 3]), latvec=None, kpoints=None, n_spin=1)
    a = 5.431 / 0.529177249
    si2 = dict(name="si2", names=["Si"], coord0=np.array([[0.0, 0, 0], [0.25 * a, 0.25 * a, 0.25 * a]]),
               species0=np.array([1, 1]),
               latvec=0.5 * a * np.array([[0.0, 1, 1], [1, 0, 1], [1, 1, 0]]),
               kpoints=np.array([[0.0, 0, 0], [0.5, 0, 0], [0.25, 0.25, 0.25], [0.1, -0.3, 0.45]]),
               kweights=np.array([0.125, 0.125, 0.5, 0.25]), n_spin=1)
    # one Fe atom at the origin and neighbours at chosen distances: every branch of both SK
    # interpolation methods (grid 0.02 x 500: polynomial, last-interval cubic spline of the old
    # method, 5th-order tails 10.0-10.28 (old) / 10.0-11.0 (new), beyond the cutoff)
    rs = [1.37, 2.5, 4.013, 7.7, 9.93, 9.965, 9.9850001, 9.995, 10.0000001, 10.013, 10.1, 10.27,
          10.2799, 10.2801, 10.29, 10.5, 10.9, 10.9999, 11.0001]
    pts = [np.zeros(3)]
    for k, r in enumerate(rs):
        u = rng.normal(size=3)
        pts.append(r * u / np.linalg.norm(u))
    scan = dict(name="scan", names=["Fe", "C"], coord0=np.array(pts),
                species0=np.array([1] + [1 + (k % 2) for k in range(len(rs))]), latvec=None, kpoints=None,
                n_spin=1)
    return {"ce_o": ce, "si2": si2, "scan": scan}


@functools.lru_cache(maxsize=None)
def case(name, method=2):
    """(workload, SlakoSet, geometry dict, oracle Model, selfegy) for a named workload.
    method: 2 = skEqGridNew, 1 = skEqGridOld (OldSKInterpolation = Yes)."""
    ex = extra_workloads()
    w = ex[name] if name in ex else workloads.BY_NAME[name]()
    sk, geo, directory = workloads.prepare(w, method=method)
    model, selfegy = oldskdata.load_model(directory, w["names"], sk.ang_shells, method=method)
    return w, sk, geo, model, selfegy


def fermi_fillings(eigen, n_elec, kT=0.001, spin_deg=2.0):
    """Fermi filling by bisection (enough for test inputs; the reference's Efilling is host code)."""
    lo, hi = eigen.min() - 1.0, eigen.max() + 1.0
    for _ in range(200):
        mu = 0.5 * (lo + hi)
        f = spin_deg / (1.0 + np.exp(np.clip((eigen - mu) / kT, -700, 700)))
        if f.sum() > n_elec:
            hi = mu
        else:
            lo = mu
    return f


def n_electrons(w, sk):
    return float(sum(sum(sk.occ[s - 1]) for s in w["species0"]))

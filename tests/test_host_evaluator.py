"""The PRODUCT's evaluator (dftbplus_b200/csrc/sk_lane_eval.cuh + sk_gaunt_code.cuh + the table
preparation of set_sk_table), compiled for the host, against the oracle -- no GPU needed.

The kernels K1 / K5 call exactly these templates per lane; what this cannot cover is the kernel
plumbing around them (schedule, staging, write-out), which the `-m gpu` parity tests cover."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import oracle, ROOT

SRC = os.path.join(ROOT, "tests", "host_eval", "host_eval.cpp")
SO = os.path.join(ROOT, "oracle", "_build", "libhosteval.so")
DEPS = [SRC] + [os.path.join(ROOT, "dftbplus_b200", "csrc", f)
                for f in ("sk_lane_eval.cuh", "sk_gaunt_code.cuh", "sk_table_prep.h", "common.cuh")]


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-I/usr/local/cuda/include", "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    common = [ctypes.c_int, ip, ctypes.c_int, ip, ctypes.c_int, ctypes.c_double, ctypes.c_int, dp, ctypes.c_double, dp,
              ctypes.c_int, dp]
    lib.host_eval_block.argtypes = common
    lib.host_eval_block.restype = ctypes.c_int
    lib.host_eval_wdot.argtypes = common
    lib.host_eval_block_d.argtypes = common
    lib.host_eval_block_d.restype = ctypes.c_int
    lib.host_eval_wdot.restype = ctypes.c_double
    return lib


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _table(rng, n_grid, n_int, dist):
    r = dist * np.arange(1, n_grid + 1)
    amp = rng.uniform(0.1, 0.5, n_int) * rng.choice([-1.0, 1.0], n_int)
    a, b = rng.uniform(0.2, 0.6, n_int), rng.uniform(0.8, 1.4, n_int)
    return np.asfortranarray(amp[None, :] * (1.0 + a[None, :] * r[:, None]) * np.exp(-b[None, :] * r[:, None]))


SHELLS = [([0], [0]), ([0, 1], [0]), ([0], [0, 1, 2]), ([0, 1, 2], [0, 1]), ([0, 1, 2], [0, 1, 2]),
          ([0, 1, 2, 3], [0, 1, 2]), ([0, 1, 2, 3], [0, 1, 2, 3]), ([1, 0], [2]), ([2, 2], [0, 3]), ([3], [1, 1])]
# distances through every branch of both interpolation methods (grid 0.02 x 500)
DISTS = [0.171, 1.37, 4.013, 9.93, 9.965, 9.9850001, 9.995, 10.0000001, 10.1, 10.27, 10.2799, 10.2801, 10.5,
         10.9999, 11.0001]


@pytest.mark.parametrize("method", [1, 2])
def test_blocks_match_the_oracle(method, host):
    rng = np.random.default_rng(100 + method)
    dist, n_grid = 0.02, 500
    worst = 0.0
    for sh1, sh2 in SHELLS:
        m = oracle.Model([sh1, sh2])
        n_int = m.n_integrals(1, 2)
        tab = _table(rng, n_grid, n_int, dist)
        n1, n2 = sum(2 * l + 1 for l in sh1), sum(2 * l + 1 for l in sh2)
        a1, a2 = np.array(sh1, dtype=np.int32), np.array(sh2, dtype=np.int32)
        for r in DISTS:
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            sk = oracle.sk_eval(tab, dist, r, method=method)
            ref = oracle.rotate_h0(m, sk, d, 1, 2)[:n2, :n1]          # rows = species 2
            for generic in (0, 1):
                out = np.full(n1 * n2, np.nan)
                n = host.host_eval_block(len(sh1), _ptr(a1, ctypes.c_int), len(sh2), _ptr(a2, ctypes.c_int), n_grid,
                                         dist, method, _ptr(tab, ctypes.c_double), r, _ptr(d, ctypes.c_double), generic,
                                         _ptr(out, ctypes.c_double))
                assert n == n1 * n2
                got = out.reshape(n1, n2).T                              # column-major nOrb2 x nOrb1
                scale = max(np.abs(ref).max(), 1e-300)
                err = np.abs(got - ref).max()
                # 1e-12 relative to the block (the tail values are tiny); exact zeros beyond the cutoff
                assert err <= 1e-11 * scale + 1e-18, (sh1, sh2, r, generic, err, scale)
                if np.abs(ref).max() == 0.0:
                    assert np.all(got == 0.0)
                worst = max(worst, err / scale if scale > 1e-200 else 0.0)
    assert worst < 1e-11


# poles and axes of the frame construction of the D-matrix evaluator
DIRS = [(0, 0, 1), (0, 0, -1), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0), (1e-9, 0, -1), (1e-160, 1e-160, -1),
        (0.6, 0.0, -0.8), (1, 1, 1), (-1, 2, -3)]


@pytest.mark.parametrize("method", [1, 2])
def test_dmatrix_blocks_match_the_oracle(method, host):
    """K1's compact evaluator (sk_dmat_eval.cuh: frame -> D^1, D^2 -> shell-pair loops; f shells through
    the Gaunt code), every instantiation that can take the species pair, poles of the frame included."""
    rng = np.random.default_rng(300 + method)
    dist, n_grid = 0.02, 500
    worst = 0.0
    for sh1, sh2 in SHELLS + [([0, 0, 1], [0, 1, 2, 2]), ([2, 1, 0, 1], [1])]:
        m = oracle.Model([sh1, sh2])
        n_int = m.n_integrals(1, 2)
        tab = _table(rng, n_grid, n_int, dist)
        n1, n2 = sum(2 * l + 1 for l in sh1), sum(2 * l + 1 for l in sh2)
        a1, a2 = np.array(sh1, dtype=np.int32), np.array(sh2, dtype=np.int32)
        lmax = max(sh1 + sh2)
        dirs = [np.array(d, dtype=np.float64) for d in DIRS] + [rng.normal(size=3) for _ in range(6)]
        for i, d in enumerate(dirs):
            d = d / np.linalg.norm(d)
            r = DISTS[i % len(DISTS)]
            sk = oracle.sk_eval(tab, dist, r, method=method)
            ref = oracle.rotate_h0(m, sk, d, 1, 2)[:n2, :n1]
            for lmax_t in range(max(lmax, 1), 4):
                out = np.full(n1 * n2, np.nan)
                n = host.host_eval_block_d(len(sh1), _ptr(a1, ctypes.c_int), len(sh2), _ptr(a2, ctypes.c_int), n_grid,
                                           dist, method, _ptr(tab, ctypes.c_double), r, _ptr(d, ctypes.c_double), lmax_t,
                                           _ptr(out, ctypes.c_double))
                assert n == n1 * n2
                got = out.reshape(n1, n2).T
                scale = max(np.abs(ref).max(), 1e-300)
                err = np.abs(got - ref).max()
                assert err <= 1e-12 * scale + 1e-18, (sh1, sh2, r, d, lmax_t, err, scale)
                worst = max(worst, err / scale if scale > 1e-200 else 0.0)
    assert worst < 1e-12


def test_weight_contracted_form_equals_direct_contraction(host):
    """sum_e w_e E_e(d) via skLaneWBuild / skLaneWDot (K5) equals the contraction of the block from
    skLaneEval (K1) with the same weights, canonical and generic paths."""
    rng = np.random.default_rng(7)
    dist, n_grid = 0.02, 500
    for sh1, sh2 in SHELLS:
        m = oracle.Model([sh1, sh2])
        tab = _table(rng, n_grid, m.n_integrals(1, 2), dist)
        n1, n2 = sum(2 * l + 1 for l in sh1), sum(2 * l + 1 for l in sh2)
        a1, a2 = np.array(sh1, dtype=np.int32), np.array(sh2, dtype=np.int32)
        for r in (1.37, 5.5, 9.99, 10.4):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            w = rng.normal(size=n1 * n2)
            for generic in (0, 1):
                out = np.zeros(n1 * n2)
                args = (len(sh1), _ptr(a1, ctypes.c_int), len(sh2), _ptr(a2, ctypes.c_int), n_grid, dist, 2,
                        _ptr(tab, ctypes.c_double), r, _ptr(d, ctypes.c_double), generic)
                host.host_eval_block(*args, _ptr(out, ctypes.c_double))
                got = host.host_eval_wdot(*args, _ptr(w, ctypes.c_double))
                ref = float(np.dot(w, out))
                assert abs(got - ref) <= 1e-12 * max(1.0, np.abs(w).max() * np.abs(out).max() * n1 * n2), (sh1, sh2, r)

"""Efilling on the device (dftbp_b200_efilling, SURVEY 8(f) N4 first piece) against the oracle's restatement
of dftb/etemp.F90 (oracle/scc_model.py)."""
import numpy as np
import pytest

from helpers import ROOT  # noqa: F401
from oracle import scc_model as sm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hp():
    from dftbplus_b200.api import HotPath
    h = HotPath(0)
    yield h
    h.close()


def _check(hp, eig_f, n_el, kT, kw, tol=1e-12):
    """eig_f: Fortran-order view [level, k, spin] for the oracle; the device takes [spin][k][level]."""
    ref = sm.efilling(eig_f, n_el, kT, kw)
    got = hp.Efilling(np.ascontiguousarray(eig_f.transpose(2, 1, 0)), n_el, kT, kw)
    fill = got["filling"].transpose(2, 1, 0)
    assert np.abs(fill - ref["filling"]).max() < tol, np.abs(fill - ref["filling"]).max()
    assert abs(got["Ef"] - ref["Ef"]) < tol * max(1.0, abs(ref["Ef"]))
    for key in ("Ebs", "TS", "E0"):
        assert np.abs(got[key] - ref[key]).max() < 10 * tol * max(1.0, np.abs(ref[key]).max()), key
    return got, ref


def test_gapped_integer_electrons_at_min_temp(hp):
    """Mid-gap shortcut: the fillings are a step function, the Fermi level the middle of the gap, exactly."""
    eig = np.array([-0.9, -0.5, -0.45, -0.1, 0.2, 0.3])[:, None, None]
    got, ref = _check(hp, eig, 4.0, sm.MIN_TEMP, [1.0], tol=1e-15)
    assert np.array_equal(got["filling"].ravel(), [1.0, 1.0, 1.0, 1.0, 0.0, 0.0])
    assert got["Ef"] == 0.5 * (-0.1 + 0.2)


@pytest.mark.parametrize("n_el", [7.0, 9.5, 13.25, 23.0])
def test_finite_temperature_kpoints_two_spins(hp, n_el):
    rng = np.random.default_rng(5)
    eig = np.sort(rng.uniform(-1.0, 1.0, size=(12, 3, 2)), axis=0)
    _check(hp, eig, n_el, 0.01, np.array([0.5, 0.25, 0.25]))


def test_large_spectrum_and_device_pointers(hp):
    """n = 31744 levels, 2 spins (the size of BASELINE config 4); eigenvalues and fillings as torch tensors."""
    import torch
    rng = np.random.default_rng(7)
    eig = np.sort(rng.normal(scale=0.4, size=(31744, 1, 2)), axis=0)
    n_el = 28672.0
    ref = sm.efilling(eig, n_el, 0.002, [1.0])
    d = torch.from_numpy(np.ascontiguousarray(eig.transpose(2, 1, 0))).cuda()
    got = hp.Efilling(d, n_el, 0.002, [1.0])
    assert got["filling"].is_cuda
    fill = got["filling"].cpu().numpy().transpose(2, 1, 0)
    assert np.abs(fill - ref["filling"]).max() < 1e-11
    assert abs(fill.sum() - n_el) < 1e-9 and abs(got["Ef"] - ref["Ef"]) < 1e-12
    assert np.abs(got["Ebs"] - ref["Ebs"]).max() < 1e-9 * np.abs(ref["Ebs"]).max()


def test_degenerate_homo(hp):
    eig = np.array([-1.0, 0.0, 0.0, 1.0])[:, None, None]
    got, ref = _check(hp, eig, 2.0, 0.001, [1.0], tol=1e-10)
    f = got["filling"].ravel()
    assert abs(f[1] - f[2]) < 1e-12 and abs(f.sum() - 2.0) < 1e-12


def test_limits_and_unsupported_distribution(hp):
    from dftbplus_b200.api import DftbB200Error
    eig = np.array([-0.3, 0.1])[:, None, None]
    got = hp.Efilling(np.ascontiguousarray(eig.transpose(2, 1, 0)), 0.0, 0.001, [1.0])
    assert got["filling"].sum() == 0.0 and got["Ef"] < -0.3
    got = hp.Efilling(np.ascontiguousarray(eig.transpose(2, 1, 0)), 2.0, 0.001, [1.0])
    assert np.allclose(got["filling"], 1.0) and got["Ef"] > 0.1
    import ctypes
    small = np.zeros(7)
    f = np.zeros(2)
    e = np.array([-0.3, 0.1])
    kw = np.ones(1)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    rc = hp._L.dftbp_b200_efilling(hp._h, p(small[0:2]), p(small[6:7]), p(small[2:4]), p(small[4:6]), p(f), p(e),
                                   2, 1, 1, 1.0, 0.001, p(kw), 2)
    assert rc == 4                                   # DFTBP_B200_ENOTIMPL: Methfessel-Paxton stays on the host

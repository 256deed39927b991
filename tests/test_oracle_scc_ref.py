"""The oracle's scalar SCC-potential functions against the reference's own code: oracle/gen_ref_scc.py machine-translates
expGamma / gammaSubExprn_ / expGammaCutoff (dftb/shortgammafuncs.F90) and gTerm / rTerm / diffRecReal / getMaxGEwald /
getMaxREwald / getOptimalAlphaEwald (dftb/coulomb.F90) into oracle/_ref/libsccref.so; tests/golden/scc_ref_golden.npz holds
its outputs for machines without /root/reference.  Checked: oracle/dftb_oracle.c (orc_exp_gamma, orc_exp_gamma_cutoff),
oracle/coulomb_model.py and the product's host mirror dftbplus_b200/coulomb.py (the Ewald parameter searches)."""
import ctypes
import math
import os
import sys

import numpy as np
import pytest

from helpers import oracle, ROOT  # noqa: F401
from oracle import coulomb_model as cm
from dftbplus_b200 import coulomb as pc

GOLD = os.path.join(ROOT, "tests", "golden", "scc_ref_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_exp_gamma_equals_the_reference_formulas_bit_for_bit(gold):
    got = np.array([cm.exp_gamma(r, a, b) for r, (a, b) in zip(gold["gam_r"], gold["gam_u"])])
    assert np.array_equal(got, gold["gam_out"])
    # the product's host mirror (Python arithmetic: pow instead of products) to rounding
    mir = np.array([pc.exp_gamma(r, a, b) for r, (a, b) in zip(gold["gam_r"], gold["gam_u"])])
    assert np.abs(mir - gold["gam_out"]).max() <= 1e-13 * np.abs(gold["gam_out"]).max()


def test_exp_gamma_cutoff_equals_the_reference_search(gold):
    lib = oracle.lib()
    lib.orc_exp_gamma_cutoff.restype = ctypes.c_double
    lib.orc_exp_gamma_cutoff.argtypes = [ctypes.c_double] * 2
    got = np.array([lib.orc_exp_gamma_cutoff(a, b) for a, b in gold["cut_u"]])
    assert np.array_equal(got, gold["cut_out"])
    mir = np.array([pc.exp_gamma_cutoff(a, b) for a, b in gold["cut_u"]])
    assert np.abs(mir - gold["cut_out"]).max() <= 1e-9


def test_ewald_parameter_searches_equal_the_reference(gold):
    """getOptimalAlphaEwald, getMaxREwald, getMaxGEwald for given (minG, minR, volume): a cubic cell realises
    minR = a, minG = 2 pi / a only for matching values, so the searches are called on the scalars directly."""
    for (min_g, min_r, vol), (alpha, max_r, max_g, diff) in zip(gold["ew"], gold["ew_out"]):
        # oracle/coulomb_model.py
        lat = np.eye(3)
        a = cm.optimal_alpha_ewald(lat * min_r, lat * min_g, vol)
        assert a == alpha
        assert cm.max_r_ewald(a) == max_r and cm.max_g_ewald(a, vol) == max_g
        assert abs(cm.diff_rec_real(a, min_g, min_r, vol) - diff) <= 1e-22
        # dftbplus_b200/coulomb.py
        assert pc.optimal_alpha(lat * min_r, lat * min_g, vol) == alpha
        assert pc._bracket_and_bisect(lambda r: pc._r_term(r, alpha), 1.0e-9, "r") == max_r
        assert pc._bracket_and_bisect(lambda g: pc._g_term(g, alpha, vol), 1.0e-9, "g") == max_g


@pytest.mark.skipif(not os.path.exists("/root/reference/src/dftbp/dftb/shortgammafuncs.F90"), reason="reference tree not present")
def test_golden_vectors_are_what_the_translation_gives_today(gold):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_ref_scc
    lib = gen_ref_scc.load(gen_ref_scc.translate())
    got = np.array([lib.expGamma(r, a, b) for r, (a, b) in zip(gold["gam_r"], gold["gam_u"])])
    assert np.array_equal(got, gold["gam_out"])
    assert np.array_equal(np.array([lib.expGammaCutoff(a, b) for a, b in gold["cut_u"]]), gold["cut_out"])
    assert math.isfinite(float(np.abs(gold["ew_out"]).max()))

"""The oracle's Efilling (oracle/scc_model.py) against vectors made by EXECUTING THE REFERENCE'S OWN SOURCE: Efilling,
electronCount, electronFill, middleGap (dftb/etemp.F90) with kahan (math/summation.F90), run through oracle/fortran_exec.py
by oracle/gen_ref_filling.py -> tests/golden/filling_ref_golden.npz.  Cases: a gapped spectrum at minTemp (mid-gap
shortcut), a degenerate HOMO, a warm metal, k-points with two spin channels and a fractional electron count, the empty and
the full band.  The device routine (dftbp_b200_efilling) is compared with the oracle in tests/test_zx_efilling_gpu.py."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT
from oracle import scc_model

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_filling as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "filling_ref_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", list(G.cases()))
def test_oracle_efilling_equals_the_executed_reference_source(name, gold):
    eig, n_el, kT, kw = G.cases()[name]
    r = scc_model.efilling(eig, n_el, kT, kw)
    assert r["Ef"] == float(gold[name + "__Ef"])                     # same control flow, same bisection steps
    assert np.abs(r["filling"] - gold[name + "__filling"]).max() <= 2.3e-16      # at most the last bit of an occupation
    for k in ("Ebs", "TS", "E0"):
        assert np.abs(np.asarray(r[k]) - gold["%s__%s" % (name, k)]).max() <= 1e-15 * max(1.0, np.abs(gold["%s__%s" % (name, k)]).max()), k


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today(gold):
    f = G.routines()
    for name in ("kpoints_spin", "gap_T0"):
        eig, n_el, kT, kw = G.cases()[name]
        r = G.run(f, eig, n_el, kT, kw)
        for k, v in r.items():
            assert np.array_equal(np.asarray(v), gold["%s__%s" % (name, k)]), (name, k)

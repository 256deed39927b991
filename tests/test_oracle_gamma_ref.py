"""The oracle's short-range gamma (orc_short_gamma_values / orc_short_gamma_shifts) against vectors made by EXECUTING THE
REFERENCE'S OWN SOURCE (oracle/gen_ref_gamma.py: updateNrOfNeighbours_, updateShortGammaValues_, buildShifts_ of
dftb/shortgamma.F90 through oracle/fortran_exec.py; expGamma = its machine translation): shell-resolved and one-U-per-species
Hubbard sets, a molecule, an f-shell cluster and periodic bcc Fe/C."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT
from oracle import coulomb_model as cm

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_gamma as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "gamma_ref_golden.npz")


@pytest.fixture(scope="module")
def cases():
    return G.cases()


@pytest.mark.parametrize("name", ["h2o", "ce_o", "fec128", "fec128r"])
def test_oracle_short_gamma_equals_the_executed_reference_source(name, cases):
    gold = np.load(GOLD)
    nl, hub, cuts, n_shell, dq = cases[name]
    vals = cm.short_gamma_values(nl, hub, cuts)
    assert np.array_equal(vals["nNeighU"], gold[name + "__nNeighU"])
    gam = vals["shortGamma"]
    if name + "__shortGamma_shape" in gold:
        shape = tuple(int(x) for x in gold[name + "__shortGamma_shape"])
        assert np.all(gam[:, shape[1]:] == 0.0)                        # the reference's array ends at maxval(nNeigh)
        got = G.sample(np.ascontiguousarray(gam[:, :shape[1]]))
    else:
        ref_ld = gold[name + "__shortGamma"].shape[1]
        assert np.all(gam[:, ref_ld:] == 0.0)
        got = gam[:, :ref_ld]
    assert np.array_equal(got, gold[name + "__shortGamma"])            # bit for bit: same formulas in the same order
    sh = cm.short_gamma_shifts(nl, hub, vals, dq)
    ref = gold[name + "__shiftShell"]
    assert np.abs(sh - ref).max() <= 1e-15 * max(1.0, np.abs(ref).max())


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today(cases):
    from oracle import oracle as o
    gold = np.load(GOLD)
    R = G.routines(o.lib())
    nl, hub, cuts, n_shell, dq = cases["ce_o"]
    r = G.run(R, nl, hub, cuts, n_shell, dq)
    for k, v in r.items():
        assert np.array_equal(v, gold["ce_o__" + k]), k

"""The oracle's Coulomb 1/R matrix (orc_invr_periodic, orc_invr_cluster) against vectors made by EXECUTING THE REFERENCE'S OWN
SOURCE (oracle/gen_ref_ewald.py): invRPeriodicSerial / addNeighbourContribsIS / ewaldReciprocal / invRClusterSerial of
dftb/coulomb.F90 with the dynamic neighbour list iterator of type/dynneighlist.F90 and type/latpointiter.F90, all through
oracle/fortran_exec.py.  Cases: rock salt, a skewed cell with four atoms, sixteen atoms of the Fe/C cell."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT
from oracle import coulomb_model as cm

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_ewald as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ewald_ref_golden.npz")


@pytest.mark.parametrize("name", ["nacl", "skew", "fec16"])
def test_oracle_ewald_matrix_equals_the_executed_reference_source(name):
    gold = np.load(GOLD)
    c0, lat = G.cases()[name]
    got = cm.invr_periodic(c0, cm.ewald_setup(lat))
    ref = gold[name + "__invr"]
    assert np.all(np.triu(ref, 1) == 0.0) and np.all(np.triu(got, 1) == 0.0)
    assert np.abs(got - ref).max() <= 1e-15 * np.abs(ref).max(), np.abs(got - ref).max()
    assert np.abs(cm.invr_cluster(c0) - gold[name + "__invr_cluster"]).max() <= 1e-16


@pytest.mark.skipif(not os.path.exists(G.REF_C), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today():
    gold = np.load(GOLD)
    per, clu = G.routines()
    c0, lat = G.cases()["skew"]
    assert np.array_equal(G.run_periodic(per, c0, lat, cm.ewald_setup(lat)), gold["skew__invr"])

"""The oracle's neighbour list against vectors made by EXECUTING THE REFERENCE'S OWN SOURCE: oracle/gen_ref_neigh.py runs
updateNeighbourList (dftb/periodic.F90:446-694) through the transpiler of oracle/fortran_exec.py, with the reference's
tolerance heap sort (machine translation) as its index_heap_sort; tests/golden/neigh_ref_golden.npz holds the outputs.
Pinned this way: the numbering of the images, the half-list rule, the order inside shells of equal distance (perfect
lattices), the symmetric list and the pair named by "Atoms i and j too close together" -- i.e. everything the device
neighbour list is compared against bit for bit in the GPU tests."""
import os
import sys

import numpy as np
import pytest

from helpers import oracle, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_neigh as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "neigh_ref_golden.npz")
KEYS = ("coords", "img2CentCell", "iCellVec", "nNeighbour", "iNeighbour", "neighDist2")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def cases():
    return G.cases()


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128", "crystal", "crystal_symmetric"])
def test_oracle_list_equals_the_executed_reference_source(name, gold, cases):
    c0, rcell, cutoff, sym = cases[name]
    species = np.array([1 + i % 3 for i in range(len(c0))], dtype=np.int32)      # 9 / 4 / 1 orbitals: gen_ref_neigh.orbitals_of
    got = oracle.neighbour_list_full([9, 4, 1], c0, species, rcell, cutoff, cutoff, symmetric=sym)
    assert got["nAllAtom"] == int(gold[name + "__nAllAtom"]) and got["mNeigh"] == int(gold[name + "__mNeigh"])
    for k in KEYS:
        ref = gold["%s__%s" % (name, k)]
        assert got[k].shape == ref.shape, k
        assert np.array_equal(got[k], ref), k          # bit for bit, distances and coordinates included
    if not sym:
        # getSparseDescriptor (dftb/sparse2dense.F90:3617-3662), also executed from its source
        assert got["sparseSize"] == int(gold[name + "__sparseSize"])
        assert np.array_equal(got["iPair"], gold[name + "__iPair"])
    if name.startswith("crystal"):
        d = gold[name + "__neighDist2"][:, 1:]
        assert np.any((np.abs(np.diff(d, axis=1)) < 1e-10) & (d[:, 1:] > 0))      # shells of equal distance are present


def test_too_close_pair_is_the_one_the_reference_names(gold, cases):
    c0, rcell, cutoff, sym = cases["too_close"]
    with pytest.raises(RuntimeError) as e:
        oracle.neighbour_list_full([4], c0, np.ones(len(c0), dtype=np.int32), rcell, cutoff, cutoff)
    i, j = (int(v) for v in gold["too_close__error"])
    assert "Atoms %d and %d too close together" % (i, j) in str(e.value)


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not present")
def test_golden_vectors_are_what_the_reference_source_gives_today(gold, cases):
    f = G.routine()
    for name in ("si2", "crystal_symmetric", "too_close"):
        c0, rcell, cutoff, sym = cases[name]
        r = G.run(f, c0, rcell, cutoff, sym)
        for k, v in r.items():
            assert np.array_equal(np.asarray(v), gold["%s__%s" % (name, k)]), (name, k)

"""Known-answer tests of the reference's own unit suite for the partition logic
(test/src/dftbp/unit/common/schedule.F90), run against the oracle restatement and the product's
host mirror; plus TParallelKS (type/parallelks.F90:78-84)."""
import numpy as np
import pytest

from helpers import oracle
from dftbplus_b200 import schedule


def _impls():
    return [("oracle", oracle.chunk_ranges, oracle.indices_with_workload),
            ("host", schedule.get_chunk_ranges, schedule.get_indices_with_workload)]


@pytest.mark.parametrize("name,chunk,_", _impls())
def test_chunk_ranges(name, chunk, _):
    assert chunk(1, 0, 1, 42) == (1, 42)
    assert chunk(1, 0, 7, 42) == (7, 42)
    assert [chunk(3, r, 1, 17) for r in range(3)] == [(1, 6), (7, 12), (13, 17)]
    assert [chunk(3, r, 3, 20) for r in range(3)] == [(3, 8), (9, 14), (15, 20)]


@pytest.mark.parametrize("name,_,idx", _impls())
def test_indices_with_workload(name, _, idx):
    assert list(idx(1, 0, 1, 10, [1] * 10)) == list(range(1, 11))
    assert list(idx(1, 0, 1, 10, [5] * 10)) == list(range(1, 11))
    assert list(idx(1, 0, 1, 10, [0] * 5 + [1] * 5)) == list(range(1, 11))
    assert list(idx(1, 0, 1, 10, [0] * 10)) == list(range(1, 11))
    assert list(idx(1, 0, 7, 10, [-100] * 6 + [1] * 4)) == [7, 8, 9, 10]
    w = [1, 1, 5, 0, 0, 1, 1, 1, 10, 2]
    assert list(idx(2, 0, 1, 10, w)) == [1, 3, 9]
    assert list(idx(2, 1, 1, 10, w)) == [2, 4, 5, 6, 7, 8, 10]
    w = [10, 2, 3, 0, 10, 1, 1, 1, 10, 1]
    assert list(idx(3, 0, 1, 10, w)) == [1, 10]
    assert list(idx(3, 1, 1, 10, w)) == [2, 4, 5]
    assert list(idx(3, 2, 1, 10, w)) == [3, 6, 7, 8, 9]
    w = [100] + [1] * 9
    assert list(idx(3, 0, 1, 10, w)) == [1]
    assert list(idx(3, 1, 1, 10, w)) == [2, 4, 6, 8, 10]
    assert list(idx(3, 2, 1, 10, w)) == [3, 5, 7, 9]
    w = [1] * 9 + [100]
    assert list(idx(3, 0, 1, 10, w)) == [1, 4, 7, 10]
    assert list(idx(3, 1, 1, 10, w)) == [2, 5, 8]
    assert list(idx(3, 2, 1, 10, w)) == [3, 6, 9]


def test_parallel_ks_round_robin():
    # 4 k-points x 2 spins over 3 groups: group = mod((iS-1)*nK + iK-1, nGroup)
    got = [schedule.parallel_ks(3, g, 4, 2) for g in range(3)]
    assert got[0] == [(1, 1), (4, 1), (3, 2)]
    assert got[1] == [(2, 1), (1, 2), (4, 2)]
    assert got[2] == [(3, 1), (2, 2)]
    allks = sorted(sum(got, []))
    assert allks == sorted((k, s) for s in (1, 2) for k in (1, 2, 3, 4))
    assert schedule.parallel_ks(1, 0, 2, 2) == [(1, 1), (2, 1), (1, 2), (2, 2)]


def test_dm_panels_cover_and_balance():
    """Column panels of one density matrix (multi-GPU split beyond (k, spin) sharding): contiguous,
    disjoint, complete, on atom boundaries, and about equal shares of the lower triangle."""
    from dftbplus_b200 import schedule
    rng = np.random.default_rng(4)
    n_orb_atom = rng.choice([1, 4, 9], size=500)
    start = np.concatenate([[1], 1 + np.cumsum(n_orb_atom)])
    n = int(start[-1] - 1)
    for n_panel in (1, 2, 3, 4, 8):
        panels = schedule.dm_panels(start, n_panel)
        assert len(panels) == n_panel
        assert panels[0][0] == 1 and panels[-1][1] == 500 and panels[0][2] == 1 and panels[-1][3] == n
        area = []
        for k, (a0, a1, c0, c1) in enumerate(panels):
            assert c0 == start[a0 - 1] and c1 == start[a1] - 1
            if k:
                assert a0 == panels[k - 1][1] + 1 and c0 == panels[k - 1][3] + 1
            cols = np.arange(c0 - 1, c1)
            area.append(float(np.sum(n - cols)))
        assert max(area) / (sum(area) / n_panel) < 1.05


def test_reference_periodic_unit_tests():
    """Known-answer tests of the reference's own unit suite, test/src/dftbp/unit/dftb/periodic.F90:
    distributeAtoms (singleRank, multipleRanks: same arithmetic as getChunkRanges) and
    neighbourdistance (two atoms at minNeighDist = 1e-2 bohr are accepted, at minNeighDist**2 the
    neighbour list must raise "too close")."""
    from dftbplus_b200 import schedule
    for fn in (schedule.get_chunk_ranges, oracle.chunk_ranges):
        assert tuple(fn(1, 0, 1, 42)) == (1, 42)              # singleRank
        assert tuple(fn(2, 0, 1, 13)) == (1, 7)               # multipleRanks
        assert tuple(fn(2, 1, 1, 13)) == (8, 13)
    m = oracle.Model([[0]])
    min_neigh_dist = 1.0e-2                                     # common/accuracy.F90:69
    ok = oracle.neighbour_geometry(m, np.array([[0.0, 0, 0], [min_neigh_dist, 0, 0]]), np.array([1, 1]), cutoff=1.0)
    assert ok["nNeighbourSK"].tolist() == [0, 1] or ok["nNeighbourSK"].sum() == 1
    with pytest.raises(RuntimeError, match="too close"):
        oracle.neighbour_geometry(m, np.array([[0.0, 0, 0], [min_neigh_dist ** 2, 0, 0]]), np.array([1, 1]), cutoff=1.0)

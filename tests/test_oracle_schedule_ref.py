"""getChunkRanges / getIndicesWithWorkload (common/schedule.F90:230-307) EXECUTED from the reference's source
(oracle/fortran_exec.py) against the oracle's restatement and the product's host mirror (dftbplus_b200/schedule.py): the atom
chunks of the pair build and the workload-balanced atom sets of Mulliken / forces / addShift on several ranks.  Needs the
reference tree (no stored vectors: integer logic, checked where the source is)."""
import os
import sys

import numpy as np
import pytest

from helpers import oracle, ROOT
from dftbplus_b200 import schedule

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402

REF = "/root/reference/src/dftbp/common/schedule.F90"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_partition_logic_equals_the_executed_reference_source():
    src = open(REF).read()
    chunk = fx.compile_routine(src, "subroutine", "getChunkRanges")[0]
    work = fx.compile_routine(src, "subroutine", "getIndicesWithWorkload")[0]
    rng = np.random.default_rng(8)
    for n, size in ((1, 1), (7, 3), (64, 5), (4096, 8), (5, 8)):
        load = rng.integers(0, 40, size=n).astype(np.int64)
        seen = []
        for rank in range(size):
            loc = chunk(size, rank, 1, n, 0, 0)
            lo, hi = int(loc["localFirst"]), int(loc["localLast"])
            assert (lo, hi) == tuple(oracle.chunk_ranges(size, rank, 1, n))
            assert (lo, hi) == tuple(schedule.get_chunk_ranges(size, rank, 1, n))
            idx = work(size, rank, 1, n, fx.FArr(load), None)["indices"].a
            assert np.array_equal(idx, oracle.indices_with_workload(size, rank, 1, n, load.astype(np.int32)))
            assert np.array_equal(idx, schedule.get_indices_with_workload(size, rank, 1, n, load))
            seen.extend(int(i) for i in idx)
        assert sorted(seen) == list(range(1, n + 1))


@pytest.mark.skipif(not os.path.exists("/root/reference/src/dftbp/type/parallelks.F90"), reason="reference tree not present")
def test_kpoint_spin_distribution_equals_the_executed_reference_source():
    """TParallelKS_init (type/parallelks.F90:46-89): the (k-point, spin) pairs of every group, round-robin."""
    import types
    f = fx.compile_routine(open("/root/reference/src/dftbp/type/parallelks.F90").read(), "subroutine", "TParallelKS_init",
                           extra_arrays=("nGroupKS", "groupKS", "localKS"))[0]
    for n_group, n_k, n_s in ((1, 1, 1), (2, 1, 2), (8, 32, 1), (3, 5, 2), (8, 3, 2)):
        seen = []
        for my in range(n_group):
            this = types.SimpleNamespace()
            f(this, types.SimpleNamespace(nGroup=n_group, myGroup=my), n_k, n_s)
            local = [tuple(int(v) for v in col) for col in np.asarray(this.localKS).T] if this.nLocalKS else []
            assert local == schedule.parallel_ks(n_group, my, n_k, n_s)
            seen += local
        assert sorted(seen) == sorted((k, s) for s in range(1, n_s + 1) for k in range(1, n_k + 1))

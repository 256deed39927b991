"""getFullTable (dftbplus/parser.F90:4061-4130: which column of the 20-column .skf tables becomes which Slater-Koster integral
of a species pair, and from which of the two files) EXECUTED from the reference's source (oracle/fortran_exec.py) against the
oracle's restatement (oracle/oldskdata.py) and the product's host code (dftbplus_b200/slako.py).  Needs the reference tree."""
import os
import sys
import types

import numpy as np
import pytest

from helpers import ROOT
from oracle import oldskdata
from dftbplus_b200 import slako

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402

REF = "/root/reference/src/dftbp/dftbplus/parser.F90"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
@pytest.mark.parametrize("sh1,sh2", [([0], [0]), ([0, 1], [0]), ([0], [0, 1, 2]), ([0, 1, 2], [0, 1]), ([0, 1, 2, 3], [0, 1, 2, 3]),
                                      ([1, 0], [2, 0, 1]), ([0, 0, 1], [0, 1, 1])])
def test_full_table_equals_the_executed_reference_source(sh1, sh2):
    def into_array(lst, arr, n_dummy, i_sk):
        vals = lst[int(i_sk) - 1]
        arr.a[:len(vals)] = vals
        return len(vals)

    calls = {"intoArray": lambda a: "%s = intoArray(%s, %s, None, %s)" % (a[2], a[0], a[1], a[3])}
    f = fx.compile_routine(open(REF).read(), "subroutine", "getFullTable", calls, extra_env=dict(intoArray=into_array, maxL=3),
                           extra_arrays=("skHam", "skOver"))[0]
    rng = np.random.default_rng(len(sh1) * 10 + len(sh2))
    n_grid = 7

    def sk_file():
        return types.SimpleNamespace(skHam=fx.FArr(np.asfortranarray(rng.normal(size=(n_grid, 20)))),
                                     skOver=fx.FArr(np.asfortranarray(rng.normal(size=(n_grid, 20)))))

    f12, f21 = sk_file(), sk_file()
    n_int = sum(min(a, b) + 1 for a in sh1 for b in sh2)
    ham = np.full((n_grid, n_int), 7.0, order="F")
    over = np.full((n_grid, n_int), 7.0, order="F")
    d12 = np.empty((1, 1), dtype=object)
    d21 = np.empty((1, 1), dtype=object)
    d12[0, 0], d21[0, 0] = f12, f21
    f(fx.FArr(ham), fx.FArr(over), fx.FArr(d12), fx.FArr(d21), [list(sh1)], [list(sh2)])
    # oracle restatement
    o12 = dict(nGrid=n_grid, skHam=f12.skHam.a, skOver=f12.skOver.a)
    o21 = dict(nGrid=n_grid, skHam=f21.skHam.a, skOver=f21.skOver.a)
    oh, oo = oldskdata.get_full_table(o12, o21, sh1, sh2)
    assert np.array_equal(oh, ham) and np.array_equal(oo, over)
    # product host code
    p12 = types.SimpleNamespace(nGrid=n_grid, skHam=f12.skHam.a, skOver=f12.skOver.a)
    p21 = types.SimpleNamespace(nGrid=n_grid, skHam=f21.skHam.a, skOver=f21.skOver.a)
    ph, po = slako.get_full_table(p12, p21, sh1, sh2)
    assert np.array_equal(ph, ham) and np.array_equal(po, over)

"""The modified Broyden mixer of the oracle (oracle/scc_model.py) and of the product's host mirror (dftbplus_b200/mixer.py)
against iterates made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/gen_ref_mixer.py: modifiedBroydenMixing of
mixer/broydenmixer.F90 through oracle/fortran_exec.py, LAPACK getrf / getrs from SciPy): fourteen iterations of three linear
fixed-point problems, iterate by iterate.  The reference's known-answer unit test is in tests/test_scc_model.py."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT
from oracle import scc_model
from dftbplus_b200.mixer import BroydenMixer as HostMixer

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_ref_mixer as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "mixer_ref_golden.npz")


class _Oracle:
    def __init__(self, n):
        self.m = scc_model.BroydenMixer()
        self.m.reset(n)

    def mix(self, q, diff):
        self.m.mix(q, diff)


class _Host:
    def __init__(self, n):
        self.m = HostMixer()
        self.m.reset(n)

    def mix(self, q, diff):
        q[:] = self.m.mix(q, diff)


@pytest.mark.parametrize("name", ["lin3", "lin11", "lin40"])
@pytest.mark.parametrize("arm", [_Oracle, _Host])
def test_broyden_iterates_equal_the_executed_reference_source(name, arm):
    gold = np.load(GOLD)
    n, d, low, b = G.problems()[name]
    got = G.iterate(arm(n), n, d, low, b)
    ref = gold[name]
    assert got.shape == ref.shape
    for it in range(len(ref)):       # dot products and the small solves round differently (observed <= 5e-14)
        assert np.abs(got[it] - ref[it]).max() <= 1e-12 * max(1.0, np.abs(ref[it]).max()), it
    assert np.array_equal(got[0], ref[0])                 # the first step is the simple mixing, exactly


@pytest.mark.skipif(not os.path.exists(G.REF), reason="reference tree not present")
def test_reference_source_passes_its_own_known_answer_test():
    """test/src/dftbp/unit/mixer/broyden.F90 through the executed source: the transpiler reproduces the reference's numbers."""
    f = G.routine()
    expected = np.array([[6.1761632532030193, 0.91191837339849169, 0.91191837339849169],
                         [6.4798000093100434, 0.76009999534497941, 0.76009999534497941],
                         [6.4625859389302871, 0.76870703053485734, 0.76870703053485745],
                         [6.4635561795150211, 0.76822191024249031, 0.76822191024249042]])
    difference = np.array([[0.88081626601508667, -0.44040813300754278, -0.44040813300754278],
                           [0.55742893386488124, -0.27871446693244140, -0.27871446693244140],
                           [-3.2497351757903914E-002, 1.6248675878950958E-002, 1.6248675878951291E-002],
                           [1.9288771697691942E-003, -9.6443858488526324E-004, -9.6443858488515222E-004]])
    m = G.RefMixer(f, 3, m_iter=100, max_w=100000.0)
    q = np.array([6.0000000000000018, 1.0000000000000002, 1.0000000000000002])
    for ii in range(4):
        m.mix(q, difference[ii])
        assert np.abs(q - expected[ii]).max() <= 100.0 * np.finfo(np.float64).eps

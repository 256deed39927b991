"""The oracle's SCC model (oracle/scc_model.py): the Broyden mixer against the reference's own
known-answer test, the Fermi filling against its defining properties."""
import numpy as np

from helpers import ROOT  # noqa: F401  (puts the repo root on sys.path)
from oracle import scc_model as sm


def test_broyden_reference_known_answer():
    """test/src/dftbp/unit/mixer/broyden.F90 ("broyden_mixer_H20"): charges and charge differences of
    a water molecule, parser defaults, atol = 100 * epsilon."""
    expected = np.array([
        [6.0000000000000018, 1.0000000000000002, 1.0000000000000002],
        [6.1761632532030193, 0.91191837339849169, 0.91191837339849169],
        [6.4798000093100434, 0.76009999534497941, 0.76009999534497941],
        [6.4625859389302871, 0.76870703053485734, 0.76870703053485745],
        [6.4635561795150211, 0.76822191024249031, 0.76822191024249042]])
    difference = np.array([
        [0.88081626601508667, -0.44040813300754278, -0.44040813300754278],
        [0.55742893386488124, -0.27871446693244140, -0.27871446693244140],
        [-3.2497351757903914E-002, 1.6248675878950958E-002, 1.6248675878951291E-002],
        [1.9288771697691942E-003, -9.6443858488526324E-004, -9.6443858488515222E-004]])
    mixer = sm.BroydenMixer(max_scc_iter=100, mix_param=0.2, omega0=0.01, min_weight=1.0, max_weight=100000.0,
                            weight_fac=0.01)
    mixer.reset(3)
    q = expected[0].copy()
    atol = 100.0 * np.finfo(np.float64).eps
    for ii in range(4):
        mixer.mix(q, difference[ii])
        assert np.abs(q - expected[ii + 1]).max() <= atol, (ii, q, expected[ii + 1])


def test_broyden_solves_a_linear_fixed_point():
    """q_out = A q + b: the modified Broyden iteration must reach the fixed point to 1e-12 well inside
    the iteration budget (the Jacobian of a linear map is recovered after n secant pairs)."""
    rng = np.random.default_rng(3)
    n = 6
    A = 0.4 * rng.normal(size=(n, n)) / np.sqrt(n)
    b = rng.normal(size=n)
    fix = np.linalg.solve(np.eye(n) - A, b)
    mixer = sm.BroydenMixer()
    mixer.reset(n)
    q = np.zeros(n)
    for it in range(60):
        diff = A @ q + b - q
        if np.abs(diff).max() < 1e-12:
            break
        mixer.mix(q, diff)
    assert it < 40 and np.abs(q - fix).max() < 1e-11


def test_efilling_integer_electrons_sit_in_the_gap():
    eig = np.array([-0.9, -0.5, -0.45, -0.1, 0.2, 0.3])
    r = sm.fillings_and_band_energies(eig, 8.0)          # minTemp: a step function
    f = r["filling"][:, 0, 0]
    assert np.array_equal(f, [2.0, 2.0, 2.0, 2.0, 0.0, 0.0])
    assert abs(r["Ef"] - 0.5 * (-0.1 + 0.2)) < 1e-15       # middleGap
    assert abs(r["Ebs"][0] - 2.0 * eig[:4].sum()) < 1e-14 and abs(r["TS"][0]) < 1e-20


def test_efilling_counts_electrons_at_finite_temperature():
    rng = np.random.default_rng(5)
    eig = np.sort(rng.uniform(-1.0, 1.0, size=(12, 3, 2)), axis=0)
    kw = np.array([0.5, 0.25, 0.25])
    for n_el in (7.0, 9.5, 13.25):
        r = sm.efilling(eig, n_el, 0.01, kw)
        total = (r["filling"] * kw[None, :, None]).sum()
        assert abs(total - n_el) < 1e-12
        # Fermi-Dirac at the returned level (up to the final rescaling)
        ref = 1.0 / (1.0 + np.exp((eig - r["Ef"]) / 0.01))
        assert np.abs(r["filling"] - ref).max() < 1e-9
        assert np.all(r["TS"] >= 0.0) and np.allclose(r["E0"], r["Ebs"] - 0.5 * r["TS"])


def test_efilling_degenerate_homo_is_shared():
    """Four electrons (spin-unpolarised: two per level) over levels (-1, 0, 0, 1): the mid-gap guess
    lands on the degenerate pair, the bisection then splits the remaining electrons evenly."""
    eig = np.array([-1.0, 0.0, 0.0, 1.0])
    r = sm.fillings_and_band_energies(eig, 4.0, kT=0.001)
    f = r["filling"][:, 0, 0]
    assert abs(f.sum() - 4.0) < 1e-12 and abs(f[1] - f[2]) < 1e-12 and abs(f[1] - 1.0) < 1e-9


def test_efilling_limits():
    eig = np.array([-0.3, 0.1])[:, None, None]
    assert sm.efilling(eig, 0.0, 0.001, [1.0])["filling"].sum() == 0.0
    full = sm.efilling(eig, 2.0, 0.001, [1.0])             # nOrbs * spin <= nElec
    assert np.allclose(full["filling"], 1.0) and full["Ef"] > 0.1


def test_product_mixer_matches_reference_known_answer_and_oracle():
    """dftbplus_b200/mixer.py (the host-side mirror bench.py's MD section uses) against the same reference
    vector and, on a longer sequence, against the oracle's restatement."""
    from dftbplus_b200.mixer import BroydenMixer
    expected = np.array([
        [6.0000000000000018, 1.0000000000000002, 1.0000000000000002],
        [6.1761632532030193, 0.91191837339849169, 0.91191837339849169],
        [6.4798000093100434, 0.76009999534497941, 0.76009999534497941],
        [6.4625859389302871, 0.76870703053485734, 0.76870703053485745],
        [6.4635561795150211, 0.76822191024249031, 0.76822191024249042]])
    difference = np.array([
        [0.88081626601508667, -0.44040813300754278, -0.44040813300754278],
        [0.55742893386488124, -0.27871446693244140, -0.27871446693244140],
        [-3.2497351757903914E-002, 1.6248675878950958E-002, 1.6248675878951291E-002],
        [1.9288771697691942E-003, -9.6443858488526324E-004, -9.6443858488515222E-004]])
    mixer = BroydenMixer(100, 0.2, 0.01, 1.0, 100000.0, 0.01)
    mixer.reset(3)
    q = expected[0].copy()
    for ii in range(4):
        q = mixer.mix(q, difference[ii])
        assert np.abs(q - expected[ii + 1]).max() <= 100.0 * np.finfo(np.float64).eps
    rng = np.random.default_rng(9)
    n = 40
    A = 0.5 * rng.normal(size=(n, n)) / np.sqrt(n)
    b = rng.normal(size=n)
    a, o = BroydenMixer(), sm.BroydenMixer()
    a.reset(n)
    o.reset(n)
    qa, qo = np.zeros(n), np.zeros(n)
    for _ in range(25):
        qa = a.mix(qa, A @ qa + b - qa)
        o.mix(qo, A @ qo + b - qo)
        assert np.abs(qa - qo).max() < 1e-12

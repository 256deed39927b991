"""SURVEY 8(f) N4, second piece: the Coulomb 1/R matrix (cluster and Ewald) and the shift product on the
device, through the C ABI, against the oracle restatement (oracle/dftb_oracle.c orc_invr_*, pinned to the
Madelung constants in test_coulomb_oracle.py)."""
import math

import numpy as np
import pytest

from helpers import oracle, case  # noqa: F401
from oracle import coulomb_model as cm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hp():
    from dftbplus_b200.api import HotPath
    h = HotPath(0)
    yield h
    h.close()


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("n", [3, 64, 257, 1003])
def test_invr_cluster(n, hp):
    rng = np.random.default_rng(n)
    x = rng.uniform(-20, 20, (n, 3))
    got = hp.invRCluster(x)
    ref = cm.invr_cluster(x)
    assert np.all(np.triu(got) == 0.0)
    assert _rel(got, ref) <= 2e-16      # one division and one square root, same operation order


@pytest.mark.parametrize("name,take", [("fec128", None), ("fec128", 127), ("si512", None), ("water1002", None),
                                       ("water1002", 333)])
def test_invr_periodic_equals_the_oracle(name, take, hp):
    """Even atom counts run the reciprocal part through the TMA-fed Gram kernel, odd ones through the
    cp.async kernel; the real-space sums have the reference's order of terms."""
    w, sk, geo, model, selfegy = case(name)
    coord = w["coord0"] if take is None else w["coord0"][:take]
    s = cm.ewald_setup(w["latvec"])
    ref = cm.invr_periodic(coord, s)
    got = hp.invRPeriodic(coord, s["rLat"], s["gLat"], s["alpha"], s["volume"], s["maxR"])
    assert np.all(np.triu(got, 1) == 0.0)
    assert _rel(got, ref) <= 1e-12
    # shifts: the symmetric product with charges that sum to zero
    rng = np.random.default_rng(7)
    dq = rng.normal(size=len(coord))
    dq -= dq.mean()
    sh = hp.coulombShift(got, dq)
    assert _rel(sh, cm.coulomb_shift(ref, dq)) <= 1e-12


@pytest.mark.parametrize("structure,madelung", [("nacl", 1.7475645946331822), ("cscl", 1.7626747730709)])
def test_madelung_constants_on_the_device(structure, madelung, hp):
    a = 5.0 if structure == "nacl" else 4.0
    lat = np.eye(3) * a
    if structure == "nacl":
        pos = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=np.float64) * a / 2
        q = np.array([1.0 if (i + j + k) % 2 == 0 else -1.0 for i in range(2) for j in range(2) for k in range(2)])
        d = a / 2
    else:
        pos = np.array([[0, 0, 0], [a / 2, a / 2, a / 2]])
        q = np.array([1.0, -1.0])
        d = a * math.sqrt(3) / 2
    s = cm.ewald_setup(lat)
    invr = hp.invRPeriodic(pos, s["rLat"], s["gLat"], s["alpha"], s["volume"], s["maxR"])
    v = hp.coulombShift(invr, q)
    assert np.allclose(v * q, -madelung / d, rtol=0, atol=3e-7)


def test_device_resident_chain(hp):
    """Device pointers: the matrix is built and multiplied without leaving the GPU; a padded leading dimension."""
    import torch
    w, sk, geo, model, selfegy = case("si512")
    s = cm.ewald_setup(w["latvec"])
    n = len(w["coord0"])
    ld = n + 6
    hp.use_torch_stream()
    coord = torch.from_numpy(np.ascontiguousarray(w["coord0"])).cuda()
    invr = torch.full((n, ld), 7.0, dtype=torch.float64, device="cuda").t()   # column-major [ld, n] storage, junk inside
    hp.invRPeriodic(coord, s["rLat"], s["gLat"], s["alpha"], s["volume"], s["maxR"], invr=invr)
    dq = torch.from_numpy(np.random.default_rng(3).normal(size=n)).cuda()
    sh = hp.coulombShift(invr, dq)
    hp.synchronize()
    ref = cm.invr_periodic(w["coord0"], s)
    got = invr.cpu().numpy()[:n, :]
    assert _rel(got, ref) <= 1e-12
    assert np.all(invr.cpu().numpy()[n:, :] == 0.0)
    assert _rel(sh.cpu().numpy(), cm.coulomb_shift(ref, dq.cpu().numpy())) <= 1e-12
    hp.set_stream(None)


# ---- short-range gamma (third piece) ------------------------------------------------------------------------
def _hubbard_for(sk, resolved):
    """Synthetic Hubbard U per shell: shell-resolved (distinct values, two shells of one species sharing one) or one
    value per species (the reference's default, orbital-resolved SCC off)."""
    n_shell = [len(a) for a in sk.ang_shells]
    if resolved:
        base = [[0.31, 0.27, 0.27, 0.22], [0.42, 0.36, 0.30, 0.25], [0.5, 0.45, 0.4, 0.35]]
    else:
        base = [[0.31] * 4, [0.42] * 4, [0.5] * 4]
    return [base[i % 3][:n] for i, n in enumerate(n_shell)], n_shell


@pytest.mark.parametrize("name,resolved", [("h2o", True), ("ce_o", True), ("fec128", True), ("fec128", False),
                                           ("si512", False), ("water1002", True)])
def test_short_gamma_on_the_device_list(name, resolved):
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = case(name)
    hub_u, n_shell = _hubbard_for(sk, resolved)
    hub = cm.unique_hubbard(hub_u, n_shell)
    cut = cm.short_gamma_cutoffs(hub)
    cutoff = max(sk.cutoff(), float(cut.max()))          # mCutOff: the list covers every short-range cutoff
    rcell = np.zeros((1, 3)) if w["latvec"] is None else oracle.cell_translations(w["latvec"], cutoff)[1]
    nl = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, sk.cutoff())
    ref = cm.short_gamma_values(nl, hub, cut)
    hp = HotPath(0)
    sk.upload(hp)
    g = hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], sk_cutoff=sk.cutoff(), adopt=True)
    assert g["mNeigh"] == nl["mNeigh"] and np.array_equal(g["iNeighbour"], nl["iNeighbour"])
    got = hp.shortGammaUpdate(hub["nHubbU"], hub["uniqHubbU"], cut, fetch=True)
    assert np.array_equal(got["nNeighU"], ref["nNeighU"]) and np.array_equal(got["maxNeigh"], ref["maxNeigh"])
    scale = np.abs(ref["shortGamma"]).max()
    assert np.abs(got["shortGamma"] - ref["shortGamma"]).max() <= 1e-13 * scale
    assert int(ref["maxNeigh"].max()) > 0 or nl["nAtom"] == 1
    rng = np.random.default_rng(11)
    dq = rng.normal(size=(nl["nAtom"], hub["mShell"])) * 0.1
    for a in range(nl["nAtom"]):                          # no charge on shells a species does not have
        dq[a, n_shell[w["species0"][a] - 1]:] = 0.0
    sh = hp.shortGammaShifts(hub["iHubbU"], dq)
    sh_ref = cm.short_gamma_shifts(nl, hub, ref, dq)
    assert np.abs(sh - sh_ref).max() <= 1e-12 * max(1.0, np.abs(sh_ref).max())
    # a new list invalidates the values until updateCoords is repeated
    from dftbplus_b200.api import DftbB200Error
    hp.updateNeighbourListAndSpecies(w["coord0"][: max(1, nl["nAtom"] - 1)], w["species0"][: max(1, nl["nAtom"] - 1)],
                                     cutoff, w["latvec"], sk_cutoff=sk.cutoff(), adopt=False, fetch=False)
    if nl["nAtom"] > 1:
        with pytest.raises(DftbB200Error):
            hp.shortGammaShifts(hub["iHubbU"], dq)
    hp.close()


# ---- Broyden mixer with device-resident vectors ------------------------------------------------------------
def test_device_broyden_mixer_reference_known_answer(hp):
    """test/src/dftbp/unit/mixer/broyden.F90 ("broyden_mixer_H20"): the reference's own vectors, atol 100 eps."""
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
    hp.mixerInit(100, 0.2, 0.01, 1.0, 100000.0, 0.01)
    hp.mixerReset(3)
    q = expected[0].copy()
    for ii in range(4):
        hp.mixerMix(q, difference[ii])
        assert np.abs(q - expected[ii + 1]).max() <= 100.0 * np.finfo(np.float64).eps, (ii, q, expected[ii + 1])


@pytest.mark.parametrize("n", [7, 1000, 4096])
def test_device_broyden_mixer_follows_the_oracle(n, hp):
    """A linear fixed-point problem q_out = A q + b driven by the device mixer (device-resident vectors) and by the
    oracle restatement side by side: the same iterates to 1e-11 of the charge scale, iteration by iteration, the
    same number of iterations to convergence; too many calls after a reset are refused like in the reference."""
    import torch
    from oracle import scc_model
    from dftbplus_b200.api import DftbB200Error
    rng = np.random.default_rng(n)
    d = rng.uniform(-0.5, 0.5, n)
    lowrank = rng.normal(size=(n, 3)) / np.sqrt(n)
    b = rng.normal(size=n)

    def out(q):
        return d * q + 0.3 * lowrank @ (lowrank.T @ q) + b

    ref = scc_model.BroydenMixer(max_scc_iter=40)
    ref.reset(n)
    hp.use_torch_stream()
    hp.mixerInit(40)
    hp.mixerReset(n)
    q_ref = np.zeros(n)
    q_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
    its_ref = its_dev = None
    for it in range(1, 41):
        diff_ref = out(q_ref) - q_ref
        q_host = q_dev.cpu().numpy()
        diff_dev = out(q_host) - q_host
        if its_ref is None and np.abs(diff_ref).max() < 1e-10:
            its_ref = it
        if its_dev is None and np.abs(diff_dev).max() < 1e-10:
            its_dev = it
        if its_ref is not None and its_dev is not None:
            break
        ref.mix(q_ref, diff_ref)
        hp.mixerMix(q_dev, torch.from_numpy(diff_dev).cuda())
        hp.synchronize()
        assert np.abs(q_dev.cpu().numpy() - q_ref).max() <= 1e-11 * max(1.0, np.abs(q_ref).max()), it
    assert its_ref is not None and its_ref == its_dev
    hp.mixerReset(n)
    for _ in range(40):
        hp.mixerMix(q_dev, torch.zeros(n, dtype=torch.float64, device="cuda") + 1e-3)
    with pytest.raises(DftbB200Error):
        hp.mixerMix(q_dev, torch.zeros(n, dtype=torch.float64, device="cuda"))
    hp.set_stream(None)

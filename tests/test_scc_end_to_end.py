"""End-to-end SCC single point (BASELINE config 1 and two larger clusters): the device arm against
the oracle arm with a shared host-side SCC model (tests/scc_harness.py).

north_star bars: total energy within 1e-8 Ha, forces within 1e-7 Ha/bohr, IDENTICAL SCC iteration
counts; Mulliken charges within 1e-9."""
import numpy as np
import pytest

from helpers import oracle, case
from oracle import oldskdata
from dftbplus_b200 import workloads
import scc_harness as sh

TOL_ENERGY = 1e-8    # Ha
TOL_FORCE = 1e-7     # Ha / bohr


def _setup(name):
    if name == "watercluster24":
        w = workloads.water_box(8, 1)
        w["latvec"], w["name"] = None, name
        sk, geo, directory = workloads.prepare(w)
        model, selfegy = oldskdata.load_model(directory, w["names"], sk.ang_shells)
        return w, sk, geo, model, selfegy
    return case(name)


MIX = {"h2o": 0.4, "watercluster24": 0.1, "ce_o": 0.05}


def test_oracle_scc_converges_and_is_variational():
    """CPU: the oracle arm converges, conserves the electron count, and its SCC forces are the
    derivative of nothing less than the SCC-consistent energy only up to the gamma' term that the
    harness leaves out -- so check what must hold exactly: sum of forces = 0 and q sums to N."""
    w, sk, geo, model, selfegy = _setup("h2o")
    r = sh.scc_single_point(sh.OracleArm(oracle, model, geo, selfegy), w, sk, geo, model.mOrb, mix=MIX["h2o"])
    assert r["err"] < 1e-8 and 5 < r["n_iter"] < 100
    assert abs(r["q"].sum() - 8.0) < 1e-10
    assert np.abs(r["forces"].sum(axis=0)).max() < 1e-10
    assert r["e_scc"] > 0.0


def test_oracle_scc_reference_defaults_converge():
    """CPU: the same single point with the reference's default SCC machinery (Broyden mixer on the
    atom charges, Efilling at minTemp; oracle/scc_model.py) reaches the same fixed point as the
    simple-mixing loop, in far fewer iterations."""
    w, sk, geo, model, selfegy = _setup("h2o")
    arm = sh.OracleArm(oracle, model, geo, selfegy)
    slow = sh.scc_single_point(arm, w, sk, geo, model.mOrb, mix=MIX["h2o"], kT=1e-8, tol=1e-10, max_iter=400)
    fast = sh.scc_single_point(arm, w, sk, geo, model.mOrb, reference_defaults=True, tol=1e-10)
    assert fast["err"] < 1e-10 and fast["n_iter"] < slow["n_iter"] and fast["n_iter"] <= 20
    assert abs(fast["energy"] - slow["energy"]) < 1e-9
    assert np.abs(fast["q"].sum(axis=1) - slow["q"].sum(axis=1)).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h2o", "watercluster24", "ce_o"])
def test_scc_reference_defaults_device_vs_oracle(name):
    """The SCC loop with the reference's defaults (Broyden, Efilling; see scc_harness): identical
    iteration counts, energy 1e-8 Ha, forces 1e-7 Ha/bohr between the oracle arm and the device arm."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = _setup(name)
    ref = sh.scc_single_point(sh.OracleArm(oracle, model, geo, selfegy), w, sk, geo, model.mOrb,
                              reference_defaults=True, tol=1e-9, max_iter=100)
    hp = HotPath(0)
    try:
        sk.upload(hp)
        hp.set_geometry(geo)
        got = sh.scc_single_point(sh.DeviceArm(hp, sk, geo), w, sk, geo, model.mOrb, reference_defaults=True,
                                  tol=1e-9, max_iter=100)
    finally:
        hp.close()
    assert ref["err"] < 1e-9, "oracle arm did not converge"
    assert got["n_iter"] == ref["n_iter"], (got["n_iter"], ref["n_iter"])
    assert abs(got["energy"] - ref["energy"]) < TOL_ENERGY, (got["energy"], ref["energy"])
    assert np.abs(got["q"] - ref["q"]).max() < 1e-9
    assert np.abs(got["forces"] - ref["forces"]).max() < TOL_FORCE


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h2o", "watercluster24", "ce_o"])
def test_scc_single_point_device_vs_oracle(name):
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = _setup(name)
    ref = sh.scc_single_point(sh.OracleArm(oracle, model, geo, selfegy), w, sk, geo, model.mOrb, mix=MIX[name],
                              max_iter=600)
    hp = HotPath(0)
    try:
        sk.upload(hp)
        hp.set_geometry(geo)
        got = sh.scc_single_point(sh.DeviceArm(hp, sk, geo), w, sk, geo, model.mOrb, mix=MIX[name], max_iter=600)
        assert hp.kernel_launches > 5 * got["n_iter"]          # the CUDA path ran, every iteration
    finally:
        hp.close()
    assert ref["err"] < 1e-8, "oracle arm did not converge"
    assert got["n_iter"] == ref["n_iter"], (got["n_iter"], ref["n_iter"])
    assert abs(got["energy"] - ref["energy"]) < TOL_ENERGY, (got["energy"], ref["energy"])
    assert abs(got["e_h0"] - ref["e_h0"]) < TOL_ENERGY and abs(got["e_scc"] - ref["e_scc"]) < TOL_ENERGY
    assert np.abs(got["q"] - ref["q"]).max() < 1e-9
    assert np.abs(got["forces"] - ref["forces"]).max() < TOL_FORCE


def _oracle_arm_factory(model, selfegy, species0):
    def make(x):
        geo = oracle.neighbour_geometry(model, x, species0, None)
        return sh.OracleArm(oracle, model, geo, selfegy), geo
    return make


def test_oracle_md_steps_restart_and_momentum():
    """CPU: the MD driver of the config-5 pattern on the oracle arm -- the SCC restart converges in
    fewer iterations than the cold start, the list is rebuilt per step, total momentum stays zero."""
    w, sk, geo, model, selfegy = _setup("watercluster24")
    r = sh.velocity_verlet(_oracle_arm_factory(model, selfegy, w["species0"]), w, sk, model.mOrb, n_steps=2)
    its = [s["n_iter"] for s in r["steps"]]
    assert its[1] < its[0] and its[2] < its[0]
    mass = np.array([sk.mass[s - 1] for s in w["species0"]])[:, None]
    assert np.abs((mass * r["v"]).sum(axis=0)).max() < 1e-9 * mass.mean()
    assert np.abs(r["forces"].sum(axis=0)).max() < 1e-9


# ---- the reference's SCC potential: 1/R (cluster / Ewald) - short-range gamma -------------------
def _hub(sk):
    """One Hubbard U per species (OrbitalResolvedSCC = No: every shell takes the U of the first, as the parser
    sets hubbU(2:, :) = hubbU(1, :))."""
    from oracle import coulomb_model as cm
    return cm.unique_hubbard([[u[0]] * len(u) for u in sk.hubbU], [len(a) for a in sk.ang_shells])


def test_oracle_scc_with_ewald_potential_converges():
    """CPU: a periodic SCC single point (bcc Fe/C, 128 atoms) with the Ewald 1/R matrix and the short-range
    gamma of the restatement: converges with the reference's mixer, conserves the electrons; a neutral charge
    pattern feels no net potential from the uniform background term."""
    w, sk, geo, model, selfegy = case("fec128")
    gam = sh.OracleGamma(oracle, sk, w, _hub(sk))
    r = sh.scc_reference_potential(sh.OracleArm(oracle, model, geo, selfegy), gam, w, sk, geo, model.mOrb, tol=1e-8,
                                   max_iter=150)
    assert r["err"] < 1e-8 and r["n_iter"] < 150          # a metal at minTemp: 69 iterations with the Broyden mixer
    n_elec = sum(sum(sk.occ[s - 1]) for s in w["species0"])
    assert abs(r["q"].sum() - n_elec) < 1e-8
    # cluster case through the same driver
    w, sk, geo, model, selfegy = case("h2o")
    gam = sh.OracleGamma(oracle, sk, w, _hub(sk))
    r2 = sh.scc_reference_potential(sh.OracleArm(oracle, model, geo, selfegy), gam, w, sk, geo, model.mOrb, tol=1e-9)
    # one U per species and no periodic images: the potential equals the closed-form cluster gamma of round 1
    r1 = sh.scc_single_point(sh.OracleArm(oracle, model, geo, selfegy), w, sk, geo, model.mOrb, reference_defaults=True,
                             tol=1e-9)
    assert r2["n_iter"] == r1["n_iter"] and abs(r2["energy"] - r1["energy"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h2o", "fec128", "watercluster24"])
def test_scc_with_reference_potential_device_vs_oracle(name):
    """Every SCC piece that has a device routine runs on the device: neighbour list (+ adopt), H0 / S, 1/R matrix
    (cluster or Ewald), short-range gamma, shifts, addShift_block, eigensolve, density matrix, pack, Mulliken,
    forces; against the oracle arm: identical iteration counts, energy 1e-8 Ha, charges 1e-9, forces 1e-7."""
    from dftbplus_b200.api import HotPath
    w, sk, geo, model, selfegy = _setup(name)
    hub = _hub(sk)
    ref = sh.scc_reference_potential(sh.OracleArm(oracle, model, geo, selfegy), sh.OracleGamma(oracle, sk, w, hub), w, sk,
                                     geo, model.mOrb, tol=1e-9, max_iter=150)
    hp = HotPath(0)
    try:
        sk.upload(hp)
        gam = sh.DeviceGamma(hp, sk, w, hub)          # builds and adopts the list: geometry of the pair kernels
        assert gam.geo["sparseSize"] == geo["sparseSize"]
        got = sh.scc_reference_potential(sh.DeviceArm(hp, sk, geo), gam, w, sk, geo, model.mOrb, tol=1e-9, max_iter=150)
    finally:
        hp.close()
    assert ref["err"] < 1e-9, "oracle arm did not converge"
    assert got["n_iter"] == ref["n_iter"], (got["n_iter"], ref["n_iter"])
    assert abs(got["energy"] - ref["energy"]) < TOL_ENERGY, (got["energy"], ref["energy"])
    assert np.abs(got["q"] - ref["q"]).max() < 1e-9
    assert np.abs(got["forces"] - ref["forces"]).max() < TOL_FORCE

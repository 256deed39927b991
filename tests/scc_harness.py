"""Host-side SCC single-point driver used by the end-to-end parity test (BASELINE config 1: "H2O
molecule SCC single point"; north_star: total energy within 1e-8 Ha, identical SCC iteration
counts, forces within 1e-7 Ha/bohr).

Everything that SURVEY.md marks OUT OF SCOPE stays on the host and is SHARED by the two arms: the
atom-resolved short-range gamma (the closed forms behind dftb/shortgamma.F90 / shortgammafuncs.F90:
gamma_ij = 1/R - S(tau_i, tau_j, R), tau = 16/5 U, gamma_ii = U_i; cluster boundary conditions, so
no Ewald sum), Fermi filling, simple charge mixing and the convergence test on the Mulliken charges
(dftbplus/main.F90:4884-4930).  What differs between the arms is only the hot path: the oracle arm
calls the CPU restatement (+ scipy.linalg.eigh), the device arm calls the CUDA library through
dftbplus_b200.api (+ cuSOLVER) in the order of processGeometry (dftbplus/main.F90:1219-1221, 1449,
3536-3566, 4101-4108, 4705, 6572-6691, 7264): buildH0/buildS -> [addShift -> unpackHS -> diag ->
getDensityMatrix -> packHS -> mulliken]* -> getEDensityMatrix -> packHS -> derivative_shift."""
import numpy as np
import scipy.linalg

from oracle import scc_model


def short_gamma(coords, species0, hubb_u):
    """gamma[nAtom, nAtom] for a cluster (Elstner et al. 1998, eqs. (12)-(14))."""
    n = len(species0)
    tau = np.array([3.2 * hubb_u[s - 1] for s in species0])
    g = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i == j:
                g[i, j] = hubb_u[species0[i] - 1]
                continue
            r = np.linalg.norm(coords[i] - coords[j])
            a, b = tau[i], tau[j]
            if abs(a - b) < 1e-12:
                s = np.exp(-a * r) * (48.0 + 33.0 * a * r + 9.0 * (a * r) ** 2 + (a * r) ** 3) / (48.0 * r)
            else:
                s = (np.exp(-a * r) * (b ** 4 * a / (2.0 * (a * a - b * b) ** 2)
                                       - (b ** 6 - 3.0 * b ** 4 * a * a) / ((a * a - b * b) ** 3 * r))
                     + np.exp(-b * r) * (a ** 4 * b / (2.0 * (b * b - a * a) ** 2)
                                         - (a ** 6 - 3.0 * a ** 4 * b * b) / ((b * b - a * a) ** 3 * r)))
            g[i, j] = 1.0 / r - s
    return g


def fermi(eigen, n_elec, kT=0.001):
    lo, hi = eigen.min() - 1.0, eigen.max() + 1.0
    for _ in range(200):
        mu = 0.5 * (lo + hi)
        f = 2.0 / (1.0 + np.exp(np.clip((eigen - mu) / kT, -700, 700)))
        if f.sum() > n_elec:
            hi = mu
        else:
            lo = mu
    return f


class OracleArm:
    def __init__(self, oracle, model, geo, selfegy):
        self.o, self.m, self.geo, self.selfegy = oracle, model, geo, selfegy

    def build(self):
        return self.o.build_h0(self.m, self.geo, self.selfegy), self.o.build_s(self.m, self.geo)

    def add_shift(self, h0, over, shift_atom):
        ham = self.o.add_shift(self.m, self.geo, np.zeros((1, self.geo["sparseSize"])), over, shift_atom[None, :],
                               is_input_zero=True)
        ham[0] += h0
        return ham[0]

    def add_shift_block(self, h0, over, shift_block):
        ham = self.o.add_shift(self.m, self.geo, np.zeros((1, self.geo["sparseSize"])), over, shift_block,
                               is_input_zero=True)
        ham[0] += h0
        return ham[0]

    def solve(self, ham, over):
        H, S = self.o.unpack_hs_real(self.geo, ham), self.o.unpack_hs_real(self.geo, over)
        eig, C = scipy.linalg.eigh(H, S, lower=True)
        return eig, np.asfortranarray(C)

    def rho(self, C, f):
        return self.o.pack_hs_real(self.geo, np.zeros(self.geo["sparseSize"]), self.o.density_matrix(C, f))

    def erho(self, C, f, eig):
        return self.o.pack_hs_real(self.geo, np.zeros(self.geo["sparseSize"]), self.o.energy_density_matrix(C, f, eig))

    def mulliken(self, over, rho):
        return self.o.mulliken_per_orbital(self.m, self.geo, over, rho)

    def forces(self, rho, erho, shift_block):
        return self.o.derivative_shift(self.m, self.geo, rho, erho, shift=shift_block)


class DeviceArm:
    """The same calls through the C ABI (host numpy buffers in, host buffers out)."""

    def __init__(self, hp, sk, geo):
        self.hp, self.sk, self.geo = hp, sk, geo

    def build(self):
        return self.hp.buildH0S(self.sk.selfegy)

    def add_shift(self, h0, over, shift_atom):
        ham = np.zeros((1, self.geo["sparseSize"]))
        self.hp.addShift(ham, over, np.ascontiguousarray(shift_atom[None, :]), is_input_zero=True)
        ham[0] += h0
        return ham[0]

    def add_shift_block(self, h0, over, shift_block):
        ham = np.zeros((1, self.geo["sparseSize"]))
        self.hp.addShift(ham, over, np.ascontiguousarray(shift_block), is_input_zero=True)
        ham[0] += h0
        return ham[0]

    def solve(self, ham, over):
        H, S = self.hp.unpackHS(ham), self.hp.unpackHS(over)
        eig = self.hp.diagDenseMtx(H, S)        # H <- eigenvectors
        return eig, H

    def rho(self, C, f):
        return self.hp.packHS(np.zeros(self.geo["sparseSize"]), self.hp.getDensityMatrix(C, f))

    def erho(self, C, f, eig):
        return self.hp.packHS(np.zeros(self.geo["sparseSize"]), self.hp.getEDensityMatrix(C, f, eig))

    def mulliken(self, over, rho):
        return self.hp.mulliken(over, rho)

    def forces(self, rho, erho, shift_block):
        return self.hp.derivative_shift(rho, erho, shift=shift_block)


def scc_single_point(arm, workload, sk, geo, m_orb, tol=1e-8, mix=0.4, max_iter=200, q_init=None,
                     reference_defaults=False, kT=None):
    """Returns dict(n_iter, energy, q, forces).  q_init: input charges of the first iteration (the
    previous geometry step's output: "SCC restart").

    reference_defaults = False: simple mixing of the orbital-resolved charges with parameter `mix`,
    bisection Fermi level at kT = 0.001 (the round-1 model).
    reference_defaults = True: the SCC loop as DFTB+ runs it by default (getNextInputCharges,
    dftbplus/main.F90:4890-4935): the charges are reduced to one value per atom (non-orbital-resolved
    SCC: all orbitals of an atom are equivalent, reduceCharges), the error is max |qOutRed - qInpRed|,
    the reduced charges go through the modified Broyden mixer with the parser's defaults
    (oracle/scc_model.py, pinned to the reference's unit test), and the fillings come from the
    restated Efilling at the reference's default electronic temperature (0 K -> minTemp)."""
    species0 = np.asarray(workload["species0"])
    n_atom = geo["nAtom"]
    hubb = [sk.hubbU[s][0] for s in range(len(sk.hubbU))]
    gamma = short_gamma(np.asarray(workload["coord0"], dtype=float), species0, hubb)
    q0 = np.zeros((n_atom, m_orb))
    n_orb_atom = np.zeros(n_atom, dtype=int)
    for i, s in enumerate(species0):           # reference (neutral atom) orbital populations
        k = 0
        for l, occ in zip(sk.ang_shells[s - 1], sk.occ[s - 1]):
            q0[i, k:k + 2 * l + 1] = occ / (2 * l + 1)
            k += 2 * l + 1
        n_orb_atom[i] = k
    n_elec = q0.sum()
    if kT is None:
        kT = scc_model.MIN_TEMP if reference_defaults else 0.001

    def fill(eig):
        if reference_defaults:
            return scc_model.fillings_and_band_energies(np.asarray(eig), n_elec, kT)["filling"][:, 0, 0]
        return fermi(np.asarray(eig), n_elec, kT)

    def expand(q_red):
        # expandCharges for atom-wise equivalence: the atom's charge spread evenly over its orbitals
        q = np.zeros((n_atom, m_orb))
        for i in range(n_atom):
            q[i, :n_orb_atom[i]] = q_red[i] / n_orb_atom[i]
        return q

    mixer = scc_model.BroydenMixer(max_scc_iter=max(max_iter, 2)) if reference_defaults else None
    if mixer is not None:
        mixer.reset(n_atom)
    h0, over = arm.build()
    q_in = q0.copy() if q_init is None else np.array(q_init, dtype=float, copy=True)
    q_in_red = q_in.sum(axis=1)
    for it in range(1, max_iter + 1):
        dq = (q_in - q0).sum(axis=1)
        shift_atom = gamma @ dq
        ham = arm.add_shift(h0, over, shift_atom)
        eig, C = arm.solve(ham, over)
        f = fill(eig)
        rho = arm.rho(C, f)
        q_out = arm.mulliken(over, rho)
        if reference_defaults:
            q_diff_red = q_out.sum(axis=1) - q_in_red
            err = np.abs(q_diff_red).max()
            if err < tol:
                break
            mixer.mix(q_in_red, q_diff_red)
            q_in = expand(q_in_red)
        else:
            err = np.abs(q_out - q_in).max()
            if err < tol:
                break
            q_in = q_in + mix * (q_out - q_in)
    dq = (q_out - q0).sum(axis=1)
    e_h0 = arm.mulliken(h0, rho).sum()                     # Tr[rho H0], getenergies.F90:198
    e_scc = 0.5 * dq @ gamma @ dq
    erho = arm.erho(C, f, np.asarray(eig))
    shift_block = np.zeros((1, n_atom, m_orb, m_orb))
    idx = np.arange(m_orb)
    shift_block[0, :, idx, idx] = (gamma @ dq)[None, :]
    forces = arm.forces(rho, erho, shift_block)
    return dict(n_iter=it, energy=e_h0 + e_scc, e_h0=e_h0, e_scc=e_scc, q=q_out, forces=forces, err=err)


def velocity_verlet(make_arm, workload, sk, m_orb, n_steps, dt=20.670687, mix=0.1, seed=7, v_scale=2.0e-4):
    """BASELINE config 5's pattern on a cluster: velocity-Verlet steps, every step rebuilds the
    neighbour list / sparse descriptor for the new coordinates (handleCoordinateChange) and runs an
    SCC cycle restarted from the previous step's charges.  make_arm(coords) -> (arm, geo) builds the
    arm's geometry its own way (oracle: CPU restatement; device: dftbp_b200_update_neighbour_list +
    adopt).  dt in atomic time units (20.67 = 0.5 fs).  Returns per-step records and the final state."""
    species0 = np.asarray(workload["species0"])
    mass = np.array([sk.mass[s - 1] for s in species0])[:, None]
    rng = np.random.default_rng(seed)
    x = np.array(workload["coord0"], dtype=float, copy=True)
    v = rng.normal(scale=v_scale, size=x.shape)
    v -= (mass * v).sum(axis=0) / mass.sum()

    def evaluate(x, q_init):
        arm, geo = make_arm(x)
        w = dict(workload, coord0=x)
        return scc_single_point(arm, w, sk, geo, m_orb, mix=mix, max_iter=600, q_init=q_init), geo

    r, geo = evaluate(x, None)
    steps = [dict(n_iter=r["n_iter"], energy=r["energy"], sparseSize=geo["sparseSize"])]
    for _ in range(n_steps):
        a = -r["forces"] / mass
        v = v + 0.5 * dt * a
        x = x + dt * v
        r, geo = evaluate(x, r["q"])
        v = v + 0.5 * dt * (-r["forces"] / mass)
        steps.append(dict(n_iter=r["n_iter"], energy=r["energy"], sparseSize=geo["sparseSize"],
                          e_kin=0.5 * float((mass * v * v).sum())))
    return dict(steps=steps, x=x, v=v, q=r["q"], forces=r["forces"])


# ---- the reference's own SCC potential: gamma = 1/R (cluster or Ewald) - short-range gamma ------
class OracleGamma:
    """TCoulomb + TShortGamma restated (oracle/coulomb_model.py, oracle/dftb_oracle.c): 1/R matrix once per
    geometry, short-range gamma values on the full neighbour list, shifts per SCC iteration."""

    def __init__(self, oracle, sk, workload, hub):
        from oracle import coulomb_model as cm
        self.cm, self.hub = cm, hub
        coord0, species0, lat = np.asarray(workload["coord0"], dtype=float), workload["species0"], workload["latvec"]
        cuts = cm.short_gamma_cutoffs(hub)
        cutoff = max(sk.cutoff(), float(cuts.max()))
        rcell = np.zeros((1, 3)) if lat is None else oracle.cell_translations(lat, cutoff)[1]
        self.nl = oracle.neighbour_list_full(sk.n_orb_species, coord0, species0, rcell, cutoff, sk.cutoff())
        self.vals = cm.short_gamma_values(self.nl, hub, cuts)
        self.invr = cm.invr_cluster(coord0) if lat is None else cm.invr_periodic(coord0, cm.ewald_setup(lat))

    def shifts(self, dq_shell):
        return (self.cm.coulomb_shift(self.invr, dq_shell.sum(axis=1)),
                self.cm.short_gamma_shifts(self.nl, self.hub, self.vals, dq_shell))


class DeviceGamma:
    """The same through the C ABI: dftbp_b200_update_neighbour_list (list cutoff = the largest of all cutoffs,
    sparse structure from the SK cutoff, adopted as the geometry of the pair kernels), _invr_cluster /
    _invr_periodic, _short_gamma_update once per geometry; _coulomb_shift + _short_gamma_shifts per iteration."""

    def __init__(self, hp, sk, workload, hub):
        from oracle import coulomb_model as cm          # input set-up only (Ewald parameters, cutoffs): host logic
        self.hp, self.hub = hp, hub
        coord0, species0, lat = np.asarray(workload["coord0"], dtype=float), workload["species0"], workload["latvec"]
        cuts = cm.short_gamma_cutoffs(hub)
        cutoff = max(sk.cutoff(), float(cuts.max()))
        self.geo = hp.updateNeighbourListAndSpecies(coord0, species0, cutoff, lat, sk_cutoff=sk.cutoff(), adopt=True)
        hp.shortGammaUpdate(hub["nHubbU"], hub["uniqHubbU"], cuts)
        if lat is None:
            self.invr = hp.invRCluster(coord0)
        else:
            ew = cm.ewald_setup(lat)
            self.invr = hp.invRPeriodic(coord0, ew["rLat"], ew["gLat"], ew["alpha"], ew["volume"], ew["maxR"])

    def shifts(self, dq_shell):
        return (self.hp.coulombShift(self.invr, np.ascontiguousarray(dq_shell.sum(axis=1))),
                self.hp.shortGammaShifts(self.hub["iHubbU"], np.ascontiguousarray(dq_shell)))


def scc_reference_potential(arm, gam, workload, sk, geo, m_orb, tol=1e-8, max_iter=100):
    """SCC single point with the reference's potential and defaults: shifts from `gam` (1/R - short-range
    gamma, one Hubbard U per species as without OrbitalResolvedSCC), added as diagonal block shifts
    (addShift_block), atom-reduced charges through the Broyden mixer, Efilling at minTemp.  Both arms run this
    same host loop; the arm and gam objects decide whether a step is the CPU restatement or the CUDA library.
    The gamma-derivative force terms (addGradientsDc, addInvRPrime) are outside the hot path and left out on
    both sides: `forces` is the derivative_shift part."""
    species0 = np.asarray(workload["species0"])
    n_atom, m_shell = geo["nAtom"], max(len(a) for a in sk.ang_shells)
    q0 = np.zeros((n_atom, m_orb))
    shell_of = np.zeros((n_atom, m_orb), dtype=int)
    n_orb_atom = np.zeros(n_atom, dtype=int)
    for i, s in enumerate(species0):
        k = 0
        for ish, (l, occ) in enumerate(zip(sk.ang_shells[s - 1], sk.occ[s - 1])):
            q0[i, k:k + 2 * l + 1] = occ / (2 * l + 1)
            shell_of[i, k:k + 2 * l + 1] = ish
            k += 2 * l + 1
        n_orb_atom[i] = k
    n_elec = q0.sum()

    def shell_charges(dq_orb):
        out = np.zeros((n_atom, m_shell))
        for i in range(n_atom):
            np.add.at(out[i], shell_of[i, :n_orb_atom[i]], dq_orb[i, :n_orb_atom[i]])
        return out

    def potential(q):
        dq_shell = shell_charges(q - q0)
        sh_atom, sh_shell = gam.shifts(dq_shell)
        v = np.zeros((n_atom, m_orb))
        for i in range(n_atom):
            v[i, :n_orb_atom[i]] = sh_atom[i] + sh_shell[i, shell_of[i, :n_orb_atom[i]]]
        e = 0.5 * float(sh_atom @ dq_shell.sum(axis=1)) + 0.5 * float((sh_shell * dq_shell).sum())
        return v, e

    def block(v):
        b = np.zeros((1, n_atom, m_orb, m_orb))
        idx = np.arange(m_orb)
        b[0, :, idx, idx] = v.T
        return b

    mixer = scc_model.BroydenMixer(max_scc_iter=max(max_iter, 2))
    mixer.reset(n_atom)
    h0, over = arm.build()
    q_in = q0.copy()
    q_in_red = q_in.sum(axis=1)
    for it in range(1, max_iter + 1):
        v, _ = potential(q_in)
        ham = arm.add_shift_block(h0, over, block(v))
        eig, C = arm.solve(ham, over)
        f = scc_model.fillings_and_band_energies(np.asarray(eig), n_elec, scc_model.MIN_TEMP)["filling"][:, 0, 0]
        rho = arm.rho(C, f)
        q_out = arm.mulliken(over, rho)
        q_diff_red = q_out.sum(axis=1) - q_in_red
        err = np.abs(q_diff_red).max()
        if err < tol:
            break
        mixer.mix(q_in_red, q_diff_red)
        q_in = np.zeros((n_atom, m_orb))
        for i in range(n_atom):      # expandCharges, atom-wise equivalence
            q_in[i, :n_orb_atom[i]] = q_in_red[i] / n_orb_atom[i]
    v, e_scc = potential(q_out)
    e_h0 = arm.mulliken(h0, rho).sum()
    erho = arm.erho(C, f, np.asarray(eig))
    forces = arm.forces(rho, erho, block(v))
    return dict(n_iter=it, energy=e_h0 + e_scc, e_h0=e_h0, e_scc=e_scc, q=q_out, forces=forces, err=err)

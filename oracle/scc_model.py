"""oracle/scc_model.py -- CPU restatement of the two SCC-loop pieces the hot path is embedded in:
the modified Broyden charge mixer and the Fermi-level search / filling.  TEST INFRASTRUCTURE ONLY
(tests/scc_harness.py): nothing under dftbplus_b200/ imports it.

* BroydenMixer restates TBroydenMixerReal (mixer/broydenmixer.F90:110-345; D. D. Johnson, PRB 38,
  12807 (1988)) statement by statement, with the parser's defaults (dftbplus/parser.F90:5786-5790:
  MixingParameter 0.2, InverseJacobiWeight 0.01, MinimalWeight 1, MaximalWeight 1e5, WeightFactor
  1e-2).  PINNED to the reference: tests/test_scc_model.py replays the known-answer test the
  reference holds for it (test/src/dftbp/unit/mixer/broyden.F90, charges of a water molecule,
  atol = 100 epsilon).
* efilling restates Efilling / electronCount / electronFill / middleGap of dftb/etemp.F90:59-191,
  195-290, 299-466, 493-525 for the Fermi distribution (the reference default; Gaussian /
  Methfessel-Paxton smearing is not restated), including the Neumaier-compensated sums
  (math/summation.F90:20-42), the mid-gap shortcut for integer electron numbers, the bracket
  widening, the bisection termination rule and the final rescaling.  The reference holds no unit
  test for it: checked through its defining properties (electron count, mid gap, aufbau limit).
  fillings_and_band_energies restates the caller's spin convention
  (getFillingsAndBandEnergies, dftbplus/main.F90:4597-4648: half the electrons per level for a
  spin-unpolarised calculation, then everything times two).
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps
ELEC_TOL = 1.0e-15       # common/accuracy.F90:55
ELEC_TOL_MAX = 1.0e-7    # common/accuracy.F90:60
MIN_TEMP = 1.0e-8        # common/accuracy.F90:64 (ElectronicTemperature defaults to 0 -> minTemp)
M_EXP_ARG = 36.043653389117154  # common/accuracy.F90:119


class BroydenMixer:
    """TBroydenMixerReal: init (broydenmixer.F90:110-139), reset (:143-170), mix (:180-201) and
    modifiedBroydenMixing (:205-342)."""

    def __init__(self, max_scc_iter=100, mix_param=0.2, omega0=0.01, min_weight=1.0, max_weight=1.0e5,
                 weight_fac=1.0e-2):
        assert max_scc_iter > 0 and mix_param > 0.0 and omega0 > 0.0
        self.m_iter = max_scc_iter
        self.alpha, self.omega0 = mix_param, omega0
        self.min_weight, self.max_weight, self.weight_fac = min_weight, max_weight, weight_fac
        self.n_elem = 0
        self.reset(0)

    def reset(self, n_elem):
        self.n_elem = n_elem
        self.i_iter = 0
        self.ww = np.zeros(self.m_iter - 1)
        self.aa = np.zeros((self.m_iter - 1, self.m_iter - 1))
        self.q_inp_last = np.zeros(n_elem)
        self.q_diff_last = np.zeros(n_elem)
        self.dF = np.zeros((n_elem, self.m_iter - 1))
        self.uu = np.zeros((n_elem, self.m_iter - 1))

    def mix(self, q_inp_result, q_diff):
        """q_inp_result: input charges on entry, mixed charges on exit (in place); q_diff = q_out - q_in."""
        self.i_iter += 1
        if self.i_iter > self.m_iter:
            raise RuntimeError("Broyden mixer: Maximal nr. of steps exceeded")
        nn = self.i_iter
        q, dq = q_inp_result, np.asarray(q_diff, dtype=np.float64)
        alpha = self.alpha
        if nn == 1:   # first iteration: simple mix and storage of qInp and qDiff
            self.q_inp_last[:] = q
            self.q_diff_last[:] = dq
            q += alpha * dq
            return q
        n1 = nn - 1        # Fortran index nn_1; 0-based position n1 - 1
        ww, aa = self.ww, self.aa
        # weight factor omega of the current iteration
        w = math.sqrt(float(np.dot(dq, dq)))
        w = self.weight_fac / w if w > self.weight_fac / self.max_weight else self.max_weight
        if w < self.min_weight:
            w = self.min_weight
        ww[n1 - 1] = w
        # |DF(m-1)>
        dF_uu = dq - self.q_diff_last
        inv_norm = math.sqrt(float(np.dot(dF_uu, dF_uu)))
        inv_norm = max(inv_norm, EPS)
        inv_norm = 1.0 / inv_norm
        dF_uu = inv_norm * dF_uu
        # a, beta, c
        cc = np.zeros(n1)
        for ii in range(nn - 2):
            aa[ii, n1 - 1] = float(np.dot(self.dF[:, ii], dF_uu))
            aa[n1 - 1, ii] = aa[ii, n1 - 1]
            cc[ii] = ww[ii] * float(np.dot(self.dF[:, ii], dq))
        aa[n1 - 1, n1 - 1] = 1.0
        cc[n1 - 1] = ww[n1 - 1] * float(np.dot(dF_uu, dq))
        beta = np.empty((n1, n1))
        for ii in range(n1):
            beta[:, ii] = ww[:n1] * ww[ii] * aa[:n1, ii]
            beta[ii, ii] += self.omega0 ** 2
        # getrf + getrs(trans='t'): solve beta^T gamma = c (beta is symmetric)
        cc = np.linalg.solve(beta.T, cc)
        self.dF[:, n1 - 1] = dF_uu
        # |u(m-1)>
        dF_uu = alpha * dF_uu + inv_norm * (q - self.q_inp_last)
        self.q_inp_last[:] = q
        self.q_diff_last[:] = dq
        q += alpha * dq
        for ii in range(nn - 2):
            q -= ww[ii] * cc[ii] * self.uu[:, ii]
        q -= ww[n1 - 1] * cc[n1 - 1] * dF_uu
        self.uu[:, n1 - 1] = dF_uu
        return q


class SimpleMixer:
    """TSimpleMixerReal (mixer/simplemixer.F90): q_in + mixParam * (q_out - q_in)."""

    def __init__(self, mix_param=0.05):
        self.alpha = mix_param

    def reset(self, n_elem):
        pass

    def mix(self, q_inp_result, q_diff):
        q_inp_result += self.alpha * np.asarray(q_diff)
        return q_inp_result


def _kahan(total, value, comp):
    """math/summation.F90:20-42 (Neumaier variant)."""
    tmp = total + value
    if abs(total) >= abs(value):
        comp = comp + (total - tmp) + value
    else:
        comp = comp + (value - tmp) + total
    return tmp, comp


def electron_count(ef, eigenvals, kT, k_weight):
    """electronCount, etemp.F90:195-290, Fermi branch (:268-285 without EXP_TRAP).
    eigenvals[level, kpoint, spin]."""
    beta = 1.0 / kT
    total, comp = 0.0, 0.0
    n_lev, n_k, n_spin = eigenvals.shape
    for i_spin in range(n_spin):
        for ii in range(n_k):
            for jj in range(n_lev):
                x = (eigenvals[jj, ii, i_spin] - ef) * beta
                occ = 0.0 if x > 700.0 else k_weight[ii] / (1.0 + math.exp(x))   # exp overflow -> inf -> 0
                total, comp = _kahan(total, occ, comp)
    return total + comp


def electron_fill(ef, eigenvals, kT, k_weight):
    """electronFill, etemp.F90:299-466, Fermi branch (:420-462).  Returns Eband, filling, TS, E0."""
    beta = 1.0 / kT
    n_lev, n_k, n_spin = eigenvals.shape
    filling = np.zeros_like(eigenvals)
    eband, ts = np.zeros(n_spin), np.zeros(n_spin)
    for i_spin in range(n_spin):
        comp_bnd = comp_ts = 0.0
        for ii in range(n_k):
            for jj in range(n_lev):
                x = (eigenvals[jj, ii, i_spin] - ef) * beta
                f = 0.0 if x > 700.0 else 1.0 / (1.0 + math.exp(x))
                filling[jj, ii, i_spin] = f
                if f <= ELEC_TOL:
                    break      # exit lpLevels: the remaining (higher) levels keep filling 0
                if EPS < f < 1.0 - EPS:
                    ts[i_spin], comp_ts = _kahan(ts[i_spin], -k_weight[ii] * (f * math.log(f) + (1.0 - f) * math.log(1.0 - f)),
                                                 comp_ts)
                eband[i_spin], comp_bnd = _kahan(eband[i_spin], k_weight[ii] * (f * eigenvals[jj, ii, i_spin]), comp_bnd)
        eband[i_spin] += comp_bnd
        ts[i_spin] += comp_ts
    ts = ts * kT
    return eband, filling, ts, eband - 0.5 * ts


def middle_gap(eigenvals, k_weight, n_electrons):
    """middleGap, etemp.F90:493-525: aufbau filling of the levels sorted over (level, k, spin)."""
    n_lev, n_k, n_spin = eigenvals.shape
    flat = eigenvals.reshape(-1, order="F")          # reshape(eigenvals, [size]) in Fortran order
    idx = np.argsort(flat, kind="stable")            # (index_heap_sort: the order inside ties does not matter here)
    ind, n_elec = 0, 0.0
    i_lev = idx[0]
    while n_elec < n_electrons:
        i_lev = idx[ind]
        i_k = (i_lev // n_lev) % n_k
        n_elec += k_weight[i_k]
        ind += 1
    ind = min(flat.size - 1, ind)
    j_lev = idx[ind]
    return 0.5 * (flat[j_lev] + flat[i_lev])


def efilling(eigenvals, n_electrons, kT, k_weight):
    """Efilling, etemp.F90:59-191 (Fermi distribution).  eigenvals[level, kpoint, spin], one electron
    per level.  Returns dict(Ebs, Ef, TS, E0, filling)."""
    eigenvals = np.asarray(eigenvals, dtype=np.float64)
    k_weight = np.asarray(k_weight, dtype=np.float64)
    n_lev, n_k, n_spin = eigenvals.shape
    assert kT > 0.0 and n_electrons >= 0.0
    epsilon2 = 2.0 * EPS
    eps_sp = float(np.finfo(np.float32).eps)

    def done(ef, scale=None):
        eb, fill, ts, e0 = electron_fill(ef, eigenvals, kT, k_weight)
        if scale is not None:
            fill = fill * scale[0] / scale[1]        # filling(:,:,:) = filling * nElectrons / nElec (left to right)
        return dict(Ebs=eb, Ef=ef, TS=ts, E0=e0, filling=fill)

    if n_electrons < EPS:
        z = np.zeros(n_spin)
        return dict(Ebs=z, Ef=eigenvals.min() - 1000.0 * (kT + eps_sp), TS=z.copy(), E0=z.copy(),
                    filling=np.zeros_like(eigenvals))
    if n_lev * n_spin <= n_electrons:
        return done(eigenvals.max() + 1000.0 * (kT + eps_sp))
    # integer number of electrons: try the middle of the gap
    if abs(n_electrons - round(n_electrons)) <= ELEC_TOL:
        ef = middle_gap(eigenvals, k_weight, n_electrons)
        n_elec = electron_count(ef, eigenvals, kT, k_weight)
        if abs(n_electrons - n_elec) <= ELEC_TOL_MAX:
            return done(ef)
    upper = eigenvals[-1, :, :].max() + 0.01
    lower = eigenvals[0, :, :].min() - 0.01
    n_max = electron_count(upper, eigenvals, kT, k_weight)
    n_min = electron_count(lower, eigenvals, kT, k_weight)
    while n_min > n_electrons:
        lower = 2.0 * (lower - upper) + lower
        n_min = electron_count(lower, eigenvals, kT, k_weight)
    while n_max < n_electrons:
        upper = 2.0 * (upper - lower) + lower
        n_max = electron_count(upper, eigenvals, kT, k_weight)
    ef = 0.5 * (upper + lower)
    n_elec = electron_count(ef, eigenvals, kT, k_weight)
    while abs(n_electrons - n_elec) > ELEC_TOL and abs(upper - lower) >= max(abs(ef) * epsilon2, epsilon2):
        if (n_max >= n_min) == (n_electrons >= n_elec):
            lower, n_min = ef, n_elec
        else:
            upper, n_max = ef, n_elec
        ef = 0.5 * (upper + lower)
        n_elec = electron_count(ef, eigenvals, kT, k_weight)
    if abs(n_electrons - n_elec) > ELEC_TOL_MAX:
        raise RuntimeError("Fermi level search did not converge.")
    scale = None
    if n_electrons > epsilon2 and abs(n_electrons - n_elec) > epsilon2:
        scale = (n_electrons, n_elec)    # "Re-scale to give exact number of electrons"
    return done(ef, scale)


def fillings_and_band_energies(eigenvals, n_electrons, kT=MIN_TEMP, k_weight=(1.0,)):
    """getFillingsAndBandEnergies, dftbplus/main.F90:4597-4648, for one spin block (spin-unpolarised:
    nElecFill = nElectrons / 2, results times two) or two collinear channels with a common Fermi
    level.  eigenvals[level] / [level, kpoint] / [level, kpoint, spin]."""
    e = np.asarray(eigenvals, dtype=np.float64)
    while e.ndim < 3:
        e = e[..., None]
    unpolarised = e.shape[2] == 1
    r = efilling(e, 0.5 * n_electrons if unpolarised else n_electrons, kT, k_weight)
    if unpolarised:
        for key in ("Ebs", "E0", "TS", "filling"):
            r[key] = 2.0 * r[key]
    return r

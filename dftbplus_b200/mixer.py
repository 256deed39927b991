"""Host-side charge mixer of the SCC loop: the modified Broyden scheme DFTB+ uses by default
(TBroydenMixerReal, mixer/broydenmixer.F90:110-345; D. D. Johnson, PRB 38, 12807 (1988); defaults of
dftbplus/parser.F90:5786-5790).  The mixer belongs to the Fortran host in a DFTB+ integration (SURVEY
8(f) N4); this mirror exists for drivers written against `dftbplus_b200.api` -- bench.py's MD section
(BASELINE config 5) runs its SCC cycles with it -- so that they iterate the way the reference does.
The vectors are O(nAtom): NumPy on the host.  Checked against the reference's own known-answer test
(tests/test_scc_model.py)."""
import math

import numpy as np


class BroydenMixer:
    def __init__(self, max_scc_iter=100, mix_param=0.2, omega0=0.01, min_weight=1.0, max_weight=1.0e5,
                 weight_fac=1.0e-2):
        if not (max_scc_iter > 0 and mix_param > 0.0 and omega0 > 0.0):
            raise ValueError("BroydenMixer: maxSccIter, mixParam and omega0 must be positive")
        self.m_iter, self.alpha, self.omega0 = int(max_scc_iter), float(mix_param), float(omega0)
        self.min_weight, self.max_weight, self.weight_fac = float(min_weight), float(max_weight), float(weight_fac)
        self.reset(0)

    def reset(self, n_elem):
        """Makes the mixer ready for a new SCC cycle (the reference resets it for every geometry)."""
        m = self.m_iter - 1
        self.n_elem, self.i_iter = int(n_elem), 0
        self._w = np.zeros(m)
        self._a = np.zeros((m, m))
        self._q_last = np.zeros(self.n_elem)
        self._d_last = np.zeros(self.n_elem)
        self._df = np.zeros((m, self.n_elem))
        self._u = np.zeros((m, self.n_elem))

    def mix(self, q_inp, q_diff):
        """q_inp: input charges -> mixed charges (returned; q_inp is not modified); q_diff = q_out - q_inp."""
        q = np.array(q_inp, dtype=np.float64, copy=True)
        d = np.asarray(q_diff, dtype=np.float64)
        if q.shape != (self.n_elem,) or d.shape != (self.n_elem,):
            raise ValueError("BroydenMixer.mix: vectors of length %d expected" % self.n_elem)
        self.i_iter += 1
        if self.i_iter > self.m_iter:
            raise RuntimeError("Broyden mixer: Maximal nr. of steps exceeded")
        n = self.i_iter
        if n == 1:
            self._q_last[:], self._d_last[:] = q, d
            return q + self.alpha * d
        m = n - 2                                   # 0-based position of iteration n - 1
        norm_d = math.sqrt(float(d @ d))
        w = self.weight_fac / norm_d if norm_d > self.weight_fac / self.max_weight else self.max_weight
        self._w[m] = max(w, self.min_weight)
        df = d - self._d_last
        inv = 1.0 / max(math.sqrt(float(df @ df)), np.finfo(np.float64).eps)
        df = inv * df
        c = np.zeros(m + 1)
        for i in range(m):
            self._a[i, m] = self._a[m, i] = float(self._df[i] @ df)
            c[i] = self._w[i] * float(self._df[i] @ d)
        self._a[m, m] = 1.0
        c[m] = self._w[m] * float(df @ d)
        ww = self._w[:m + 1]
        beta = ww[:, None] * ww[None, :] * self._a[:m + 1, :m + 1]
        beta[np.diag_indices(m + 1)] += self.omega0 ** 2
        gamma = np.linalg.solve(beta.T, c)
        self._df[m] = df
        u = self.alpha * df + inv * (q - self._q_last)
        self._q_last[:], self._d_last[:] = q, d
        out = q + self.alpha * d
        for i in range(m):
            out -= self._w[i] * gamma[i] * self._u[i]
        out -= self._w[m] * gamma[m] * u
        self._u[m] = u
        return out

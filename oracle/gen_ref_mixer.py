#!/usr/bin/env python3
"""Golden vectors of the modified Broyden mixer made by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/fortran_exec.py):
modifiedBroydenMixing (mixer/broydenmixer.F90:204-342; the Fypp template instantiated for real vectors), getrf / getrs = the
LAPACK routines (SciPy).  TEST INFRASTRUCTURE ONLY.  Outputs -> tests/golden/mixer_ref_golden.npz;
tests/test_oracle_mixer_ref.py compares oracle/scc_model.py and the product's host mirror dftbplus_b200/mixer.py with them."""
import os
import sys

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import fortran_exec as fx  # noqa: E402

REF = "/root/reference/src/dftbp/mixer/broydenmixer.F90"


def routine():
    src = open(REF).read().replace("${TYPE}$", "real").replace("${LABEL}$", "Real")
    state = {}

    def getrf(beta, ipiv):
        state["lu"] = scipy.linalg.lu_factor(np.array(beta.a, order="F"))

    def getrs(beta, ipiv, cc, trans="n"):
        cc.a[:] = scipy.linalg.lu_solve(state["lu"], cc.a, trans=1 if trans == "t" else 0)

    calls = {"getrf": lambda a: "getrf(%s)" % ", ".join(a), "getrs": lambda a: "getrs(%s)" % ", ".join(a)}
    return fx.compile_routine(src, "subroutine", "modifiedBroydenMixingReal", calls, extra_env=dict(getrf=getrf, getrs=getrs))[0]


class RefMixer:
    """TBroydenMixerReal_init / _reset / _mix (broydenmixer.F90:114-200) around the executed routine."""

    def __init__(self, f, n, m_iter=100, alpha=0.2, omega0=0.01, min_w=1.0, max_w=1.0e5, w_fac=1.0e-2):
        self.f, self.n, self.p = f, n, (alpha, omega0, min_w, max_w, w_fac)
        m = m_iter - 1
        self.it = 0
        self.ww = fx.FArr(np.zeros(m))
        self.aa = fx.FArr(np.zeros((m, m), order="F"))
        self.q_last, self.d_last = fx.FArr(np.zeros(n)), fx.FArr(np.zeros(n))
        self.dF, self.uu = fx.FArr(np.zeros((n, m), order="F")), fx.FArr(np.zeros((n, m), order="F"))

    def mix(self, q, diff):
        self.it += 1
        a, w0, mn, mx, wf = self.p
        self.f(fx.FArr(q), self.q_last, self.d_last, self.aa, self.ww, self.it, fx.FArr(np.asarray(diff, dtype=np.float64)), a, w0, mn, mx,
               wf, self.n, self.dF, self.uu)
        return q


def problems():
    out = {}
    for n in (3, 11, 40):
        rng = np.random.default_rng(n)
        d = rng.uniform(-0.5, 0.5, n)
        low = rng.normal(size=(n, 2)) / np.sqrt(n)
        b = rng.normal(size=n)
        out["lin%d" % n] = (n, d, low, b)
    return out


def iterate(mixer, n, d, low, b, steps=14):
    q = np.zeros(n)
    hist = []
    for _ in range(steps):
        diff = d * q + 0.3 * low @ (low.T @ q) + b - q
        mixer.mix(q, diff)
        hist.append(q.copy())
    return np.array(hist)


def main():
    if not os.path.exists(REF):
        print("reference not present; nothing to do", file=sys.stderr)
        return
    f = routine()
    g = {}
    for name, (n, d, low, b) in problems().items():
        g[name] = iterate(RefMixer(f, n), n, d, low, b)
        print(name, "last iterate", g[name][-1][:3])
    path = os.path.join(ROOT, "tests", "golden", "mixer_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path)


if __name__ == "__main__":
    main()

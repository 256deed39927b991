#!/usr/bin/env python3
"""Generate dftbplus_b200/csrc/sk_gaunt_code.cuh (and, with --tables, sk_gaunt_tables.h) -- the constants behind the CUDA
Slater-Koster rotation.

Design (B200-first, not a transcription of sk.F90's Maple polynomials): a two-centre block is
a rotationally covariant function of the bond direction d, so it has a finite expansion in
tesseral harmonics of d,

    E^{l1 l2}(m2, m1; d) = sum_L  A_L  sum_M  G^{l1 l2}(m1, m2; L M)  S_LM(d),
    L = l1+l2, l1+l2-2, ..., |l1-l2|,

with G the real Gaunt coefficients of Racah-normalised tesseral harmonics and A_L a fixed
linear map of the min(l1,l2)+1 diatomic integrals V_sigma, V_pi, ... (fixed by the d = z case,
where E is diagonal: E(m,m; z) = V_|m|).  That makes the device code uniform: every lane
evaluates one output element as a short list of (coef, S_LM index, A slot) terms -- no
per-shell-pair branches, coalesced stores, s/p/d/f all handled by the same loop.

Everything is computed numerically in float64 with an exact product quadrature (Gauss-
Legendre x uniform phi, exact beyond the degree 12 needed) and self-checked against the
definition E = sum_sigma D^{l1} D^{l2} V before being written.

Orbital order within a shell: m = -l..l, tesseral (reference: src/dftbp/type/orbitals.F90:59-64;
sign/normalisation read off the sigma columns of src/dftbp/dftb/sk.F90:160-240).
"""
import math
import os

import numpy as np

LMAX = 3            # orbitals up to f (mAngRot_ = 3, sk.F90:28)
LY = 2 * LMAX       # harmonics of the direction up to L = 6
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "dftbplus_b200", "csrc", "sk_gaunt_tables.h")


def legendre_deriv_coeffs(L, m):
    """coefficients (ascending powers of z) of sqrt((2-d_m0)(L-m)!/(L+m)!) d^m P_L / dz^m."""
    p = np.polynomial.legendre.Legendre.basis(L).convert(kind=np.polynomial.Polynomial)
    for _ in range(m):
        p = p.deriv()
    norm = math.sqrt((2.0 if m > 0 else 1.0) * math.factorial(L - m) / math.factorial(L + m))
    c = np.zeros(L - m + 1)
    c[:len(p.coef)] = p.coef
    return norm * c


def tesseral(L, x, y, z):
    """Racah-normalised tesseral harmonics S_LM(d), M = -L..L, on unit vectors (arrays ok)."""
    out = []
    xy = x + 1j * y
    for M in range(-L, L + 1):
        m = abs(M)
        c = legendre_deriv_coeffs(L, m)
        pz = sum(ck * z ** k for k, ck in enumerate(c))
        ab = xy ** m
        out.append(pz * (ab.imag if M < 0 else ab.real))
    return np.array(out)


def quadrature(n_theta=10, n_phi=24):
    t, w = np.polynomial.legendre.leggauss(n_theta)
    phi = (np.arange(n_phi) + 0.5) * 2.0 * np.pi / n_phi
    T, P = np.meshgrid(t, phi, indexing="ij")
    W = np.repeat(w[:, None], n_phi, axis=1) / (2.0 * n_phi)   # sums to 1
    s = np.sqrt(1.0 - T ** 2)
    return (s * np.cos(P)).ravel(), (s * np.sin(P)).ravel(), T.ravel(), W.ravel()


def build():
    x, y, z, w = quadrature()
    S = [tesseral(L, x, y, z) for L in range(LY + 1)]
    # orthonormality self-check: <S_LM S_LM'> = delta / (2L+1)
    for L in range(LY + 1):
        g = (S[L] * w) @ S[L].T
        assert np.allclose(g, np.eye(2 * L + 1) / (2 * L + 1), atol=1e-14), L
    tables = {}
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            lmin = min(l1, l2)
            Ls = [l1 + l2 - 2 * k for k in range(lmin + 1)]
            # G[k][a, b, M]: a indexes m1 (shell on atom 1), b indexes m2
            G = []
            for L in Ls:
                g = np.einsum("aq,bq,Mq,q->abM", S[l1], S[l2], S[L], w) * (2 * L + 1)
                g[np.abs(g) < 1e-13] = 0.0
                G.append(g)
            # d = z: S_LM(z) = delta_M0  =>  V_mu = sum_k G[k][mu,mu,M=0] A_k
            gm = np.array([[G[k][l1 + mu, l2 + mu, Ls[k]] for k in range(lmin + 1)]
                           for mu in range(lmin + 1)])
            T = np.linalg.inv(gm)                      # A_k = sum_mu T[k, mu] V_mu
            # reference convention for l1 > l2 (sk.F90:128-133): (-1)^(l1+l2) * transpose of the
            # (l2, l1) routine fed with the column taken from the B-A file
            sgn = -1.0 if (l1 > l2 and (l1 + l2) % 2) else 1.0
            tables[(l1, l2)] = (Ls, [sgn * g for g in G], T)
    return tables


def self_check(tables):
    """E from the tables == sum_sigma D^{l2} D^{l1} V with D from a projection of rotated
    harmonics (independent of the Gaunt route)."""
    rng = np.random.default_rng(5)
    x, y, z, w = quadrature()
    pts = np.stack([x, y, z])
    worst = 0.0
    for trial in range(6):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if trial == 0:
            d = np.array([0.0, 0.0, -1.0])
        # frame with third column d
        a = np.array([1.0, 0, 0]) if abs(d[0]) < 0.9 else np.array([0, 1.0, 0])
        e1 = np.cross(a, d); e1 /= np.linalg.norm(e1)
        e2 = np.cross(d, e1)
        R = np.stack([e1, e2, d], axis=1)
        loc = R.T @ pts
        D = []
        for l in range(LMAX + 1):
            D.append(((tesseral(l, *pts) * w) @ tesseral(l, *loc).T) * (2 * l + 1))
        Y = [tesseral(L, d[0], d[1], d[2]) for L in range(LY + 1)]
        for (l1, l2), (Ls, G, T) in tables.items():
            lmin = min(l1, l2)
            V = rng.normal(size=lmin + 1)
            A = T @ V
            E = sum(A[k] * (G[k] @ Y[Ls[k]]) for k in range(lmin + 1))   # [a, b]
            sgn = -1.0 if (l1 > l2 and (l1 + l2) % 2) else 1.0
            Eref = np.zeros_like(E)
            for s in range(-lmin, lmin + 1):
                Eref += sgn * np.outer(D[l1][:, l1 + s], D[l2][:, l2 + s]) * V[abs(s)]
            worst = max(worst, np.abs(E - Eref).max())
    assert worst < 5e-14, worst
    return worst


def emit(tables, write=False):
    lines = []
    ap = lines.append
    ap("// GENERATED by tools/gen_sk_gaunt.py -- do not edit.  See that script for the derivation.")
    ap("#pragma once")
    ap("namespace dftbp_b200 { namespace skg {")
    ap("constexpr int kLMax = %d;            // highest orbital angular momentum (f)" % LMAX)
    ap("constexpr int kLY = %d;              // highest harmonic of the bond direction" % LY)
    ap("constexpr int kNY = %d;              // number of S_LM values, index L*L + L + M" % ((LY + 1) ** 2))
    # --- harmonics of the direction in product form
    ap("// S_LM(d) = zpoly(z) * (M<0 ? Im : Re)(x+iy)^|M|,  zpoly(z) = z^par * (c0 + c1 z^2 + c2 z^4 + c3 z^6)")
    ap("struct YDesc { double c[4]; int m; int par; int isSin; int pad; };")
    ap("static const YDesc kYDesc[kNY] = {")
    for L in range(LY + 1):
        for M in range(-L, L + 1):
            m = abs(M)
            c = legendre_deriv_coeffs(L, m)
            par = (L - m) & 1
            ev = list(c[par::2]) + [0.0] * 4
            ap("  {{%.17e, %.17e, %.17e, %.17e}, %d, %d, %d, 0}, // L=%d M=%d" %
               (ev[0], ev[1], ev[2], ev[3], m, par, 1 if M < 0 else 0, L, M))
    ap("};")
    # --- shell-pair tables
    ap("// One term of an output element: value += coef * S[y] * A[k]   (k-th L of the shell pair)")
    ap("struct GTerm { double coef; short y; short k; };")
    ap("// A_k = sum_mu T[k][mu] * V_mu for shell pair (l1, l2); element (m1 index a, m2 index b) owns")
    ap("// terms [elemStart[a*(2*l2+1)+b], elemStart[.. + 1]) of its shell-pair term list.")
    terms_all = []
    offs = {}
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            Ls, G, T = tables[(l1, l2)]
            starts = []
            for a in range(2 * l1 + 1):
                for b in range(2 * l2 + 1):
                    starts.append(len(terms_all))
                    for k, L in enumerate(Ls):
                        for M in range(-L, L + 1):
                            g = G[k][a, b, M + L]
                            if g != 0.0:
                                terms_all.append((g, L * L + L + M, k))
            starts.append(len(terms_all))
            offs[(l1, l2)] = starts
    ap("constexpr int kNTerms = %d;" % len(terms_all))
    ap("static const GTerm kTerms[kNTerms] = {")
    for g, yi, k in terms_all:
        ap("  {%.17e, %d, %d}," % (g, yi, k))
    ap("};")
    ap("// kElemStart[l1][l2][a*(2*l2+1)+b], one extra sentinel entry per shell pair")
    ap("static const int kElemStart[kLMax + 1][kLMax + 1][50] = {")
    for l1 in range(LMAX + 1):
        ap(" {")
        for l2 in range(LMAX + 1):
            s = offs[(l1, l2)] + [0] * (50 - len(offs[(l1, l2)]))
            ap("  {" + ", ".join(str(v) for v in s) + "},")
        ap(" },")
    ap("};")
    ap("// kT[l1][l2][k][mu]")
    ap("static const double kT[kLMax + 1][kLMax + 1][kLMax + 1][kLMax + 1] = {")
    for l1 in range(LMAX + 1):
        ap(" {")
        for l2 in range(LMAX + 1):
            T = np.zeros((LMAX + 1, LMAX + 1))
            t = tables[(l1, l2)][2]
            T[:t.shape[0], :t.shape[1]] = t
            ap("  {" + ", ".join("{" + ", ".join("%.17e" % v for v in row) + "}" for row in T) + "},")
        ap(" },")
    ap("};")
    ap("}}  // namespace")
    if write:
        with open(OUT, "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return len(terms_all)


OUT_CODE = os.path.join(HERE, "..", "dftbplus_b200", "csrc", "sk_gaunt_code.cuh")


def emit_code(tables):
    """Straight-line device code for the lane-per-pair kernels (K1/K5 v2): every thread owns one
    neighbour pair, all 32 lanes of a warp run the same shell-pair type, so the Gaunt sums are
    emitted as unrolled FMAs with the coefficients as immediates -- no tables, no branches."""
    L = []
    ap = L.append
    pool = []
    pool_index = {}

    def C(v):
        """reference to a coefficient in the __constant__ pool: DFMA / DMUL take constant-bank
        operands directly, whereas a 64-bit immediate costs two UMOVs per use"""
        key = float(v).hex()
        if key not in pool_index:
            pool_index[key] = len(pool)
            pool.append(float(v))
        return "skgC[%d]" % pool_index[key]

    ap("// GENERATED by tools/gen_sk_gaunt.py (emit_code) -- do not edit.")
    ap("// Straight-line Slater-Koster rotation: harmonics of the bond direction and one function per")
    ap("// ordered shell-pair type (l1 = shell of the atom, l2 = shell of the neighbour).")
    ap("// emit(a, b, h, s): a = m1 index (column inside the shell of atom 1), b = m2 index (row).")
    ap("#pragma once")
    ap("namespace dftbp_b200 { namespace skg {")
    # ---- harmonics
    ap("template <int LY>")
    ap("__device__ __forceinline__ void harmonics(double x, double y, double z, double (&Y)[(LY + 1) * (LY + 1)]) {")
    ap("  const double z2 = z * z;")
    ap("  const double c1 = x, s1 = y;")
    for m in range(2, LY + 1):
        ap("  [[maybe_unused]] const double c%d = c%d * x - s%d * y, s%d = c%d * y + s%d * x;" % (m, m - 1, m - 1, m, m - 1, m - 1))
    ap("  Y[0] = 1.0;")
    for Lq in range(1, LY + 1):
        ap("  if constexpr (LY >= %d) {" % Lq)
        for m in range(0, Lq + 1):
            c = legendre_deriv_coeffs(Lq, m)
            par = (Lq - m) & 1
            ev = list(c[par::2])
            # Horner in z2
            expr = "%.17e" % ev[-1]
            for ck in reversed(ev[:-1]):
                expr = "fma(z2, %s, %.17e)" % (expr, ck)
            if par:
                expr = "(%s) * z" % expr
            if m == 0:
                ap("    Y[%d] = %s;" % (Lq * Lq + Lq, expr))
            else:
                ap("    { const double p = %s; Y[%d] = p * c%d; Y[%d] = p * s%d; }" %
                   (expr, Lq * Lq + Lq + m, m, Lq * Lq + Lq - m, m))
        ap("  }")
    ap("}")
    # ---- shell-pair blocks
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            Ls, G, T = tables[(l1, l2)]
            lmin = min(l1, l2)
            nk = lmin + 1
            ap("// (l1, l2) = (%d, %d): L = %s" % (l1, l2, Ls))
            ap("template <bool DO_H, bool DO_S, int NY, class Emit>")
            ap("__device__ __forceinline__ void block_%d_%d(const double (&Y)[NY], const double (&vH)[4], const double (&vS)[4], Emit &&emit) {" % (l1, l2))
            ap("  static_assert(NY >= %d, \"harmonics up to L = %d needed\");" % ((l1 + l2 + 1) ** 2, l1 + l2))
            for which in ("H", "S"):
                for k in range(nk):
                    terms = []
                    for mu in range(nk):
                        if T[k, mu] != 0.0:
                            terms.append("%s * v%s[%d]" % (C(T[k, mu]), which, mu))
                    expr = terms[0]
                    for t in terms[1:]:
                        c, v = t.split(" * ")
                        expr = "fma(%s, %s, %s)" % (c, v, expr)
                    ap("  [[maybe_unused]] const double a%s%d = DO_%s ? %s : 0.0;" % (which, k, which, expr))
            for a in range(2 * l1 + 1):
                for b in range(2 * l2 + 1):
                    ap("  {")
                    used = []
                    for k, Lq in enumerate(Ls):
                        terms = [(G[k][a, b, M + Lq], Lq * Lq + Lq + M) for M in range(-Lq, Lq + 1)
                                 if G[k][a, b, M + Lq] != 0.0]
                        if not terms:
                            continue
                        expr = "%s * Y[%d]" % (C(terms[0][0]), terms[0][1])
                        for g, yi in terms[1:]:
                            expr = "fma(%s, Y[%d], %s)" % (C(g), yi, expr)
                        ap("    const double t%d = %s;" % (k, expr))
                        used.append(k)
                    if used:
                        for which in ("H", "S"):
                            expr = "a%s%d * t%d" % (which, used[0], used[0])
                            for k in used[1:]:
                                expr = "fma(a%s%d, t%d, %s)" % (which, k, k, expr)
                            ap("    const double %s = DO_%s ? %s : 0.0;" % (which.lower(), which, expr))
                        ap("    emit(%d, %d, h, s);" % (a, b))
                    else:
                        ap("    emit(%d, %d, 0.0, 0.0);" % (a, b))
                    ap("  }")
            ap("}")
    # ---- column-wise forward rotation (K1) and weight-contracted forms (K5), per shell-pair type
    def wcount(l1, l2):
        return (min(l1, l2) + 1) * (2 * max(l1, l2) + 1)

    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            Ls, G, T = tables[(l1, l2)]
            lmin = min(l1, l2)
            nk = lmin + 1
            # A_k = sum_mu T[k, mu] V_mu
            ap("// (l1, l2) = (%d, %d): expansion coefficients, one column, weight contraction" % (l1, l2))
            ap("__device__ __forceinline__ void acoef_%d_%d(const double *v, double *a) {" % (l1, l2))
            for k in range(nk):
                terms = [(T[k, mu], mu) for mu in range(nk) if T[k, mu] != 0.0]
                expr = "%s * v[%d]" % (C(terms[0][0]), terms[0][1])
                for c, mu in terms[1:]:
                    expr = "fma(%s, v[%d], %s)" % (C(c), mu, expr)
                ap("  a[%d] = %s;" % (k, expr))
            ap("}")
            # column A of the block: emit(b, h, s), b = m2 index (row inside the shell of atom 2)
            ap("template <int A, bool DO_H, bool DO_S, int NY, class Emit>")
            ap("__device__ __forceinline__ void col_%d_%d(const double (&Y)[NY], const double *aH, const double *aS, Emit &&emit) {" % (l1, l2))
            ap("  static_assert(NY >= %d, \"harmonics up to L = %d needed\");" % ((l1 + l2 + 1) ** 2, l1 + l2))
            for a in range(2 * l1 + 1):
                ap("  %sif constexpr (A == %d) {" % ("" if a == 0 else "else ", a))
                for b in range(2 * l2 + 1):
                    ap("    {")
                    used = []
                    for k, Lq in enumerate(Ls):
                        terms = [(G[k][a, b, M + Lq], Lq * Lq + Lq + M) for M in range(-Lq, Lq + 1)
                                 if G[k][a, b, M + Lq] != 0.0]
                        if not terms:
                            continue
                        expr = "%s * Y[%d]" % (C(terms[0][0]), terms[0][1])
                        for g, yi in terms[1:]:
                            expr = "fma(%s, Y[%d], %s)" % (C(g), yi, expr)
                        ap("      const double t%d = %s;" % (k, expr))
                        used.append(k)
                    if used:
                        for which in ("H", "S"):
                            expr = "a%s[%d] * t%d" % (which, used[0], used[0])
                            for k in used[1:]:
                                expr = "fma(a%s[%d], t%d, %s)" % (which, k, k, expr)
                            ap("      const double %s = DO_%s ? %s : 0.0;" % (which.lower(), which, expr))
                        ap("      emit(%d, h, s);" % b)
                    else:
                        ap("      emit(%d, 0.0, 0.0);" % b)
                    ap("    }")
                ap("  }")
            ap("}")
            # W_j = sum_(a,b) G[k][a,b,M] w(a,b), j = (k, M) flattened; %d values
            ap("template <class GetW, class PutW>")
            ap("__device__ __forceinline__ void wbuild_%d_%d(GetW &&getw, PutW &&putW) {" % (l1, l2))
            for a in range(2 * l1 + 1):
                for b in range(2 * l2 + 1):
                    ap("  const double w_%d_%d = getw(%d, %d);" % (a, b, a, b))
            j = 0
            for k, Lq in enumerate(Ls):
                for M in range(-Lq, Lq + 1):
                    terms = [(G[k][a, b, M + Lq], a, b) for a in range(2 * l1 + 1) for b in range(2 * l2 + 1)
                             if G[k][a, b, M + Lq] != 0.0]
                    if terms:
                        expr = "%s * w_%d_%d" % (C(terms[0][0]), terms[0][1], terms[0][2])
                        for g, a, b in terms[1:]:
                            expr = "fma(%s, w_%d_%d, %s)" % (C(g), a, b, expr)
                    else:
                        expr = "0.0"
                    ap("  const double W%d = %s;" % (j, expr))
                    j += 1
            assert j == wcount(l1, l2)
            for jj in range(j):  # W_j is parked in the slot of the j-th element (a, b) of the sub-block
                ap("  putW(%d, %d, W%d);" % (jj // (2 * l2 + 1), jj % (2 * l2 + 1), jj))
            ap("}")
            # sum_k a_k sum_M W_(k,M) S_(L_k M)
            ap("template <int NY, class GetW>")
            ap("__device__ __forceinline__ double wdot_%d_%d(const double (&Y)[NY], const double *a, GetW &&getW) {" % (l1, l2))
            ap("  static_assert(NY >= %d, \"harmonics up to L = %d needed\");" % ((l1 + l2 + 1) ** 2, l1 + l2))
            j = 0
            for k, Lq in enumerate(Ls):
                expr = None
                for M in range(-Lq, Lq + 1):
                    yi = Lq * Lq + Lq + M
                    slot = "%d, %d" % (j // (2 * l2 + 1), j % (2 * l2 + 1))
                    expr = ("getW(%s) * Y[%d]" % (slot, yi)) if expr is None else ("fma(getW(%s), Y[%d], %s)" % (slot, yi, expr))
                    j += 1
                ap("  const double u%d = %s;" % (k, expr))
            expr = "a[0] * u0"
            for k in range(1, nk):
                expr = "fma(a[%d], u%d, %s)" % (k, k, expr)
            ap("  return %s;" % expr)
            ap("}")
    # compile-time and warp-uniform run-time selectors of the new forms
    ap("template <int L1, int L2>")
    ap("__device__ __forceinline__ void acoefT(const double *v, double *a) {")
    first = True
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("  %sif constexpr (L1 == %d && L2 == %d) acoef_%d_%d(v, a);" % ("" if first else "else ", l1, l2, l1, l2))
            first = False
    ap("}")
    ap("template <int L1, int L2, int A, bool DO_H, bool DO_S, int NY, class Emit>")
    ap("__device__ __forceinline__ void colT(const double (&Y)[NY], const double *aH, const double *aS, Emit &&emit) {")
    first = True
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("  %sif constexpr (L1 == %d && L2 == %d) col_%d_%d<A, DO_H, DO_S>(Y, aH, aS, emit);" %
               ("" if first else "else ", l1, l2, l1, l2))
            first = False
    ap("}")
    ap("template <int L1, int L2, class GetW, class PutW>")
    ap("__device__ __forceinline__ void wbuildT(GetW &&getw, PutW &&putW) {")
    first = True
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("  %sif constexpr (L1 == %d && L2 == %d) wbuild_%d_%d(getw, putW);" % ("" if first else "else ", l1, l2, l1, l2))
            first = False
    ap("}")
    ap("template <int L1, int L2, int NY, class GetW>")
    ap("__device__ __forceinline__ double wdotT(const double (&Y)[NY], const double *a, GetW &&getW) {")
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("  if constexpr (L1 == %d && L2 == %d) return wdot_%d_%d(Y, a, getW);" % (l1, l2, l1, l2))
    ap("  return 0.0;")
    ap("}")
    ap("template <int LMAX, class GetW, class PutW>")
    ap("__device__ __forceinline__ void wbuildDispatch(int l1, int l2, GetW &&getw, PutW &&putW) {")
    ap("  switch (l1 * 4 + l2) {")
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("    case %d: if constexpr (LMAX >= %d) wbuild_%d_%d(getw, putW); break;" % (l1 * 4 + l2, max(l1, l2), l1, l2))
    ap("    default: break;")
    ap("  }")
    ap("}")
    ap("template <int LMAX, class GetW>")
    ap("__device__ __forceinline__ double wdotDispatch(int l1, int l2, const double (&Y)[(2 * LMAX + 1) * (2 * LMAX + 1)],")
    ap("                                                const double *v, GetW &&getW) {")
    ap("  double a[4];")
    ap("  switch (l1 * 4 + l2) {")
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("    case %d: if constexpr (LMAX >= %d) { acoef_%d_%d(v, a); return wdot_%d_%d(Y, a, getW); } break;" %
               (l1 * 4 + l2, max(l1, l2), l1, l2, l1, l2))
    ap("    default: break;")
    ap("  }")
    ap("  return 0.0;")
    ap("}")
    # ---- dispatch
    ap("// Uniform (per warp) dispatch on the shell-pair type; LMAX = highest l of any species.")
    ap("template <bool DO_H, bool DO_S, int LMAX, class Emit>")
    ap("__device__ __forceinline__ void blockDispatch(int l1, int l2, const double (&Y)[(2 * LMAX + 1) * (2 * LMAX + 1)],")
    ap("                                              const double (&vH)[4], const double (&vS)[4], Emit &&emit) {")
    ap("  switch (l1 * 4 + l2) {")
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("    case %d: if constexpr (LMAX >= %d) block_%d_%d<DO_H, DO_S>(Y, vH, vS, emit); break;" %
               (l1 * 4 + l2, max(l1, l2), l1, l2))
    ap("    default: break;")
    ap("  }")
    ap("}")
    ap("// compile-time dispatch (canonical s/sp/spd/spdf species: the shell loops unroll completely)")
    ap("template <int L1, int L2, bool DO_H, bool DO_S, int NY, class Emit>")
    ap("__device__ __forceinline__ void blockT(const double (&Y)[NY], const double (&vH)[4], const double (&vS)[4], Emit &&emit) {")
    first = True
    for l1 in range(LMAX + 1):
        for l2 in range(LMAX + 1):
            ap("  %sif constexpr (L1 == %d && L2 == %d) block_%d_%d<DO_H, DO_S>(Y, vH, vS, emit);" %
               ("" if first else "else ", l1, l2, l1, l2))
            first = False
    ap("}")
    ap("}}  // namespace")
    # the pool goes in front of the first use
    at = L.index("namespace dftbp_b200 { namespace skg {") + 1
    decl = ["// coefficient pool (Gaunt coefficients and V -> A_L maps)",
            "static __constant__ double skgC[%d] = {" % len(pool)]
    for i in range(0, len(pool), 4):
        decl.append("  " + ", ".join("%.17e" % v for v in pool[i:i + 4]) + ",")
    decl.append("};")
    L[at:at] = decl
    with open(OUT_CODE, "w") as fh:
        fh.write("\n".join(L) + "\n")
    return len(L)


if __name__ == "__main__":
    tb = build()
    err = self_check(tb)
    import sys
    # the table form (term lists for a table-driven evaluator) is only written on request: the
    # library uses the generated straight-line code
    n = emit(tb, write="--tables" in sys.argv)
    nl = emit_code(tb)
    print("straight-line code: %d lines -> %s" % (nl, os.path.normpath(OUT_CODE)))
    print("self-check max err %.2e; %d Gaunt terms%s" % (err, n, " -> " + os.path.normpath(OUT) if "--tables" in sys.argv else ""))

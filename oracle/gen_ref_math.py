#!/usr/bin/env python3
"""Machine-translate the reference's interpolation and tolerance-sort routines to C.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).

The reference is Fortran and this image has no Fortran compiler.  Four routines of the hot path are
small, self-contained loop code:
    poly5ToZero, freeCubicSpline, polyInterU2     src/dftbp/math/interpolation.F90:31-239
    index_heap_sort                               src/dftbp/math/sorting.F90:142-199
This script reads them where they lie under /root/reference, translates them statement by statement
(assignments, do / do while loops, if blocks, array sections over the first dimension -> an explicit
loop) into C under oracle/_ref/math_ref.c (git-ignored; never committed) and compiles
oracle/_ref/libmathref.so.  That library is the reference's own arithmetic and control flow for these
routines; tests/test_oracle_math_ref.py checks the oracle's restatements (oracle/dftb_oracle.c) against
it and against the golden vectors tests/golden/math_ref_golden.npz made from it (`--golden`), which
travel to machines without /root/reference.

No reference source text is copied into the repository.
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INTERP = "/root/reference/src/dftbp/math/interpolation.F90"
REF_SORT = "/root/reference/src/dftbp/math/sorting.F90"
OUT_DIR = os.path.join(HERE, "_ref")


def routine_body(src, kind, name):
    m = re.search(r"\b%s\s+%s\s*\(.*?\n(.*?)\bend\s+%s\s+%s\b" % (kind, name, kind, name), src, re.S | re.I)
    if not m:
        raise RuntimeError("routine %s not found" % name)
    return m.group(1)


def logical_lines(body):
    """comments stripped, continuation lines joined, declarations / assertions dropped"""
    out, cur = [], ""
    for raw in body.splitlines():
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        if cur:
            line = cur + " " + line.strip().lstrip("&")
            cur = ""
        if line.rstrip().endswith("&"):
            cur = line.rstrip()[:-1]
            continue
        s = line.strip()
        if re.match(r"(real|integer|logical)\b.*::", s, re.I) or s.startswith("@:") or s.lower().startswith("pure ") \
                or s.lower().startswith("result"):
            continue
        out.append(s)
    return out


class Tx:
    """Expression / statement rewriting for one routine.
    arrays1: 1-D arrays name -> C name (1-based index -> [i-1]); arrays2: 2-D arrays whose FIRST dimension is
    the vector dimension (name(:,i) -> name[k + nv*(i-1)]); vectors: names of length-nv vectors (bare name ->
    name[k]); optional_out: optional intent(out) scalars passed as pointers."""

    def __init__(self, arrays1=(), arrays2=(), vectors=(), optional_out=(), subst=None):
        self.a1, self.a2, self.vec, self.opt = set(arrays1), set(arrays2), set(vectors), set(optional_out)
        self.subst = subst or {}

    @staticmethod
    def _match_paren(s, i):
        depth = 0
        for j in range(i, len(s)):
            if s[j] == "(":
                depth += 1
            elif s[j] == ")":
                depth -= 1
                if depth == 0:
                    return j
        raise ValueError("unbalanced: " + s)

    def expr(self, s):
        s = s.strip()
        for a, b in ((r"\.gt\.", ">"), (r"\.lt\.", "<"), (r"\.ge\.", ">="), (r"\.le\.", "<="), (r"\.eqv\.", "=="),
                     (r"\.and\.", "&&"), (r"\.or\.", "||"), (r"\.not\.", "!")):
            s = re.sub(a, b, s, flags=re.I)
        s = re.sub(r"(\d+\.\d*|\d+)_dp", r"\1", s)
        s = re.sub(r"(?<![\w.])(\d+\.)(?!\d)", r"\g<1>0", s)
        for a, b in self.subst.items():
            s = re.sub(a, b, s)
        out, i, uses_k = "", 0, False
        while i < len(s):
            m = re.match(r"[A-Za-z_]\w*", s[i:])
            if not m:
                out += s[i]
                i += 1
                continue
            name = m.group(0)
            j = i + len(name)
            k = j
            while k < len(s) and s[k] == " ":
                k += 1
            if k < len(s) and s[k] == "(" and (name in self.a1 or name in self.a2 or name in ("present", "ceiling", "abs")):
                e = self._match_paren(s, k)
                inner = s[k + 1:e]
                if name == "present":
                    out += "(%s != 0)" % inner.strip()
                elif name == "ceiling":
                    sub, uk = self.expr(inner)
                    uses_k |= uk
                    out += "((int)ceil(%s))" % sub
                elif name == "abs":
                    sub, uk = self.expr(inner)
                    uses_k |= uk
                    out += "fabs(%s)" % sub
                elif name in self.a2:
                    first, second = self._split_args(inner)
                    if first.strip() != ":":
                        raise ValueError("2-D array only as a(:,i): " + s)
                    sub, uk = self.expr(second)
                    uses_k = True
                    out += "%s[k + nv * ((%s) - 1)]" % (name, sub)
                else:
                    sub, uk = self.expr(inner)
                    uses_k |= uk
                    out += "%s[(%s) - 1]" % (name, sub)
                i = e + 1
                continue
            if name in self.vec:
                out += name + "[k]"
                uses_k = True
            elif name in self.opt:
                out += "(*%s)" % name
            else:
                out += name
            i = j
        return out, uses_k

    @staticmethod
    def _split_args(inner):
        depth = 0
        for j, ch in enumerate(inner):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 0:
                return inner[:j], inner[j + 1:]
        raise ValueError("expected two subscripts: " + inner)

    def assignment(self, lhs, rhs, ind):
        lhs = lhs.strip()
        # whole-array statements
        if lhs in self.a2:
            r, _ = self.expr(rhs) if rhs.strip() not in self.a2 else (rhs.strip(), False)
            if rhs.strip() in self.a2:
                return ["%sfor (int q = 0; q < nv * nn; ++q) %s[q] = %s[q];" % (ind, lhs, rhs.strip())]
            return ["%sfor (int q = 0; q < nv * nn; ++q) %s[q] = %s;" % (ind, lhs, r)]
        if lhs in self.a1 and "(" not in lhs:
            r, _ = self.expr(rhs)
            return ["%sfor (int q = 0; q < n_%s; ++q) %s[q] = %s;" % (ind, lhs, lhs, r)]
        l, uk1 = self.expr(lhs)
        r, uk2 = self.expr(rhs)
        if uk1 or uk2:
            return ["%sfor (int k = 0; k < nv; ++k) %s = %s;" % (ind, l, r)]
        return ["%s%s = %s;" % (ind, l, r)]

    def body(self, lines, ret="return;"):
        out, ind = [], "  "
        for s in lines:
            low = s.lower()
            m = re.match(r"if\s*\((.*)\)\s*then$", s, re.I)
            if m:
                out.append("%sif (%s) {" % (ind, self.expr(m.group(1))[0]))
                ind += "  "
                continue
            m = re.match(r"else\s*if\s*\((.*)\)\s*then$", s, re.I)
            if m:
                ind = ind[:-2]
                out.append("%s} else if (%s) {" % (ind, self.expr(m.group(1))[0]))
                ind += "  "
                continue
            if low == "else":
                ind = ind[:-2]
                out.append(ind + "} else {")
                ind += "  "
                continue
            if re.match(r"end\s*(if|do)$", low):
                ind = ind[:-2]
                out.append(ind + "}")
                continue
            m = re.match(r"do\s+while\s*\((.*)\)$", s, re.I)
            if m:
                out.append("%swhile (%s) {" % (ind, self.expr(m.group(1))[0]))
                ind += "  "
                continue
            m = re.match(r"do\s+(\w+)\s*=\s*(.*)$", s, re.I)
            if m:
                lo, hi = self._split_args(m.group(2))
                out.append("%sfor (%s = %s; %s <= %s; ++%s) {" % (ind, m.group(1), self.expr(lo)[0], m.group(1),
                                                               self.expr(hi)[0], m.group(1)))
                ind += "  "
                continue
            if low == "return":
                out.append(ind + ret)
                continue
            m = re.match(r"if\s*(\(.*)$", s, re.I)
            if m:                       # one-line if
                e = self._match_paren(m.group(1), 0)
                cond, stmt = m.group(1)[1:e], m.group(1)[e + 1:].strip()
                inner = self.body([stmt], ret)
                out.append("%sif (%s) { %s }" % (ind, self.expr(cond)[0], " ".join(x.strip() for x in inner)))
                continue
            m = re.match(r"([^=]+?)=(?!=)(.*)$", s)
            if m:
                out.extend(self.assignment(m.group(1), m.group(2), ind))
                continue
            raise ValueError("cannot translate: " + s)
        return out


def translate():
    isrc, ssrc = open(REF_INTERP).read(), open(REF_SORT).read()
    out = ["/* GENERATED by oracle/gen_ref_math.py from the reference's math/interpolation.F90 and math/sorting.F90",
           " * -- do not commit. */", "#include <math.h>", "#include <float.h>", ""]
    # poly5ToZero: straight-line function
    tx = Tx()
    lines = logical_lines(routine_body(isrc, "function", "poly5ToZero"))
    out.append("double mref_poly5_to_zero(double y0, double y0p, double y0pp, double xx, double dx, double invdx)\n{")
    out.append("  double dx1, dx2, dd, ee, ff, xr, yy;")
    out.extend(tx.body(lines))
    out.append("  return yy;\n}\n")
    # freeCubicSpline: optional outputs as pointers (0 = absent)
    tx = Tx(optional_out=("yy", "yp", "ypp"))
    lines = logical_lines(routine_body(isrc, "subroutine", "freeCubicSpline"))
    out.append("void mref_free_cubic_spline(double y0, double y0p, double y0pp, double dx, double ydx, double xx,\n"
               "                            double *yy, double *yp, double *ypp)\n{")
    out.append("  double aa, bb, cc, dd, dx1;")
    body = tx.body(lines)
    # present(yy) was rewritten through the optional-pointer rule: (*yy) != 0 must test the pointer
    body = [b.replace("((*yy) != 0)", "(yy != 0)").replace("((*yp) != 0)", "(yp != 0)").replace("((*ypp) != 0)", "(ypp != 0)")
            for b in body]
    out.extend(body)
    out.append("}\n")
    # polyInterU2: vector dimension nv = size(yp, dim=1), nn = size(xp)
    tx = Tx(arrays1=("xp", "delta"), arrays2=("yp", "cc", "dd"), vectors=("yy", "dyy", "r2Tmp"),
            subst={r"size\(xp\)": "nn"})
    lines = [ln for ln in logical_lines(routine_body(isrc, "function", "polyInterU2"))
             if not re.match(r"nn\s*=\s*size", ln) and "present(dy)" not in ln and not re.match(r"dy\s*=", ln)]
    # the `if (present(dy)) then ... end if` block lost its head and body above: drop its `end if` too
    depth, kept = 0, []
    for ln in lines:
        low = ln.lower()
        if re.match(r"(if\b.*then$|do\b)", low):
            depth += 1
        if re.match(r"end\s*(if|do)$", low):
            if depth == 0:
                continue
            depth -= 1
        kept.append(ln)
    out.append("/* yp, cc, dd: (nv, nn) column-major like the reference's yp(:,ii) */")
    out.append("void mref_poly_inter_u2(int nn, const double *xp, int nv, const double *yp, double xx, double *yy)\n{")
    out.append("  double cc[64 * 16], dd[64 * 16], dyy[64], r2Tmp[64], delta[16];")
    out.append("  int iCl, ii, mm; const int n_delta = nn - 1;")
    out.extend(tx.body(kept))
    out.append("}\n")
    # index_heap_sort with the tolerance given
    tx = Tx(arrays1=("indx", "array"), subst={r"size\(indx\)": "n0", r"size\(array\)": "n0", r"present\(tolerance\)": "1",
                                               r"epsilon\(0\.0\)": "DBL_EPSILON"})
    lines = logical_lines(routine_body(ssrc, "subroutine", "index_heap_sort"))
    out.append("void mref_index_heap_sort(int n0, int *indx, const double *array, double tolerance)\n{")
    out.append("  int n, ir, ij, il, ii, ik, indxTmp; double arrayTmp, tol;")
    out.extend(tx.body(lines))
    out.append("}\n")
    os.makedirs(OUT_DIR, exist_ok=True)
    cpath = os.path.join(OUT_DIR, "math_ref.c")
    with open(cpath, "w") as fh:
        fh.write("\n".join(out) + "\n")
    so = os.path.join(OUT_DIR, "libmathref.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, cpath, "-lm"])
    return so


def load(so):
    import ctypes
    lib = ctypes.CDLL(so)
    d, dp, ip = ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.mref_poly5_to_zero.restype = d
    lib.mref_poly5_to_zero.argtypes = [d] * 6
    lib.mref_free_cubic_spline.argtypes = [d] * 6 + [dp, dp, dp]
    lib.mref_poly_inter_u2.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp, d, dp]
    lib.mref_index_heap_sort.argtypes = [ctypes.c_int, ip, dp, d]
    return lib


def golden_inputs():
    """the inputs of the golden vectors (shared with the test)"""
    import numpy as np
    rng = np.random.default_rng(4242)
    g = {}
    g["p5_in"] = rng.normal(size=(40, 3))                      # y0, y0p, y0pp
    g["p5_x"] = rng.uniform(0.0, 1.0, size=40)                 # xx / dx
    g["fcs_in"] = rng.normal(size=(40, 4))                     # y0, y0p, y0pp, ydx
    g["fcs_x"] = rng.uniform(0.0, 0.02, size=40)
    g["u2_xp0"] = rng.uniform(0.5, 9.0, size=30)
    g["u2_yp"] = rng.normal(size=(30, 5, 8))                   # [case, integral, point]
    g["u2_frac"] = rng.uniform(0.0, 7.0, size=30)              # xx = xp0 + frac * 0.02
    vals = []
    for _ in range(20):
        n = int(rng.integers(1, 60))
        base = rng.uniform(1.0, 100.0, size=max(1, n // 3))
        v = rng.choice(base, size=n) + rng.choice([0.0, 3e-11, -4e-11, 9e-11], size=n)   # clusters within tolSameDist2
        vals.append(v)
    g["sort_len"] = np.array([len(v) for v in vals])
    g["sort_vals"] = np.concatenate(vals)
    return g


def golden(so):
    import ctypes
    import numpy as np
    lib = load(so)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    g = golden_inputs()
    dx = 0.3
    g["p5_out"] = np.array([lib.mref_poly5_to_zero(a[0], a[1], a[2], x * dx, dx, 1.0 / dx) for a, x in zip(g["p5_in"], g["p5_x"])])
    fo = np.zeros((40, 3))
    for k, (a, x) in enumerate(zip(g["fcs_in"], g["fcs_x"])):
        y, yp, ypp = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        lib.mref_free_cubic_spline(a[0], a[1], a[2], 0.02, a[3], x, ctypes.byref(y), ctypes.byref(yp), ctypes.byref(ypp))
        fo[k] = (y.value, yp.value, ypp.value)
    g["fcs_out"] = fo
    uo = np.zeros((30, 5))
    for k in range(30):
        xp = g["u2_xp0"][k] + 0.02 * np.arange(8)
        yp = np.asfortranarray(g["u2_yp"][k])                 # (nv, nn) column-major
        lib.mref_poly_inter_u2(8, xp.ctypes.data_as(dp), 5, yp.ctypes.data_as(dp), xp[0] + 0.02 * g["u2_frac"][k],
                               uo[k].ctypes.data_as(dp))
    g["u2_out"] = uo
    so_out, pos = [], 0
    for n in g["sort_len"]:
        v = np.ascontiguousarray(g["sort_vals"][pos:pos + n])
        idx = np.zeros(n, dtype=np.int32)
        lib.mref_index_heap_sort(int(n), idx.ctypes.data_as(ip), v.ctypes.data_as(dp), 1.0e-10)
        so_out.append(idx)
        pos += n
    g["sort_out"] = np.concatenate(so_out)
    path = os.path.join(HERE, "..", "tests", "golden", "math_ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true", help="also (re)write tests/golden/math_ref_golden.npz")
    a = ap.parse_args()
    if not (os.path.exists(REF_INTERP) and os.path.exists(REF_SORT)):
        print("reference not present; nothing to do", file=sys.stderr)
        sys.exit(0)
    so = translate()
    print("built", so)
    if a.golden:
        golden(so)

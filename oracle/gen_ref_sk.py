#!/usr/bin/env python3
"""Machine-translate the reference's own Slater-Koster shell-pair rotation routines to C.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).

The reference is Fortran and this image has no Fortran compiler, so the real routines cannot
be compiled.  The ten shell-pair routines ss..ff of src/dftbp/dftb/sk.F90:144-945 are however
pure straight-line arithmetic (`hh(i,j) = <expression in ll, mm, nn, sk(k)>`), which this
script parses where the file lies under /root/reference and re-emits, expression by
expression, as C into oracle/_ref/sk_ref.c (git-ignored; never committed), then compiles to
oracle/_ref/libskref.so.  That library is the *real reference arithmetic* for the rotations;
it is used (a) to validate the oracle's own independent rotation (oracle/dftb_oracle.c) and
the CUDA kernels' Gaunt-table rotation, and (b) to generate the committed golden vectors
tests/golden/sk_rot_golden.npz (`--golden`).

No reference source text is copied into the repository: the translation happens at build
time in the container that has /root/reference, and its output stays under oracle/_ref/.
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SK = "/root/reference/src/dftbp/dftb/sk.F90"
OUT_DIR = os.path.join(HERE, "_ref")
ROUTINES = ["ss", "sp", "sd", "sf", "pp", "pd", "pf", "dd", "df", "ff"]

TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eEdD][+-]?\d+)?(?:_dp)?|\d+)|([A-Za-z_]\w*)|(\*\*|[-+*/(),]))")


def tokenize(s):
    pos, out = 0, []
    s = s.strip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            raise ValueError("cannot tokenise %r at %d" % (s, pos))
        num, ident, op = m.groups()
        if num is not None:
            out.append(("num", num.replace("_dp", "").replace("d", "e").replace("D", "e")))
        elif ident is not None:
            out.append(("id", ident))
        else:
            out.append(("op", op))
        pos = m.end()
    return out


class Parser:
    """Fortran arithmetic expression -> C text (same association, integer powers by
    repeated multiplication as gfortran -O2 expands them)."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, val=None):
        tok = self.peek()
        if val is not None and tok[1] != val:
            raise ValueError("expected %r got %r" % (val, tok))
        self.i += 1
        return tok

    def expr(self):
        tok = self.peek()
        if tok == ("op", "-") or tok == ("op", "+"):
            self.take()
            left = "(%s%s)" % (tok[1], self.term())
        else:
            left = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            left = "(%s %s %s)" % (left, op, self.term())
        return left

    def term(self):
        left = self.power()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()[1]
            left = "(%s %s %s)" % (left, op, self.power())
        return left

    def power(self):
        base = self.primary()
        if self.peek() == ("op", "**"):
            self.take()
            kind, val = self.take()
            # the reference also writes ll**2.0_dp; a whole-number real exponent has the same
            # value (gfortran folds pow(x, 2.0) to x*x)
            if kind != "num" or float(val) != int(float(val)):
                raise ValueError("only whole-number literal powers expected")
            return "ipow%d(%s)" % (int(float(val)), base)
        return base

    def primary(self):
        kind, val = self.take()
        if kind == "num":
            return val if ("." in val or "e" in val) else val + ".0"
        if kind == "op" and val == "(":
            e = self.expr()
            self.take(")")
            return e
        if kind == "id":
            if self.peek() == ("op", "("):
                self.take()
                args = [self.expr_or_int()]
                while self.peek() == ("op", ","):
                    self.take()
                    args.append(self.expr_or_int())
                self.take(")")
                if val == "sqrt":
                    return "sqrt(%s)" % args[0]
                if val == "sk":
                    return "sk[%d]" % (int(args[0]) - 1)
                if val == "hh":
                    return "hh[%d]" % ((int(args[0]) - 1) + 7 * (int(args[1]) - 1))
                raise ValueError("unexpected call " + val)
            if val in ("ll", "mm", "nn"):
                return val
            raise ValueError("unexpected identifier " + val)
        raise ValueError("unexpected token %r" % ((kind, val),))

    def expr_or_int(self):
        kind, val = self.peek()
        nxt = self.t[self.i + 1] if self.i + 1 < len(self.t) else (None, None)
        if kind == "num" and "." not in val and nxt[1] in (",", ")"):
            self.take()
            return val
        return self.expr()


def translate():
    with open(REF_SK) as fh:
        src = fh.read()
    # join continuation lines
    src = re.sub(r"&\s*\n\s*&", "", src)
    out = ["/* GENERATED by oracle/gen_ref_sk.py from the reference's sk.F90 -- do not commit. */",
           "#include <math.h>",
           "static inline double ipow2(double x){return x*x;}",
           "static inline double ipow3(double x){return x*x*x;}",
           "static inline double ipow4(double x){double y=x*x;return y*y;}",
           "static inline double ipow5(double x){double y=x*x;return y*y*x;}",
           "static inline double ipow6(double x){double y=x*x;return y*y*y;}"]
    for name in ROUTINES:
        m = re.search(r"subroutine %s\(hh,.*?\n(.*?)end subroutine %s\b" % (name, name), src, re.S)
        if not m:
            raise RuntimeError("routine %s not found" % name)
        body = []
        for line in m.group(1).splitlines():
            line = line.strip()
            mm = re.match(r"hh\((\d+),(\d+)\)\s*=\s*(.*)$", line)
            if not mm:
                continue
            i, j, rhs = int(mm.group(1)), int(mm.group(2)), mm.group(3)
            p = Parser(tokenize(rhs))
            c = p.expr()
            if p.i != len(p.t):
                raise ValueError("trailing tokens in %s: %s" % (name, rhs))
            body.append("  hh[%d] = %s;" % ((i - 1) + 7 * (j - 1), c))
        out.append("static void ref_%s(double *hh, double ll, double mm, double nn, const double *sk)\n{" % name)
        out.append("  (void)ll; (void)mm; (void)nn;")
        out.extend(body)
        out.append("}")
    out.append("""
/* hh is a 7x7 column-major scratch; la <= lb selects the routine exactly like the
 * select-case of rotateH0 (sk.F90:82-126). */
void skref_block(int la, int lb, double ll, double mm, double nn, const double *sk, double *hh)
{
  static void (*const tab[4][4])(double *, double, double, double, const double *) = {
    {0, 0, 0, 0}, {0, ref_pp, ref_pd, ref_pf}, {0, 0, ref_dd, ref_df}, {0, 0, 0, ref_ff}};
  int i;
  for (i = 0; i < 49; ++i) hh[i] = 0.0;
  if (la == 0) {
    if (lb == 0) { hh[0] = sk[0]; (void)ref_ss; }
    else if (lb == 1) ref_sp(hh, ll, mm, nn, sk);
    else if (lb == 2) ref_sd(hh, ll, mm, nn, sk);
    else ref_sf(hh, ll, mm, nn, sk);
  } else {
    tab[la][lb](hh, ll, mm, nn, sk);
  }
}
""")
    os.makedirs(OUT_DIR, exist_ok=True)
    cpath = os.path.join(OUT_DIR, "sk_ref.c")
    with open(cpath, "w") as fh:
        fh.write("\n".join(out) + "\n")
    so = os.path.join(OUT_DIR, "libskref.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, cpath, "-lm"])
    return so


def golden(so):
    import ctypes
    import numpy as np
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.skref_block.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, dp, dp]
    rng = np.random.default_rng(777)
    dirs = rng.normal(size=(24, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    special = np.array([[0, 0, 1.0], [0, 0, -1.0], [1.0, 0, 0], [0, 1.0, 0], [-1.0, 0, 0],
                        [0, -1.0, 0], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [1e-9, 0, -1.0]])
    special /= np.linalg.norm(special, axis=1)[:, None]
    dirs = np.vstack([special, dirs])
    res = {"dirs": dirs}
    for la in range(4):
        for lb in range(la, 4):
            sk = rng.normal(size=(dirs.shape[0], la + 1))
            hh = np.zeros((dirs.shape[0], 49))
            for k in range(dirs.shape[0]):
                lib.skref_block(la, lb, dirs[k, 0], dirs[k, 1], dirs[k, 2],
                                sk[k].ctypes.data_as(dp), hh[k].ctypes.data_as(dp))
            res["sk_%d%d" % (la, lb)] = sk
            # stored as [case, col, row] == Fortran hh(row, col)
            res["hh_%d%d" % (la, lb)] = hh.reshape(-1, 7, 7)
    path = os.path.join(HERE, "..", "tests", "golden", "sk_rot_golden.npz")
    np.savez_compressed(path, **res)
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true", help="also (re)write tests/golden/sk_rot_golden.npz")
    a = ap.parse_args()
    if not os.path.exists(REF_SK):
        print("reference not present; nothing to do", file=sys.stderr)
        sys.exit(0)
    so = translate()
    print("built", so)
    if a.golden:
        golden(so)

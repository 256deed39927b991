"""Run loop routines of the reference from their own source text: a Fortran-subset -> Python (NumPy) transpiler.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference is Fortran and this image has no Fortran compiler.
The sparse-layout routines of the hot path (unpackHS / packHS / mullikenPerOrbital / addShift ...) are short loop
nests over whole-array statements -- array sections, reshape, sum(dim=), matmul -- which plain C translation
(gen_ref_math.py) does not cover.  This module reads such a routine where it lies under /root/reference, rewrites it
statement by statement into Python that operates on NumPy arrays with Fortran index semantics (FArr: lower bounds,
inclusive sections, column-major reshape) and executes it.  No reference source text is stored in the repository:
the golden vectors made with it are (tests/golden/, gen_ref_sparse.py).

Covered subset: do / do while / if blocks, scalar and array-section assignments, reshape, sum(.., dim=), matmul,
size, dot_product, exp / cos / sin / sqrt / abs / real / cmplx / conjg / transpose, derived-type components (a%b),
implied-do array constructors of the form [(i, i = a, b)], allocate of 1-D index arrays.  Calls are mapped by the
caller (call_map); assertions and preprocessor lines are dropped."""
import math
import re

import numpy as np


class LoopExit(Exception):
    def __init__(self, label):
        self.label = label


class LoopCycle(Exception):
    def __init__(self, label):
        self.label = label


class FArr:
    """NumPy array with Fortran lower bounds.  a[i, j] -> scalar; any slice -> plain ndarray view/copy (sections are
    1-based afterwards, as in Fortran); a[...] = value assigns with Fortran conformance (scalar broadcast or same shape)."""

    def __init__(self, data, lower=None):
        self.a = data
        self.lb = tuple(lower) if lower is not None else (1,) * data.ndim

    def _index(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        assert len(idx) == self.a.ndim, "rank mismatch"
        out = []
        for k, (i, lb) in enumerate(zip(idx, self.lb)):
            if isinstance(i, slice):
                lo = lb if i.start is None else int(i.start)
                hi = lb + self.a.shape[k] - 1 if i.stop is None else int(i.stop)
                assert lo >= lb and hi <= lb + self.a.shape[k] - 1 or hi < lo, "section out of bounds"
                out.append(slice(lo - lb, max(hi - lb + 1, lo - lb), None if i.step is None else int(i.step)))
            elif isinstance(i, (np.ndarray, FArr, list)):          # vector subscript
                v = np.asarray(i.a if isinstance(i, FArr) else i, dtype=np.int64) - lb
                assert v.size == 0 or (v.min() >= 0 and v.max() < self.a.shape[k]), "vector subscript out of bounds"
                out.append(v)
            else:
                j = int(i) - lb
                assert 0 <= j < self.a.shape[k], "index out of bounds"
                out.append(j)
        return tuple(out)

    def __getitem__(self, idx):
        v = self.a[self._index(idx)]
        return (v.item() if hasattr(v, "item") else v) if np.ndim(v) == 0 else v

    def __setitem__(self, idx, value):
        ix = self._index(idx)
        value = value.a if isinstance(value, FArr) else value
        if np.ndim(value) == 0:
            self.a[ix] = value
        else:
            tgt = self.a[ix]
            assert tgt.shape == np.shape(value), "non-conforming assignment %s <- %s" % (tgt.shape, np.shape(value))
            self.a[ix] = value


def _assign_whole(self, value):
    """whole-array assignment a = expr: element-wise into the existing storage (scalars broadcast); an array of another
    shape replaces the storage (allocatable left-hand sides are reallocated by Fortran 2003 semantics)"""
    value = value.a if isinstance(value, FArr) else value
    if np.ndim(value) == 0 or np.shape(value) == self.a.shape:
        self.a[...] = value
    else:
        self.a = np.array(value, order="F")
        self.lb = (1,) * self.a.ndim


FArr.assign = _assign_whole


def _whole(op):
    return lambda self, other: op(self.a, other.a if isinstance(other, FArr) else other)


# whole-array expressions (a bare array name): element-wise on the data, result a plain ndarray
for _n, _f in (("__mul__", np.multiply), ("__add__", np.add), ("__sub__", np.subtract), ("__truediv__", np.divide)):
    setattr(FArr, _n, _whole(_f))
for _n, _f in (("__rmul__", np.multiply), ("__radd__", np.add)):
    setattr(FArr, _n, _whole(_f))
FArr.__rsub__ = lambda self, other: np.subtract(other, self.a)
FArr.__neg__ = lambda self: -self.a
FArr.__pow__ = lambda self, n: self.a ** n
FArr.__eq__ = lambda self, other: self.a == (other.a if isinstance(other, FArr) else other)
FArr.__ne__ = lambda self, other: self.a != (other.a if isinstance(other, FArr) else other)
FArr.__hash__ = lambda self: id(self)
FArr.__array__ = lambda self, dtype=None, copy=None: self.a if dtype is None else self.a.astype(dtype)


def f_reshape(x, shape):
    x = x.a if isinstance(x, FArr) else np.asarray(x)
    shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
    return np.reshape(x, shape, order="F")


def f_sum(x, dim=None):
    x = x.a if isinstance(x, FArr) else np.asarray(x)
    if dim is None:
        # the intrinsic is a serial loop in array-element (column-major) order
        acc = 0.0 if not np.iscomplexobj(x) else 0j
        for v in np.ravel(x, order="F").tolist():
            acc = acc + v
        return acc
    return np.sum(x, axis=int(dim) - 1)


def f_size(x, dim=None):
    x = x.a if isinstance(x, FArr) else np.asarray(x)
    return int(x.size) if dim is None else int(x.shape[int(dim) - 1])


def f_matmul(a, b):
    """small products as the compiler's inline loops: serial sum over the contracted index, a product and an addition
    per term (no fused multiply-add -- x86-64 baseline); large ones through BLAS"""
    a = np.asarray(a.a if isinstance(a, FArr) else a)
    b = np.asarray(b.a if isinstance(b, FArr) else b)
    if a.size * (b.shape[1] if b.ndim == 2 else 1) > 20000 or np.iscomplexobj(a) or np.iscomplexobj(b):
        return np.matmul(a, b)
    al, bl = a.tolist(), b.tolist()
    if b.ndim == 1:
        return np.array([_dot_serial(row, bl) for row in (al if a.ndim == 2 else [al])]) if a.ndim == 2 else _dot_serial(al, bl)
    cols = list(zip(*bl))
    if a.ndim == 1:
        return np.array([_dot_serial(al, c) for c in cols])
    return np.array([[_dot_serial(row, c) for c in cols] for row in al])


def _dot_serial(x, y):
    acc = 0.0
    for u, v in zip(x, y):
        acc = acc + u * v
    return acc


def f_dot(a, b):
    a = a.a if isinstance(a, FArr) else a
    b = b.a if isinstance(b, FArr) else b
    s = 0.0
    for x, y in zip(np.ravel(a), np.ravel(b)):   # serial sum like the intrinsic on short vectors
        s = s + x * y
    return s


def f_cmplx(re_, im_=0.0, kind=None):
    return complex(re_, im_)


def f_real(x, kind=None):
    return np.real(x) if np.ndim(x) else float(np.real(x))


def _exp(x):
    """exp with IEEE overflow to +Inf (no trap), like the compiled reference"""
    with np.errstate(over="ignore"):
        v = np.exp(x)
    return v if np.ndim(v) else (complex(v) if isinstance(x, complex) else float(v))


ENV = dict(np=np, math=math, FArr=FArr, f_reshape=f_reshape, f_sum=f_sum, f_size=f_size, f_matmul=f_matmul, f_dot=f_dot,
           dp=None, f_cmplx=f_cmplx, f_real=f_real, f_exp=lambda x: _exp(x),
           f_conjg=np.conj, f_transpose=np.transpose, pi=3.14159265358979323846264338327950288, imag=1j)

INTRINSICS = {"reshape": "f_reshape", "sum": "f_sum", "size": "f_size", "matmul": "f_matmul", "dot_product": "f_dot",
              "cmplx": "f_cmplx", "real": "f_real", "exp": "f_exp", "cos": "math.cos", "sin": "math.sin", "sqrt": "f_sqrt", "log": "math.log", "floor": "f_floor",
              "abs": "f_abs", "minval": "f_minval", "maxval": "f_maxval", "minloc": "f_minloc", "norm2": "f_norm2", "epsilon": "f_epsilon", "sign": "math.copysign", "present": "f_present", "len": "len", "shape": "f_shape", "product": "f_product", "all": "f_all", "any": "f_any", "ubound": "f_ubound", "nint": "f_nint", "conjg": "f_conjg", "transpose": "f_transpose", "max": "max", "min": "min", "mod": "f_mod"}
ENV["f_epsilon"] = lambda x: 2.220446049250313e-16
def _f_assign(target, value):
    """whole-array assignment; an unallocated allocatable (None) is allocated with the shape of the value"""
    if isinstance(target, FArr):
        target.assign(value)
        return target
    v = value.a if isinstance(value, FArr) else value
    return FArr(np.array(v, order="F"))


ENV["f_assign"] = _f_assign
ENV["f_present"] = lambda x: x is not None
ENV["f_product"] = lambda x: int(np.prod(x.a if isinstance(x, FArr) else x)) if np.asarray(x.a if isinstance(x, FArr) else x).dtype.kind in "iu" else float(np.prod(x.a if isinstance(x, FArr) else x))
ENV["f_shape"] = lambda x: np.array((x.a if isinstance(x, FArr) else np.asarray(x)).shape)
ENV["f_sqrt"] = lambda x: np.sqrt(x.a if isinstance(x, FArr) else x) if isinstance(x, (FArr, np.ndarray)) else math.sqrt(x)
ENV["f_abs"] = lambda x: np.abs(x.a if isinstance(x, FArr) else x) if isinstance(x, (FArr, np.ndarray)) else abs(x)
ENV["f_minval"] = lambda x, dim=None: (np.min(x.a if isinstance(x, FArr) else x) if dim is None else np.min(x.a if isinstance(x, FArr) else x, axis=int(dim) - 1))
ENV["f_maxval"] = lambda x, dim=None: (np.max(x.a if isinstance(x, FArr) else x) if dim is None else np.max(x.a if isinstance(x, FArr) else x, axis=int(dim) - 1))
ENV["f_minloc"] = lambda x, dim=None: int(np.argmin(x.a if isinstance(x, FArr) else x)) + (x.lb[0] if isinstance(x, FArr) else 1)
ENV["f_norm2"] = lambda x: math.sqrt(f_sum(np.asarray(x.a if isinstance(x, FArr) else x) ** 2))
ENV["f_not"] = lambda x: np.logical_not(x.a if isinstance(x, FArr) else x) if isinstance(x, (FArr, np.ndarray)) else (not x)
ENV["f_all"] = lambda x: bool(np.all(x.a if isinstance(x, FArr) else x))
ENV["f_any"] = lambda x: bool(np.any(x.a if isinstance(x, FArr) else x))
ENV["f_floor"] = lambda x: np.floor(x.a if isinstance(x, FArr) else x).astype(np.int64) if isinstance(x, (FArr, np.ndarray)) else int(math.floor(x))
ENV["f_ubound"] = lambda x, dim: x.lb[int(dim) - 1] + x.a.shape[int(dim) - 1] - 1
ENV["f_nint"] = lambda x: int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))
ENV["LoopExit"], ENV["LoopCycle"] = LoopExit, LoopCycle
# integer mod: a first operand that came out of an integer division arrives as a float here -- truncate it like Fortran
ENV["f_mod"] = lambda a, b: int(a) % int(b) if isinstance(b, (int, np.integer)) else math.fmod(a, b)


def _logical_lines(text):
    out, cur, skip = [], "", []     # skip: stack of booleans (True = lines are dropped)
    state = {"q": None}     # quote character still open (a string literal continued over lines)

    def strip_comment(raw):
        q = state["q"]
        for i, ch in enumerate(raw):
            if q:
                if ch == q:
                    q = None
            elif ch in "\"'":
                q = ch
            elif ch == "!":
                state["q"] = None
                return raw[:i]
        state["q"] = q
        return raw

    for raw in text.splitlines():
        line = strip_comment(raw).rstrip()
        st = line.strip()
        if st.startswith("#:if") or st.startswith("#:block"):
            skip.append(True)        # every build option (WITH_MPI, EXP_TRAP, DEBUG_CODE ...) is off
            continue
        if st.startswith("#:else"):
            skip[-1] = not skip[-1]
            continue
        if st.startswith("#:endif") or st.startswith("#:endblock"):
            skip.pop()
            continue
        if any(skip):
            continue
        if not line.strip() or line.strip().startswith("#:") or line.strip().startswith("@:"):
            continue
        s = line.strip()
        if cur:
            s = cur + " " + s.lstrip("&").strip()
            cur = ""
        if s.endswith("&"):
            cur = s[:-1].rstrip()
            continue
        out.append(s)
    return out


def _match(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] in "([":
            depth += 1
        elif s[j] in ")]":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced: " + s)


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def _fix_eqv(s):
    """a .eqv. b  ->  ((a)) == ((b)): the relational operators inside a and b bind tighter in Fortran, while Python would
    chain the comparisons.  Operands end at the enclosing parenthesis or at a .and. / .or. of the same depth."""
    while True:
        m = re.search(r"\.eqv\.", s, re.I)
        if not m:
            return s
        i, depth = m.start() - 1, 0
        while i >= 0:
            if s[i] == ")":
                depth += 1
            elif s[i] == "(":
                if depth == 0:
                    break
                depth -= 1
            elif depth == 0 and re.match(r"\.(and|or)\.", s[max(0, i - 4):i + 1][-5:], re.I):
                break
            i -= 1
        left = i + 1
        j, depth = m.end(), 0
        while j < len(s):
            if s[j] == "(":
                depth += 1
            elif s[j] == ")":
                if depth == 0:
                    break
                depth -= 1
            elif depth == 0 and re.match(r"\.(and|or)\.", s[j:j + 5], re.I):
                break
            j += 1
        s = s[:left] + "((" + s[left:m.start()] + ")) == ((" + s[m.end():j] + "))" + s[j:]


def _expr(s, arrays):
    s = _fix_eqv(s.strip())
    s = re.sub(r"\.not\.\s*(\w+\([^()]*\))", r"f_not(\1)", s, flags=re.I)
    s = re.sub(r"\.not\.\s*(\w+)(?!\s*[\w(])", r"f_not(\1)", s, flags=re.I)
    for a, b in ((r"\.gt\.", ">"), (r"\.lt\.", "<"), (r"\.ge\.", ">="), (r"\.le\.", "<="), (r"\.eq\.", "=="), (r"\.ne\.", "!="),
                 (r"\.and\.", " and "), (r"\.or\.", " or "), (r"\.not\.", " not "), (r"\.true\.", "True"), (r"\.false\.", "False")):
        s = re.sub(a, b, s, flags=re.I)
    s = re.sub(r"\.not\.\s*(\w+\([^()]*\))", r"f_not(\1)", s, flags=re.I)
    s = re.sub(r"\.not\.\s*(\w+)(?!\s*\()", r"f_not(\1)", s, flags=re.I)
    s = s.replace("/=", "!=").replace("(/", "[").replace("/)", "]")
    s = re.sub(r"epsilon\(1\.0_rsp\)", "1.1920928955078125e-07", s)
    s = re.sub(r"(\d+\.\d*(?:[eE][-+]?\d+)?|\d+)_dp", r"\1", s)
    s = re.sub(r"\bdim\s*=\s*", "", s)
    out, i = "", 0
    while i < len(s):
        m = re.match(r"[A-Za-z_]\w*(?:%[A-Za-z_]\w*)*", s[i:])
        if not m:
            out += s[i]
            i += 1
            continue
        name = m.group(0)
        j = i + len(name)
        k = j
        while k < len(s) and s[k] == " ":
            k += 1
        pyname = name.replace("%", ".")
        base = name.split("%")[-1] if "%" in name else name
        if k < len(s) and s[k] == "(":
            e = _match(s, k)
            inner = _expr(s[k + 1:e], arrays)
            if name.lower() in INTRINSICS and "%" not in name:
                out += "%s(%s)" % (INTRINSICS[name.lower()], inner)
            elif name in arrays or base in arrays:
                out += "%s[%s]" % (pyname, inner)
            else:
                out += "%s(%s)" % (pyname, inner)
            i = e + 1
            continue
        out += pyname
        i = j
    return out.replace("]%", "].").replace(")%", ").")        # component of an indexed element: a(i)%b


def transpile(src, kind, name, call_map=None, extra_arrays=()):
    """-> (python source of def <name>(args), argument list).  call_map: 'callee' -> python statement template using
    {args} (list of translated arguments); unknown calls raise."""
    m = re.search(r"\b%s\s+%s\s*\(([^)]*)\)[^\n]*\n(.*?)\bend\s+%s\s+%s\b" % (kind, name, kind, name), src, re.S | re.I)
    if not m:
        raise RuntimeError("routine %s not found" % name)
    args = [a.strip() for a in m.group(1).replace("&", " ").replace("\n", " ").split(",")]
    lines = _logical_lines(m.group(2))
    # Fortran is case-insensitive: every identifier takes its first spelling (dummy arguments first)
    canon = {}
    for tok in args + re.findall(r"[A-Za-z_]\w*", "\n".join(lines)):
        canon.setdefault(tok.lower(), tok)
    lines = [re.sub(r"[A-Za-z_]\w*", lambda t: canon[t.group(0).lower()], ln) for ln in lines]
    arrays, local_alloc, body, lower, dtypes, int_scalars, declared_arrays = set(extra_arrays), [], [], {}, {}, set(), set()
    for ln in lines:
        d = re.match(r"(real\(dp\)|integer|complex\(dp\)|logical|type\(\w+\)|class\(\w+\)|character\([^)]*\))\s*(,[^:]*)?::\s*(.*)$", ln, re.I)
        d0 = re.match(r"(integer|logical)\s+(\w+(?:\s*,\s*\w+)*)$", ln, re.I)
        if d0 and not d:        # old-style declaration without '::' (scalars)
            if d0.group(1).lower() == "integer":
                int_scalars.update(x.strip() for x in d0.group(2).split(","))
            continue
        if d:
            typ, attrs, rest = d.group(1).lower(), (d.group(2) or "").lower(), d.group(3)
            # split declared entities at top-level commas
            ents, depth, cur = [], 0, ""
            for ch in rest:
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                if ch == "," and depth == 0:
                    ents.append(cur.strip())
                    cur = ""
                else:
                    cur += ch
            ents.append(cur.strip())
            for e in ents:
                pm = re.match(r"(\w+)\s*=\s*(.*)$", e)
                if pm and "parameter" in attrs:        # named constant with its value
                    local_alloc.append("%s = %s" % (pm.group(1), _expr(pm.group(2), arrays)))
                    continue
                pa = re.match(r"(\w+)\s*\(([^=]*)\)\s*=\s*(.*)$", e)
                if pa and "parameter" in attrs:        # constant array: name(bounds) = reshape(...)
                    arrays.add(pa.group(1))
                    los = [_expr(x.split(":")[0], arrays) if ":" in x else "1" for x in _split_top(pa.group(2))]
                    local_alloc.append("%s = FArr(np.array(%s, order='F'), (%s,))" % (pa.group(1), _expr(pa.group(3), arrays), ", ".join(los)))
                    continue
                em = re.match(r"(\w+)\s*\((.*)\)$", e)
                if not em:
                    if typ == "integer" and re.match(r"\w+$", e):
                        int_scalars.add(e)
                    continue
                arrays.add(em.group(1))
                if "allocatable" not in attrs or "intent" in attrs:
                    declared_arrays.add(em.group(1))
                dtypes[em.group(1)] = "complex" if typ.startswith("complex") else ("int" if typ == "integer" else "float")
                lower[em.group(1)] = [int(x.split(":")[0]) if ":" in x and x.split(":")[0].strip() else 1
                                      for x in _split_top(em.group(2))] if "intent" in attrs else None
                if "intent" not in attrs and "allocatable" not in attrs and "pointer" not in attrs:
                    shp, los = [], []
                    for x in _split_top(em.group(2)):
                        if ":" in x:
                            lo_, hi_ = x.split(":")
                            los.append(_expr(lo_, arrays))
                            shp.append("(%s) - (%s) + 1" % (_expr(hi_, arrays), _expr(lo_, arrays)))
                        else:
                            los.append("1")
                            shp.append(_expr(x, arrays))
                    dt = "complex" if typ.startswith("complex") else ("int" if typ == "integer" else ("bool" if typ == "logical" else "float"))
                    local_alloc.append("%s = FArr(np.zeros((%s,), dtype=%s, order='F'), (%s,))" % (em.group(1), ", ".join(shp), dt, ", ".join(los)))
            continue
        body.append(ln)
    py, ind, selects, loops, wheres = ["def %s(%s):" % (name, ", ".join(a_ + "=None" for a_ in args))], "    ", [], [], []
    for a_ in args:
        if lower.get(a_) is not None:
            py.append("%s%s = %s if isinstance(%s, FArr) or not isinstance(%s, np.ndarray) else FArr(%s, %r)" % (ind, a_, a_, a_, a_, a_, tuple(lower[a_])))
    py += [ind + x for x in local_alloc]
    for s in body:
        low = s.lower()
        mm = re.match(r"if\s*\((.*)\)\s*then$", s, re.I)
        if mm:
            py.append("%sif %s:" % (ind, _expr(mm.group(1), arrays)))
            ind += "    "
            continue
        mm = re.match(r"else\s*if\s*\((.*)\)\s*then$", s, re.I)
        if mm:
            ind = ind[:-4]
            py.append("%selif %s:" % (ind, _expr(mm.group(1), arrays)))
            ind += "    "
            continue
        if low == "else":
            ind = ind[:-4]
            py.append(ind + "else:")
            ind += "    "
            continue
        em_ = re.match(r"end\s*do(\s+\w+)?$", low)
        if em_:
            lab = loops.pop() if loops else None
            py.append(ind + "pass")
            if lab:
                ind = ind[:-4]
                py.append("%sexcept LoopCycle as e_:" % ind)
                py.append("%s    if e_.label.lower() != %r: raise" % (ind, lab.lower()))
                ind = ind[:-4]
                ind = ind[:-4]
                py.append("%sexcept LoopExit as e_:" % ind)
                py.append("%s    if e_.label.lower() != %r: raise" % (ind, lab.lower()))
            else:
                ind = ind[:-4]
            continue
        if re.match(r"end\s*if$", low):
            py.append(ind + "pass")
            ind = ind[:-4]
            continue
        mm = re.match(r"where\s*(\(.*)$", s, re.I)
        if mm:
            e_ = _match(mm.group(1), 0)
            cond_, rest_ = mm.group(1)[1:e_], mm.group(1)[e_ + 1:].strip()
            if rest_:       # one-line form: where (mask) a = b
                am = re.match(r"([^=]+?)=(?!=)(.*)$", rest_)
                tgt = am.group(1).strip()
                tgt_e = "%s.a" % tgt if tgt in declared_arrays else _expr(tgt, arrays)
                py.append("%snp.copyto(%s, %s, where=%s)" % (ind, tgt_e, _expr(am.group(2), arrays), _expr(cond_, arrays)))
            else:
                wheres.append(_expr(cond_, arrays))
            continue
        if re.match(r"end\s*where$", low):
            wheres.pop()
            continue
        if wheres:
            am = re.match(r"([^=]+?)=(?!=)(.*)$", s)
            lhs_, rhs_ = _expr(am.group(1), arrays), _expr(am.group(2), arrays)
            py.append("%snp.copyto(%s, %s, where=%s)" % (ind, lhs_, rhs_, wheres[-1]))      # masked assignment
            continue
        mm = re.match(r"select\s+case\s*\((.*)\)$", s, re.I)
        if mm:
            selects.append([_expr(mm.group(1), arrays), True])
            continue
        mm = re.match(r"case\s*\((.*)\)$", s, re.I)
        if mm or low == "case default":
            sel = selects[-1]
            if not sel[1]:
                ind = ind[:-4]
            if mm:
                py.append("%s%s %s == %s:" % (ind, "if" if sel[1] else "elif", sel[0], _expr(mm.group(1), arrays)))
            else:
                py.append(ind + ("if True:" if sel[1] else "else:"))
            sel[1] = False
            ind += "    "
            continue
        if re.match(r"end\s*select$", low):
            sel = selects.pop()
            if not sel[1]:
                py.append(ind + "pass")
                ind = ind[:-4]
            continue
        mm = re.match(r"([\w%]+)\s*=>\s*(.*)$", s)
        if mm:   # pointer assignment: a view / the same object
            py.append("%s%s = %s" % (ind, mm.group(1).replace("%", "."), _expr(mm.group(2), arrays)))
            continue
        if low == "do":
            loops.append(None)
            py.append(ind + "while True:")
            ind += "    "
            continue
        mm = re.match(r"do\s+while\s*\((.*)\)$", s, re.I)
        if mm:
            loops.append(None)
            py.append("%swhile %s:" % (ind, _expr(mm.group(1), arrays)))
            ind += "    "
            continue
        label = None
        lm = re.match(r"(\w+)\s*:\s*(do\b.*)$", s, re.I)
        if lm:
            label, s = lm.group(1), lm.group(2)
            low = s.lower()
        mm = re.match(r"do\s+(\w+)\s*=\s*(.*)$", s, re.I)
        if mm:
            if label:
                py.append(ind + "try:")
                ind += "    "
            loops.append(label)
            parts = _split_top(mm.group(2))
            lo, hi = _expr(parts[0], arrays), _expr(parts[1], arrays)
            if len(parts) == 3:
                st = _expr(parts[2], arrays)
                py.append("%sfor %s in range(%s, (%s) + (1 if (%s) > 0 else -1), %s):" % (ind, mm.group(1), lo, hi, st, st))
            else:
                py.append("%sfor %s in range(%s, (%s) + 1):" % (ind, mm.group(1), lo, hi))
            ind += "    "
            if label:
                py.append(ind + "try:")
                ind += "    "
            continue
        if re.match(r"deallocate\s*\(", low):
            py.append(ind + "pass")
            continue
        if low == "return":
            py.append(ind + "return locals()")
            continue
        if low in ("exit", "cycle"):
            py.append(ind + ("break" if low == "exit" else "continue"))
            continue
        xm = re.match(r"(exit|cycle)\s+(\w+)$", s, re.I)
        if xm:
            py.append("%sraise %s(%r)" % (ind, "LoopExit" if xm.group(1).lower() == "exit" else "LoopCycle", xm.group(2)))
            continue
        if low.startswith("write(") or low.startswith("write ("):
            py.append(ind + "pass")
            continue
        mm = re.match(r"call\s+(\w+(?:%\w+)*)\s*\((.*)\)$", s, re.I)
        if mm:
            callee = mm.group(1)
            cargs = [_expr(p, arrays) for p in _split_top(mm.group(2))]
            if not call_map or callee not in call_map:
                raise ValueError("unmapped call: " + s)
            tmpl = call_map[callee]
            py.append(ind + (tmpl(cargs) if callable(tmpl) else tmpl.format(args=cargs)))
            continue
        mm = re.match(r"allocate\s*\(\s*([\w%]+)\s*\)$", s, re.I)
        if mm:       # a derived-type scalar
            py.append("%s%s = new_object(%r)" % (ind, mm.group(1).replace("%", "."), mm.group(1).split("%")[-1]))
            continue
        s = re.sub(r",\s*source\s*=\s*[^()]*\)$", ")", s)        # allocate(a(n), source=0.0_dp): arrays are zeroed anyway
        mm = re.match(r"allocate\s*\(\s*([\w%]+)\s*\((.*)\)\s*\)$", s, re.I)
        if mm:
            arrays.add(mm.group(1).split("%")[-1])
            los, shp = [], []
            for dspec in _split_top(mm.group(2)):
                if ":" in dspec:
                    lo_, hi_ = dspec.split(":")
                    los.append(_expr(lo_, arrays))
                    shp.append("(%s) - (%s) + 1" % (_expr(hi_, arrays), _expr(lo_, arrays)))
                else:
                    los.append("1")
                    shp.append(_expr(dspec, arrays))
            py.append("%s%s = FArr(np.zeros((%s,), dtype=%s, order='F'), (%s,))" % (ind, mm.group(1).replace("%", "."), ", ".join(shp),
                                                                                dtypes.get(mm.group(1), "int"), ", ".join(los)))
            continue
        mm = re.match(r"(\w+)\s*\(\s*:\s*\)\s*=\s*\[\s*\(\s*(\w+)\s*,\s*\2\s*=\s*(.*?),\s*(.*?)\)\s*\]$", s)
        if mm:   # implied do: a(:) = [(i, i = lo, hi)]
            py.append("%s%s[:] = np.arange(%s, (%s) + 1)" % (ind, mm.group(1), _expr(mm.group(3), arrays), _expr(mm.group(4), arrays)))
            continue
        mm = re.match(r"if\s*(\(.*)$", s, re.I)
        if mm:  # one-line if
            e = _match(mm.group(1), 0)
            cond, stmt = mm.group(1)[1:e], mm.group(1)[e + 1:].strip()
            cm_ = re.match(r"call\s+(\w+)\s*\((.*)\)$", stmt, re.I)
            if cm_:
                cargs = [_expr(p, arrays) for p in _split_top(cm_.group(2))]
                if not call_map or cm_.group(1) not in call_map:
                    raise ValueError("unmapped call: " + s)
                py.append("%sif %s: %s" % (ind, _expr(cond, arrays), call_map[cm_.group(1)].format(args=cargs)))
                continue
            xm = re.match(r"(exit|cycle)(\s+\w+)?$", stmt, re.I)
            if xm:
                if xm.group(2):
                    act = "raise %s(%r)" % ("LoopExit" if xm.group(1).lower() == "exit" else "LoopCycle", xm.group(2).strip())
                else:
                    act = "break" if xm.group(1).lower() == "exit" else "continue"
                py.append("%sif %s: %s" % (ind, _expr(cond, arrays), act))
                continue
            am = re.match(r"([^=]+?)=(?!=)(.*)$", stmt)
            py.append("%sif %s: %s = %s" % (ind, _expr(cond, arrays), _expr(am.group(1), arrays), _expr(am.group(2), arrays)))
            continue
        mm = re.match(r"([^=]+?)=(?!=)(.*)$", s)
        if mm:
            rhs = _expr(mm.group(2), arrays)
            if mm.group(1).strip() in int_scalars and "/" in rhs:
                rhs = "int(%s)" % rhs     # integer division / real -> integer conversion: truncation (non-negative operands)
            if mm.group(1).strip() in declared_arrays:
                py.append("%s%s = f_assign(%s, %s)" % (ind, mm.group(1).strip(), mm.group(1).strip(), rhs))       # whole-array assignment
                continue
            py.append("%s%s = %s" % (ind, _expr(mm.group(1), arrays), rhs))
            continue
        raise ValueError("cannot translate: " + s)
    py.append("    return locals()")
    transpile.last_lower = {k: v for k, v in lower.items() if v is not None}
    return "\n".join(py) + "\n", args


def compile_routine(src, kind, name, call_map=None, extra_env=None, extra_arrays=()):
    code, args = transpile(src, kind, name, call_map, extra_arrays)
    env = dict(ENV)
    if extra_env:
        env.update(extra_env)
    exec(compile(code, "<%s from the reference>" % name, "exec"), env)
    fn = env[name]
    fn.lower = dict(transpile.last_lower)
    return fn, args, code


def wrap(fn, name, array):
    """NumPy array (Fortran dimension order) -> FArr with the lower bounds the routine declares for dummy `name`."""
    return FArr(array, fn.lower.get(name))

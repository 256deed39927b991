"""Slater-Koster parameter files -> equidistant tables for dftbp_b200_set_sk_table (host side).

Covers the reference's init-time path readSKFiles -> TOldSKData_readFromFile -> getFullTable
(dftbplus/parser.F90:3668-3886,4061-4130; type/oldskdata.F90:83-202): the classic 10+10 column
format and the extended ('@' first line) 20+20 column format, the column maps iSKInterOld and
skMap, and the rule that for l1 > l2 the column comes from the B-A file.  Also writes synthetic
parameter sets in the real file format: the reference's own sets (mio-1-1, pbc-0-3) are not in
the tree (external/slakos is not fetched) and there is no network.
"""
import os

import numpy as np

AMU_TO_AU = 1822.888486209  # amu__au, common/constants.F90

# column of the 20-wide table holding integral (l_min, l_max, m): parser.F90:4088-4094 skMap.
# 1-based as in the reference: ff(sigma..phi)=1..4, df=5..7, dd=8..10, pf=11..12, pd=13..14,
# pp=15..16, sf=17, sd=18, sp=19, ss=20.
SK_MAP = {(3, 3): [1, 2, 3, 4], (2, 3): [5, 6, 7], (2, 2): [8, 9, 10], (1, 3): [11, 12],
          (1, 2): [13, 14], (1, 1): [15, 16], (0, 3): [17], (0, 2): [18], (0, 1): [19], (0, 0): [20]}
# classic files carry only these 10 columns, in this order (oldskdata.F90:77 iSKInterOld)
I_SK_INTER_OLD = [8, 9, 10, 13, 14, 15, 16, 18, 19, 20]


class OldSKData:
    """TOldSKData: dist, nGrid, skHam/skOver[nGrid, 20], skSelf/skHubbU/skOcc[4] (s,p,d,f), mass."""

    def __init__(self):
        self.dist = 0.0
        self.nGrid = 0
        self.skHam = self.skOver = None
        self.skSelf = np.zeros(4)
        self.skHubbU = np.zeros(4)
        self.skOcc = np.zeros(4)
        self.mass = 0.0


def _numbers(line):
    """Fortran list-directed reads accept commas, and n*value repeat counts."""
    out = []
    for tok in line.replace(",", " ").split():
        if "*" in tok:
            n, v = tok.split("*")
            out.extend([float(v.replace("d", "e").replace("D", "e"))] * int(n))
        else:
            out.append(float(tok.replace("d", "e").replace("D", "e")))
    return out


def read_skf(path, homo):
    """TOldSKData_readFromFile (electronic part; the spline/polynomial repulsive is not on the hot path)."""
    sk = OldSKData()
    with open(path) as fh:
        lines = fh.read().splitlines()
    pos = 0
    extended = lines[0].lstrip().startswith("@")
    n_shell = 4 if extended else 3
    if extended:
        pos = 1
    head = _numbers(lines[pos]); pos += 1
    sk.dist, sk.nGrid = head[0], int(head[1]) - 1       # oldskdata.F90:143-145
    if homo:
        v = _numbers(lines[pos]); pos += 1
        # (skSelf(ii), ii = nShell, 1, -1), SPE, (skHubbU(ii), ...), (skOcc(ii), ...)
        for k in range(n_shell):
            sk.skSelf[n_shell - 1 - k] = v[k]
            sk.skHubbU[n_shell - 1 - k] = v[n_shell + 1 + k]
            sk.skOcc[n_shell - 1 - k] = v[2 * n_shell + 1 + k]
        v = _numbers(lines[pos]); pos += 1
        sk.mass = v[0] * AMU_TO_AU
    else:
        pos += 1
    sk.skHam = np.zeros((sk.nGrid, 20))
    sk.skOver = np.zeros((sk.nGrid, 20))
    for g in range(sk.nGrid):
        v = _numbers(lines[pos]); pos += 1
        if extended:
            sk.skHam[g, :] = v[:20]
            sk.skOver[g, :] = v[20:40]
        else:
            for k, col in enumerate(I_SK_INTER_OLD):
                sk.skHam[g, col - 1] = v[k]
                sk.skOver[g, col - 1] = v[10 + k]
    return sk


def n_sk_integrals(shells1, shells2):
    return sum(min(l1, l2) + 1 for l1 in shells1 for l2 in shells2)


def get_full_table(sk12, sk21, shells1, shells2):
    """getFullTable (parser.F90:4061-4130): (skHam, skOver)[nGrid, nInteg] order='F' for the ordered
    species pair (1, 2); sk12 = file '1-2', sk21 = file '2-1'."""
    n = n_sk_integrals(shells1, shells2)
    ham = np.zeros((sk12.nGrid, n), order="F")
    over = np.zeros((sk12.nGrid, n), order="F")
    ind = 0
    for l1 in shells1:
        for l2 in shells2:
            src = sk12 if l1 <= l2 else sk21
            cols = SK_MAP[(min(l1, l2), max(l1, l2))]
            for mm in range(min(l1, l2) + 1):
                ham[:, ind] = src.skHam[:, cols[mm] - 1]
                over[:, ind] = src.skOver[:, cols[mm] - 1]
                ind += 1
    return ham, over


class SlakoSet:
    """All tables of a parameter set for the given species: what readSKFiles leaves in
    TSlakoCont (H and S), plus on-site energies, Hubbard U and reference occupations."""

    def __init__(self, directory, names, ang_shells, truncation_cutoff=None, method=2):
        # method: 2 = skEqGridNew (default), 1 = skEqGridOld (OldSKInterpolation = Yes), slakoeqgrid.F90:62-65
        self.method = method
        self.names, self.ang_shells = list(names), [list(a) for a in ang_shells]
        nsp = len(names)
        files = {}
        for i, a in enumerate(names):
            for j, b in enumerate(names):
                files[(i, j)] = read_skf(os.path.join(directory, "%s-%s.skf" % (a, b)), homo=(i == j))
        self.tables = {}
        for i in range(nsp):
            for j in range(nsp):
                ham, over = get_full_table(files[(i, j)], files[(j, i)], self.ang_shells[i], self.ang_shells[j])
                dist = files[(i, j)].dist
                if truncation_cutoff is not None:     # parser.F90:3833-3838
                    n_ent = min(int(np.floor(truncation_cutoff / dist)), ham.shape[0])
                    ham, over = np.asfortranarray(ham[:n_ent]), np.asfortranarray(over[:n_ent])
                self.tables[(i + 1, j + 1)] = (dist, ham, over)
        self.selfegy = [[files[(i, i)].skSelf[l] for l in self.ang_shells[i]] for i in range(nsp)]
        self.hubbU = [[files[(i, i)].skHubbU[l] for l in self.ang_shells[i]] for i in range(nsp)]
        self.occ = [[files[(i, i)].skOcc[l] for l in self.ang_shells[i]] for i in range(nsp)]
        self.mass = [files[(i, i)].mass for i in range(nsp)]
        self.n_orb_species = np.array([sum(2 * l + 1 for l in a) for a in self.ang_shells], dtype=np.int32)

    def cutoff(self):
        fudge = 0.3 if self.method == 1 else 1.0     # distFudgeOld / distFudge, slakoeqgrid.F90:162-179
        return max(t[1].shape[0] * t[0] + fudge for t in self.tables.values())

    def upload(self, hot):
        """set_orbitals + set_sk_table for every ordered pair on a HotPath (or an oracle Model-like
        object with the same two methods)."""
        hot.set_orbitals(self.ang_shells)
        for (i, j), (dist, ham, over) in self.tables.items():
            hot.set_sk_table(0, i, j, dist, ham, self.method)
            hot.set_sk_table(1, i, j, dist, over, self.method)


# ---------------------------------------------------------------------------------------------------
# synthetic parameter sets in the real .skf format
# ---------------------------------------------------------------------------------------------------

def write_synthetic_set(directory, names, ang_shells, onsite, hubbard, occ, mass, dist=0.02, n_grid=500,
                        seed=20250101, extended=None):
    """Write A-B.skf for all ordered pairs.  Every integral column is A (1 + a r) exp(-b r) with
    per-column constants drawn from numpy's default_rng(seed); overlap columns are kept small enough
    that S stays diagonally dominant for the bench geometries (checked by the tests).
    onsite/hubbard/occ: per species dicts {l: value}."""
    os.makedirs(directory, exist_ok=True)
    rng = np.random.default_rng(seed)
    r = dist * np.arange(1, n_grid + 1)
    done = {}
    force_extended = extended
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            # the extended 20-column format is only needed when an f shell is involved
            extended = force_extended if force_extended is not None else (3 in ang_shells[i] or 3 in ang_shells[j])
            ncol = 20 if extended else 10
            order = list(range(1, 21)) if extended else I_SK_INTER_OLD
            def cols(scale_lo, scale_hi, b_lo, b_hi):
                amp = rng.uniform(scale_lo, scale_hi, ncol) * rng.choice([-1.0, 1.0], ncol)
                aa = rng.uniform(0.2, 0.6, ncol)
                bb = rng.uniform(b_lo, b_hi, ncol)
                return amp[None, :] * (1.0 + aa[None, :] * r[:, None]) * np.exp(-bb[None, :] * r[:, None])
            ham = cols(0.1, 0.5, 0.8, 1.2)
            over = cols(0.15, 0.4, 1.0, 1.4)
            # columns of shells the species do not have are zero in real files too; l-l columns of
            # B-A equal those of A-B (the same two-centre integral seen from the other end)
            ls_a, ls_b = set(ang_shells[i]), set(ang_shells[j])
            for k, col in enumerate(order):
                (lmin, lmax) = next(key for key, v in SK_MAP.items() if col in v)
                present = lmin in ls_a and lmax in ls_b   # first orbital on A, second on B
                if not present:
                    ham[:, k] = 0.0
                    over[:, k] = 0.0
                elif lmin == lmax and i > j:
                    ham[:, k] = done[(j, i)][0][:, k]
                    over[:, k] = done[(j, i)][1][:, k]
            done[(i, j)] = (ham, over)
            path = os.path.join(directory, "%s-%s.skf" % (a, b))
            with open(path, "w") as fh:
                if extended:
                    fh.write("@\n")
                fh.write("%.12g %d\n" % (dist, n_grid + 1))
                nsh = 4 if extended else 3
                if i == j:
                    e = [onsite[i].get(l, 0.0) for l in range(nsh - 1, -1, -1)]
                    u = [hubbard[i].get(l, 0.0) for l in range(nsh - 1, -1, -1)]
                    f = [occ[i].get(l, 0.0) for l in range(nsh - 1, -1, -1)]
                    fh.write(" ".join("%.12g" % v for v in e + [0.0] + u + f) + "\n")
                    fh.write(" ".join("%.12g" % v for v in [mass[i]] + [0.0] * 8 + [1.0] + [0.0] * 10) + "\n")
                else:
                    fh.write(" ".join("%.12g" % v for v in [0.0] * 9 + [1.0] + [0.0] * 10) + "\n")
                for g in range(n_grid):
                    fh.write(" ".join("%.17e" % v for v in np.concatenate([ham[g], over[g]])) + "\n")
                # minimal, valid (zero) spline repulsive so the file is complete
                fh.write("Spline\n1 1.0\n0.0 1.0 0.0\n0.5 1.0 0.0 0.0 0.0 0.0 0.0 0.0\n")
    return directory

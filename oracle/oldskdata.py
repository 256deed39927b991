"""Oracle restatement of the reference's SK-file path -- TEST INFRASTRUCTURE ONLY.

Follows TOldSKData_readFromFile (src/dftbp/type/oldskdata.F90:83-202) and getFullTable
(src/dftbp/dftbplus/parser.F90:4061-4130) loop for loop, independently of the product's
dftbplus_b200/slako.py, so the two can be checked against each other.
"""
import numpy as np

MAXL = 3
N_SK_INTER = 20
I_SK_INTER_OLD = [8, 9, 10, 13, 14, 15, 16, 18, 19, 20]          # oldskdata.F90:77

# skMap(0:maxL, 0:maxL, 0:maxL), Fortran reshape order (mm fastest, then lMax, then lMin);
# parser.F90:4088-4094
_SK_MAP_FLAT = [20, 0, 0, 0, 19, 0, 0, 0, 18, 0, 0, 0, 17, 0, 0, 0,
                0, 0, 0, 0, 15, 16, 0, 0, 13, 14, 0, 0, 11, 12, 0, 0,
                0, 0, 0, 0, 0, 0, 0, 0, 8, 9, 10, 0, 5, 6, 7, 0,
                0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4]


def sk_map(mm, l_max, l_min):
    return _SK_MAP_FLAT[mm + (MAXL + 1) * l_max + (MAXL + 1) ** 2 * l_min]


def read_from_file(file_name, homo):
    """returns dict(dist, nGrid, skHam[nGrid,20], skOver[nGrid,20], skSelf[4], skHubbU[4], skOcc[4])"""
    with open(file_name) as fh:
        lines = fh.read().split("\n")
    it = iter(lines)
    first = next(it)
    if first[:1] == "@":
        extended, n_shell = True, 4
        first = next(it)
    else:
        extended, n_shell = False, 3
    tok = first.replace(",", " ").split()
    dist, n_grid = float(tok[0]), int(tok[1])
    n_grid = n_grid - 1                                              # oldskdata.F90:145
    sk_self, sk_hubb, sk_occ = np.zeros(4), np.zeros(4), np.zeros(4)
    if homo:
        v = [float(x) for x in next(it).replace(",", " ").split()]
        k = 0
        for ii in range(n_shell, 0, -1):
            sk_self[ii - 1] = v[k]; k += 1
        k += 1                                                       # rPlaceholder (spin pol. energy)
        for ii in range(n_shell, 0, -1):
            sk_hubb[ii - 1] = v[k]; k += 1
        for ii in range(n_shell, 0, -1):
            sk_occ[ii - 1] = v[k]; k += 1
        next(it)                                                     # mass + polynomial repulsive
    else:
        next(it)
    sk_ham = np.zeros((n_grid, N_SK_INTER))
    sk_over = np.zeros((n_grid, N_SK_INTER))
    for i_grid in range(n_grid):
        v = [float(x) for x in next(it).replace(",", " ").split()]
        if extended:
            for ii in range(N_SK_INTER):
                sk_ham[i_grid, ii] = v[ii]
                sk_over[i_grid, ii] = v[N_SK_INTER + ii]
        else:
            for ii in range(10):
                sk_ham[i_grid, I_SK_INTER_OLD[ii] - 1] = v[ii]
                sk_over[i_grid, I_SK_INTER_OLD[ii] - 1] = v[10 + ii]
    return dict(dist=dist, nGrid=n_grid, skHam=sk_ham, skOver=sk_over, skSelf=sk_self,
                skHubbU=sk_hubb, skOcc=sk_occ)


def get_full_table(sk12, sk21, ang_shells1, ang_shells2):
    n = sum(min(l1, l2) + 1 for l1 in ang_shells1 for l2 in ang_shells2)
    sk_ham = np.zeros((sk12["nGrid"], n), order="F")
    sk_over = np.zeros((sk12["nGrid"], n), order="F")
    ind = 0
    for l1 in ang_shells1:
        for l2 in ang_shells2:
            if l1 <= l2:
                p_ham, p_over, l_min, l_max = sk12["skHam"], sk12["skOver"], l1, l2
            else:
                p_ham, p_over, l_min, l_max = sk21["skHam"], sk21["skOver"], l2, l1
            for mm in range(l_min + 1):
                sk_ham[:, ind] = p_ham[:, sk_map(mm, l_max, l_min) - 1]
                sk_over[:, ind] = p_over[:, sk_map(mm, l_max, l_min) - 1]
                ind += 1
    return sk_ham, sk_over


def load_model(directory, names, ang_shells, method=2):
    """readSKFiles (parser.F90:3668-3886) into an oracle Model; returns (model, selfegy)."""
    import os
    from . import oracle
    model = oracle.Model(ang_shells)
    files = {(i, j): read_from_file(os.path.join(directory, "%s-%s.skf" % (a, b)), i == j)
             for i, a in enumerate(names) for j, b in enumerate(names)}
    for i in range(len(names)):
        for j in range(len(names)):
            ham, over = get_full_table(files[(i, j)], files[(j, i)], ang_shells[i], ang_shells[j])
            model.set_table(0, i + 1, j + 1, files[(i, j)]["dist"], ham, method)
            model.set_table(1, i + 1, j + 1, files[(i, j)]["dist"], over, method)
    selfegy = [[files[(i, i)]["skSelf"][l] for l in ang_shells[i]] for i in range(len(names))]
    return model, selfegy

"""Synthetic inputs for the BASELINE.json configs (geometry + SK parameter set), host side.

There is no network, so mio-1-1 / pbc-0-3 cannot be fetched: parameter sets are synthetic files
in the real .skf format (slako.write_synthetic_set; 0.02 bohr x 500 grid, rng 20250101), and
geometries are generated from fixed seeds.
"""
import os
import tempfile

import numpy as np

from . import geometry, slako

ANG = 1.0 / 0.529177249  # Angstrom -> bohr (AA__Bohr, common/constants.F90)

SPECIES = {
    # name: (shells, onsite {l: E}, hubbard {l: U}, occupation {l: n}, mass/amu)
    "H": ([0], {0: -0.2386}, {0: 0.4195}, {0: 1.0}, 1.008),
    "O": ([0, 1], {0: -0.8788, 1: -0.3321}, {0: 0.4954, 1: 0.4954}, {0: 2.0, 1: 4.0}, 15.999),
    "C": ([0, 1], {0: -0.5049, 1: -0.1944}, {0: 0.3647, 1: 0.3647}, {0: 2.0, 1: 2.0}, 12.011),
    "Si": ([0, 1], {0: -0.3957, 1: -0.1503}, {0: 0.2476, 1: 0.2476}, {0: 2.0, 1: 2.0}, 28.086),
    "Fe": ([0, 1, 2], {0: -0.1577, 1: -0.0230, 2: -0.2563}, {0: 0.2005, 1: 0.1470, 2: 0.3668},
           {0: 2.0, 1: 0.0, 2: 6.0}, 55.845),
    # f-shell species for the s/p/d/f rotation tests (not one of the BASELINE configs)
    "Ce": ([0, 1, 2, 3], {0: -0.13, 1: -0.03, 2: -0.10, 3: -0.25}, {0: 0.2, 1: 0.15, 2: 0.25, 3: 0.45},
           {0: 2.0, 1: 0.0, 2: 1.0, 3: 1.0}, 140.116),
}


def make_slako_set(names, directory=None, n_grid=500, dist=0.02, seed=20250101, method=2):
    """Write (once) and read a synthetic parameter set for the named species."""
    if directory is None:
        directory = os.path.join(tempfile.gettempdir(), "dftbp_b200_slakos",
                                 "-".join(names) + "_%d_%g_%d" % (n_grid, dist, seed))
    shells = [SPECIES[n][0] for n in names]
    marker = os.path.join(directory, ".complete")
    if not os.path.exists(marker):
        # several ranks of a torchrun job may get here at once: every process writes its own private copy and
        # renames it into place; whoever loses the race throws its copy away (rename of a directory is atomic)
        import shutil
        tmp = "%s.tmp.%d" % (directory.rstrip(os.sep), os.getpid())
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(os.path.dirname(directory.rstrip(os.sep)) or ".", exist_ok=True)
        slako.write_synthetic_set(tmp, names, shells, [SPECIES[n][1] for n in names],
                                  [SPECIES[n][2] for n in names], [SPECIES[n][3] for n in names],
                                  [SPECIES[n][4] for n in names], dist=dist, n_grid=n_grid, seed=seed)
        open(os.path.join(tmp, ".complete"), "w").close()
        try:
            os.rename(tmp, directory)
        except OSError:
            if os.path.exists(marker):
                shutil.rmtree(tmp, ignore_errors=True)      # another process was first
            else:
                directory = tmp     # an incomplete directory of an interrupted run is in the way: use the private copy
    return slako.SlakoSet(directory, names, shells, method=method), directory


def _water_molecule():
    # geometry of test/app/dftb+/spin/H2O/dftb_in.hsd (Angstrom)
    return np.array([[0.0, -1.0, 0.0], [0.0, 0.0, 0.78306400000], [0.0, 0.0, -0.78306400000]]) * ANG


def h2o():
    """config 1: H2O molecule, non-periodic."""
    return dict(name="h2o", names=["O", "H"], coord0=_water_molecule(), species0=np.array([1, 2, 2]),
                latvec=None, kpoints=None, n_spin=1)


def water_box(n_mol=334, seed=1):
    """configs 2 / 5: n_mol H2O on a jittered cubic lattice at 0.0334 molecules / A^3, periodic."""
    rng = np.random.default_rng(seed)
    box = (n_mol / 0.0334) ** (1.0 / 3.0) * ANG
    m = int(np.ceil(n_mol ** (1.0 / 3.0)))
    sites = np.array([[i, j, k] for i in range(m) for j in range(m) for k in range(m)], dtype=np.float64)
    sites = sites[rng.permutation(len(sites))[:n_mol]]
    mol = _water_molecule()
    mol = mol - mol[0]
    coords, species = [], []
    for s in sites:
        # random orientation + small jitter of the centre
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        a, b, c, d = q
        R = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                      [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                      [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
        centre = (s + 0.5 + rng.uniform(-0.08, 0.08, 3)) * (box / m)
        coords.append(centre + mol @ R.T)
        species += [1, 2, 2]
    return dict(name="water%d" % (3 * n_mol), names=["O", "H"], coord0=np.vstack(coords),
                species0=np.array(species), latvec=np.eye(3) * box, kpoints=None, n_spin=1)


def si_diamond(n_cells=4, mp=(4, 4, 4)):
    """config 3: Si diamond n x n x n conventional cells (a = 5.431 A), Monkhorst-Pack mesh."""
    a = 5.431 * ANG
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(n_cells) for j in range(n_cells) for k in range(n_cells)])
    coord0 = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    kpts, wts = geometry.monkhorst_pack(mp)
    return dict(name="si%d" % len(coord0), names=["Si"], coord0=coord0,
                species0=np.ones(len(coord0), dtype=np.int32), latvec=np.eye(3) * a * n_cells,
                kpoints=kpts, kweights=wts, n_spin=1)


def fe_c(cells=(16, 16, 8), c_fraction=0.25, seed=2):
    """config 4: bcc Fe supercell (a = 2.87 A) with a fraction of the sites -> C, jitter 0.02 A,
    collinear spin, Gamma point.  (16,16,8) -> 4096 atoms."""
    rng = np.random.default_rng(seed)
    a = 2.87 * ANG
    basis = np.array([[0, 0, 0], [.5, .5, .5]])
    cc = np.array([[i, j, k] for i in range(cells[0]) for j in range(cells[1]) for k in range(cells[2])])
    coord0 = (cc[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    coord0 = coord0 + rng.normal(scale=0.02 * ANG, size=coord0.shape)
    species0 = np.ones(len(coord0), dtype=np.int32)
    species0[rng.permutation(len(coord0))[:int(round(c_fraction * len(coord0)))]] = 2
    return dict(name="fec%d" % len(coord0), names=["Fe", "C"], coord0=coord0, species0=species0,
                latvec=np.diag(np.array(cells, dtype=np.float64) * a), kpoints=None, n_spin=2)


def prepare(workload, n_grid=500, dist=0.02, seed=20250101, slako_dir=None, method=2):
    """geometry arrays + SlakoSet for a workload dict."""
    sk, directory = make_slako_set(workload["names"], slako_dir, n_grid=n_grid, dist=dist, seed=seed, method=method)
    cutoff = sk.cutoff()
    geo = geometry.build_geometry(workload["coord0"], workload["species0"], sk.n_orb_species,
                                  workload["latvec"], cutoff)
    return sk, geo, directory


BY_NAME = {
    "h2o": h2o,
    "water1002": lambda: water_box(334, 1),
    "water2046": lambda: water_box(682, 3),
    "si512": lambda: si_diamond(4, (4, 4, 4)),
    "fec4096": lambda: fe_c((16, 16, 8)),
    "fec1024": lambda: fe_c((8, 8, 8)),
    "fec128": lambda: fe_c((4, 4, 4)),
}

"""Producer of the sparse index arrays the hot path consumes (host side).

In production these arrays come from the Fortran host (updateNeighbourListAndSpecies,
getNrOfNeighboursForAll, getSparseDescriptor; dftb/periodic.F90:380-694,996-1054,
dftb/sparse2dense.F90:3617-3662) and are passed through dftbp_b200_set_geometry unchanged.
This module produces the same arrays for the harness, benchmarks and the workload generators
with a k-d tree instead of the reference's O(N^2 nCell) scan.  The layout, the image numbering
(cell vector outer, central atom inner, first hit appends) and the half-list rule (only images
whose central atom index >= the atom) are the reference's; the order inside a shell of equal
distances is (distance, image index) where the reference's heap sort leaves it unspecified.
"""
import numpy as np
from scipy.spatial import cKDTree

MIN_NEIGH_DIST = 1.0e-2   # common/accuracy.F90:79
TOL_SAME_DIST2 = 1.0e-10  # common/accuracy.F90:72


def cell_translations(latvec, cutoff):
    """getCellTranslations (dftb/periodic.F90:227-263): origin first, then all integer triples of
    the parallelepiped +-(floor(cutoff*|b_i|)+1), index 3 fastest.  latvec[i] = i-th lattice vector.
    Returns (cellVec[nCell,3] in lattice units, rCellVec[nCell,3] Cartesian)."""
    latvec = np.asarray(latvec, dtype=np.float64)
    rec = np.linalg.inv(latvec).T            # rec[i] = i-th reciprocal vector / 2 pi
    rng = [int(np.floor(cutoff * np.linalg.norm(rec[i]))) + 1 for i in range(3)]
    ax = [np.arange(-r, r + 1) for r in rng]
    grid = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
    grid = grid[np.any(grid != 0, axis=1)]
    cell = np.vstack([np.zeros((1, 3)), grid.astype(np.float64)])
    return cell, cell @ latvec


def build_geometry(coord0, species0, n_orb_species, latvec=None, cutoff=None):
    """All arrays of SURVEY.md section 8.0 for central-cell coordinates coord0[nAtom,3] (bohr),
    species0[nAtom] (1-based), orbitals per species, optional lattice vectors (rows)."""
    coord0 = np.ascontiguousarray(coord0, dtype=np.float64)
    species0 = np.ascontiguousarray(species0, dtype=np.int32)
    n_orb_species = np.asarray(n_orb_species, dtype=np.int32)
    nAtom = coord0.shape[0]
    if latvec is None:
        cell, rcell = np.zeros((1, 3)), np.zeros((1, 3))
    else:
        cell, rcell = cell_translations(latvec, cutoff)
    tree = cKDTree(coord0)
    # candidate images (cell ii, atom iAtom1) near the central cell
    img_cell, img_atom, pairs_a2, pairs_img, pairs_d2 = [], [], [], [], []
    lo, hi = coord0.min(axis=0) - cutoff, coord0.max(axis=0) + cutoff
    nimg = 0
    for ii in range(cell.shape[0]):
        pos = coord0 + rcell[ii]
        inside = np.all((pos >= lo) & (pos <= hi), axis=1)
        cand = np.nonzero(inside)[0]
        if cand.size == 0:
            continue
        lists = tree.query_ball_point(pos[cand], cutoff)
        for a1, neigh in zip(cand, lists):
            neigh = np.asarray(neigh, dtype=np.int64)
            neigh = neigh[neigh <= a1]          # half list: central atom index <= image's atom
            if neigh.size == 0:
                continue
            d = coord0[neigh] - pos[a1]
            d2 = np.einsum("ij,ij->i", d, d)
            keep = d2 <= cutoff * cutoff
            neigh, d2 = neigh[keep], d2[keep]
            if neigh.size == 0:
                continue
            img_cell.append(ii)
            img_atom.append(a1)
            self_mask = (ii == 0) & (neigh == a1)
            close = (d2 < MIN_NEIGH_DIST ** 2) & ~self_mask
            if np.any(close):
                bad = neigh[np.nonzero(close)[0][0]]
                raise RuntimeError("Atoms %d and %d too close together" % (a1 + 1, bad + 1))
            neigh, d2 = neigh[~self_mask], d2[~self_mask]
            pairs_a2.append(neigh)
            pairs_img.append(np.full(neigh.size, nimg, dtype=np.int64))
            pairs_d2.append(d2)
            nimg += 1
    img_cell = np.asarray(img_cell, dtype=np.int64)
    img_atom = np.asarray(img_atom, dtype=np.int64)
    assert np.array_equal(img_atom[:nAtom], np.arange(nAtom)) and np.all(img_cell[:nAtom] == 0)
    nAll = nimg
    coords = coord0[img_atom] + rcell[img_cell]
    a2 = np.concatenate(pairs_a2) if pairs_a2 else np.zeros(0, dtype=np.int64)
    im = np.concatenate(pairs_img) if pairs_img else np.zeros(0, dtype=np.int64)
    d2 = np.concatenate(pairs_d2) if pairs_d2 else np.zeros(0)
    # sort by (atom, distance, image index); distances equal within tolSameDist2 keep image order
    order = np.lexsort((im, d2, a2))
    a2, im, d2 = a2[order], im[order], d2[order]
    counts = np.bincount(a2, minlength=nAtom).astype(np.int32)
    mNeigh = int(counts.max()) if counts.size else 0
    iNeighbour = np.zeros((nAtom, mNeigh + 1), dtype=np.int32)
    neighDist2 = np.zeros((nAtom, mNeigh + 1))
    iNeighbour[:, 0] = np.arange(1, nAtom + 1)
    start = np.concatenate([[0], np.cumsum(counts)])
    slot = np.arange(a2.size) - start[a2] + 1
    iNeighbour[a2, slot] = im + 1
    neighDist2[a2, slot] = d2
    nOrbAtom = np.ascontiguousarray(n_orb_species[species0 - 1], dtype=np.int32)
    img2CentCell = (img_atom + 1).astype(np.int32)
    # getSparseDescriptor: consecutive blocks, atom outer, neighbour inner
    nOrbNeigh = np.where(iNeighbour > 0, nOrbAtom[img2CentCell[np.maximum(iNeighbour, 1) - 1] - 1], 0)
    blk = nOrbNeigh * nOrbAtom[:, None]
    valid = np.arange(mNeigh + 1)[None, :] <= counts[:, None]
    blk = np.where(valid, blk, 0)
    flat = np.cumsum(blk.ravel()) - blk.ravel()
    iPair = np.where(valid, flat.reshape(blk.shape), 0).astype(np.int32)
    sparseSize = int(blk.sum())
    iAtomStart = np.ones(nAtom + 1, dtype=np.int32)
    iAtomStart[1:] = 1 + np.cumsum(nOrbAtom)
    return dict(nAtom=nAtom, nAllAtom=nAll, mNeigh=mNeigh, coords=np.ascontiguousarray(coords),
                species=np.ascontiguousarray(species0[img_atom]), species0=species0,
                img2CentCell=img2CentCell, iCellVec=(img_cell + 1).astype(np.int32), cellVec=cell,
                rCellVec=rcell, nNeighbourSK=counts, iNeighbour=iNeighbour, iPair=iPair,
                iAtomStart=iAtomStart, sparseSize=sparseSize, nOrbAtom=nOrbAtom,
                nOrb=int(iAtomStart[-1] - 1), cutoff=cutoff, latvec=latvec, neighDist2=neighDist2)


def monkhorst_pack(n, shift=(0.5, 0.5, 0.5)):
    """Supercell-folding k-points of a diagonal n1 x n2 x n3 mesh with the reference's shift
    convention and inversion reduction (getSuperSampling, dftb/periodic.F90:1152-1278).
    Returns (kpoints[nK,3] in relative coordinates, weights[nK] summing to 1)."""
    n = np.asarray(n, dtype=np.int64)
    pts = []
    for i in range(n[0]):
        for j in range(n[1]):
            for k in range(n[2]):
                p = (np.array([i, j, k], dtype=np.float64) + np.asarray(shift)) / n
                p = p - np.floor(p)            # fold to [0, 1)
                pts.append(p)
    pts = np.array(pts)
    w = np.full(len(pts), 1.0 / len(pts))
    # reduce by inversion: k and -k (mod 1) are equivalent
    keep = np.ones(len(pts), dtype=bool)
    for a in range(len(pts)):
        if not keep[a]:
            continue
        inv = (-pts[a]) % 1.0
        for b in range(a + 1, len(pts)):
            if keep[b]:
                d = np.abs(pts[b] - inv)
                d = np.minimum(d, 1.0 - d)
                if np.all(d < 1e-4):
                    keep[b] = False
                    w[a] += w[b]
                    break
    return pts[keep], w[keep]

// n2_neighbour_kernels.cuh -- neighbour list and sparse descriptor on the device.
//
// Reference: updateNeighbourList (dftb/periodic.F90:446-694), getNrOfNeighbours (:1026-1054, the
// bisection search_asc_real_geq of math/binarysearch.F90:63-94), getSparseDescriptor
// (dftb/sparse2dense.F90:3617-3662), index_heap_sort (math/sorting.F90:142-199).
//
// The reference scans all nCellVec * nAtom * nAtom / 2 (image, atom) combinations on one thread.
// Here the central-cell atoms are binned into a uniform grid (bin edge >= cutoff); the CANDIDATE
// IMAGES (cell vector ii, central atom iAtom1) that lie within one bin edge of the central cell are
// compacted into a list, and eight threads per listed image share the atoms of the 27 bins around it.
// Everything the reference's loop order decides is reproduced from the candidate number
// c = ii * nAtom + iAtom1, which is the position of the image in that loop:
//   * image numbers  = exclusive prefix sum over c of "has a neighbour iAtom2 <= iAtom1 in range";
//   * the unsorted neighbour rows are in ascending image number (the loops append them that way);
//   * the error pair is the first (c, iAtom2) in loop order that is closer than minNeighDist.
// The distance order: the reference's heap sort compares with a tolerance, which is not transitive,
// so inside shells of equal distance only the same algorithm on the same input order gives the same
// order (and with it the same iPair layout).  Rows without two distances within the tolerance have
// a unique order and are placed by counting, fully in parallel (nl_rank); rows with near-ties run
// the reference's heap sort with its exact control flow (nl_sort).
//
// The kernels are grid-stride loops with atomics only (no shared memory, no warp collectives), so
// that tests/host_eval/host_nl.cpp can run the very same source on the CPU, one emulated thread
// after the other in a shuffled order, against the oracle (no GPU in the development container).
#pragma once

#include "emu_kernel.cuh"

namespace dftbp_b200 {
namespace nl {

typedef unsigned long long ull;

constexpr int kMaxBinsPerDim = 32;
constexpr int kMaxBins = kMaxBinsPerDim * kMaxBinsPerDim * kMaxBinsPerDim;
constexpr double kTolSameDist2 = 1.0e-5 * 1.0e-5;   // tolSameDist**2, common/accuracy.F90:68-72
constexpr double kMinNeighDist2 = 1.0e-2 * 1.0e-2;  // minNeighDist**2, common/accuracy.F90:77-82
constexpr double kBinSlack = 1.0001;
constexpr int kSearchParts = 8;  // threads that share the search around one image (consecutive lanes)                // bin edge >= kBinSlack * cutoff (rounding margin)

// slots of the small int array the host reads back
enum { kSizeAllAtom = 0, kSizeMaxNeigh = 1, kSizeSparse = 2, kSizeOrb = 3, kSizeCount = 4 };

struct Grid {
  double lo[3];    // lower corner of the bounding box of the central-cell atoms
  double invW[3];  // 1 / bin edge
  int nb[3];       // bins per dimension
  int pad;
};

// order-preserving map double -> unsigned 64 bit (for atomicMin / atomicMax on coordinates)
NL_DEV ull encodeOrdered(double x) {
  const ull u = NL_BITS(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
NL_DEV double decodeOrdered(ull e) {
  const ull u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  return NL_FROM_BITS(u);
}

NL_DEV int binCoord(double x, double lo, double invW, int nb) {
  const int i = (int)floor((x - lo) * invW);
  return i < 0 ? 0 : (i >= nb ? nb - 1 : i);
}

NL_DEV int binOf(const double *__restrict__ p, const Grid &g) {
  return binCoord(p[0], g.lo[0], g.invW[0], g.nb[0]) +
         g.nb[0] * (binCoord(p[1], g.lo[1], g.invW[1], g.nb[1]) +
                    g.nb[1] * binCoord(p[2], g.lo[2], g.invW[2], g.nb[2]));
}

// foldCoordsToCell, pbc3d without contacts (dftb/boundarycond.F90:369-380): inv = invert33(latVec),
// vecLen = column norms, all column-major; prepared on the host by foldSetup (n2_neighbour_driver.h).
struct FoldParams {
  double inv[9], lat[9], vecLen[3];
};

NL_KERNEL void nl_fold(int nAtom, FoldParams fp, double *__restrict__ coord) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    double *x = coord + 3 * a;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    double frac[3];
    for (int i = 0; i < 3; ++i) {
      // frac(:) = matmul(invLatVecs, coord(:,iAt)); frac = frac - floor(frac)
      double f = NL_ADD(NL_ADD(NL_MUL(fp.inv[i], x0), NL_MUL(fp.inv[i + 3], x1)), NL_MUL(fp.inv[i + 6], x2));
      f = NL_SUB(f, floor(f));
      // where (abs(vecLen * (1.0_dp - frac)) < epsilon(0.0_dp)) frac = 0.0_dp
      if (fabs(NL_MUL(fp.vecLen[i], NL_SUB(1.0, f))) < 2.220446049250313e-16) f = 0.0;
      frac[i] = f;
    }
    for (int i = 0; i < 3; ++i)  // coord(:, iAt) = matmul(latVec, frac)
      x[i] = NL_ADD(NL_ADD(NL_MUL(fp.lat[i], frac[0]), NL_MUL(fp.lat[i + 3], frac[1])), NL_MUL(fp.lat[i + 6], frac[2]));
  }
}

// box[0..2] = min, box[3..5] = max of coord0 (encoded); box is preset to all-ones / zero
NL_KERNEL void nl_bbox(int nAtom, const double *__restrict__ coord0, ull *__restrict__ box) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE)
    for (int k = 0; k < 3; ++k) {
      const ull e = encodeOrdered(coord0[3 * a + k]);
      atomicMin(box + k, e);
      atomicMax(box + 3 + k, e);
    }
}

// one thread: bins per dimension so that the bin edge is at least kBinSlack * cutoff
NL_KERNEL void nl_grid_setup(const ull *__restrict__ box, double cutoff, Grid *__restrict__ grid) {
  if (NL_TID != 0) return;
  const double w0 = cutoff > 0.0 ? cutoff * kBinSlack : 1.0;
  Grid g;
  for (int k = 0; k < 3; ++k) {
    const double lo = decodeOrdered(box[k]), ext = decodeOrdered(box[3 + k]) - lo;
    int n = (int)floor(ext / w0);
    n = n < 1 ? 1 : (n > kMaxBinsPerDim ? kMaxBinsPerDim : n);
    const double w = ext / n > w0 ? ext / n : w0;
    g.lo[k] = lo;
    g.invW[k] = 1.0 / w;
    g.nb[k] = n;
  }
  g.pad = 0;
  *grid = g;
}

NL_KERNEL void nl_bin_count(int nAtom, const double *__restrict__ coord0, const Grid *__restrict__ grid,
                            int *__restrict__ binCount) {
  const Grid g = *grid;
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) atomicAdd(binCount + binOf(coord0 + 3 * a, g), 1);
}

// atoms of a bin stored contiguously with their coordinates (order inside a bin: arbitrary)
NL_KERNEL void nl_bin_fill(int nAtom, const double *__restrict__ coord0, const Grid *__restrict__ grid,
                           const int *__restrict__ binStart, int *__restrict__ binCursor,
                           int *__restrict__ binAtom, double *__restrict__ binXyz) {
  const Grid g = *grid;
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    const int b = binOf(coord0 + 3 * a, g);
    const int pos = binStart[b] + atomicAdd(binCursor + b, 1);
    binAtom[pos] = (int)a;
    binXyz[3 * pos] = coord0[3 * a];
    binXyz[3 * pos + 1] = coord0[3 * a + 1];
    binXyz[3 * pos + 2] = coord0[3 * a + 2];
  }
}

// rr = coord0(:,a1) + rCellVec(:,ii) (periodic.F90:579) and its (unclamped) bin coordinates; false if
// the image is farther than one bin edge (>= cutoff) from the bounding box of the central cell, i.e.
// cannot have a neighbour there.
NL_DEV bool imageInReach(int ii, int a1, const double *__restrict__ coord0, const double *__restrict__ rCellVec,
                         const Grid &g, double *rr, int *ip) {
  for (int k = 0; k < 3; ++k) {
    rr[k] = NL_ADD(coord0[3 * a1 + k], rCellVec[3 * ii + k]);
    const double f = (rr[k] - g.lo[k]) * g.invW[k];
    if (!(f >= -1.0 && f < (double)(g.nb[k] + 1))) return false;
    ip[k] = (int)floor(f);
  }
  return true;
}

// Stream compaction of the candidate images: with a cell larger than the cutoff only the images in
// a shell of one bin edge around the central cell can interact (fec4096: 9.5 k of 110 k), and they are
// scattered over the candidate numbers -- looping over all candidates would leave one or two
// active lanes per warp in the search kernels.  The order of the list is arbitrary (atomic cursor);
// everything downstream is keyed by the candidate number itself.
NL_KERNEL void nl_active(long long nCand, int nAtom, const double *__restrict__ coord0,
                         const double *__restrict__ rCellVec, const Grid *__restrict__ grid,
                         int *__restrict__ nActive, int *__restrict__ activeList) {
  const Grid g = *grid;
  for (long long c = NL_TID; c < nCand; c += NL_STRIDE) {
    const int ii = (int)(c / nAtom), a1 = (int)(c - (long long)ii * nAtom);
    double rr[3];
    int ip[3];
    if (imageInReach(ii, a1, coord0, rCellVec, g, rr, ip)) activeList[atomicAdd(nActive, 1)] = (int)c;
  }
}

// Image of atom a1 (0-based) in cell ii (0-based): rr = coord0(:,a1) + rCellVec(:,ii)
// (periodic.F90:579), and every central atom a2 <= a1 with dist2 = sum((coord0(:,a2) - rr)**2) <=
// cutoff2 (:597-600) is handed to visit(a2, dist2) -- every central atom for the symmetric list
// (iAtom2End = nAtom).  The atoms of the 27 bins are shared by nParts threads (thread `part` takes
// every nParts-th atom of each bin row).  Returns false if the image is too far from the bounding box.
template <class F>
NL_DEV bool visitNeighbours(bool symmetric, int part, int nParts, int ii, int a1, const double *__restrict__ coord0,
                            const double *__restrict__ rCellVec, const Grid &g,
                            const int *__restrict__ binStart, const int *__restrict__ binAtom,
                            const double *__restrict__ binXyz, double cutoff2, double *rr, F &&visit) {
  int ip[3];
  if (!imageInReach(ii, a1, coord0, rCellVec, g, rr, ip)) return false;
  const int z0 = ip[2] - 1 < 0 ? 0 : ip[2] - 1, z1 = ip[2] + 1 > g.nb[2] - 1 ? g.nb[2] - 1 : ip[2] + 1;
  const int y0 = ip[1] - 1 < 0 ? 0 : ip[1] - 1, y1 = ip[1] + 1 > g.nb[1] - 1 ? g.nb[1] - 1 : ip[1] + 1;
  const int x0 = ip[0] - 1 < 0 ? 0 : ip[0] - 1, x1 = ip[0] + 1 > g.nb[0] - 1 ? g.nb[0] - 1 : ip[0] + 1;
  for (int iz = z0; iz <= z1; ++iz)
    for (int iy = y0; iy <= y1; ++iy) {
      // bins x0..x1 of one (iy, iz) row are contiguous
      const int row = g.nb[0] * (iy + g.nb[1] * iz);
      const int pEnd = binStart[row + x1 + 1];
      for (int p = binStart[row + x0] + part; p < pEnd; p += nParts) {  // this thread's share of the row
        const int a2 = binAtom[p];
        if (a2 > a1 && !symmetric) continue;  // non-symmetric list: iAtom2 = 1 .. iAtom1 (:580-584)
        const double d0 = NL_SUB(binXyz[3 * p], rr[0]);
        const double d1 = NL_SUB(binXyz[3 * p + 1], rr[1]);
        const double d2 = NL_SUB(binXyz[3 * p + 2], rr[2]);
        const double dist2 = NL_ADD(NL_ADD(NL_MUL(d0, d0), NL_MUL(d1, d1)), NL_MUL(d2, d2));
        if (dist2 > cutoff2) continue;
        visit(a2, dist2);
      }
    }
  return true;
}

// pass 1: which candidate images interact at all (flag), neighbours per atom, first too-close pair
// (flag is zero-filled by the caller: candidates that are not in the list stay 0)
NL_KERNEL void nl_count(const int *__restrict__ nActive, const int *__restrict__ activeList, int nAtom,
                        int symmetric, const double *__restrict__ coord0,
                        const double *__restrict__ rCellVec, const Grid *__restrict__ grid,
                        const int *__restrict__ binStart, const int *__restrict__ binAtom,
                        const double *__restrict__ binXyz, double cutoff2, int *__restrict__ flag,
                        int *__restrict__ nNeighbour, ull *__restrict__ firstError) {
  const Grid g = *grid;
  const long long nItems = (long long)*nActive * kSearchParts;
  for (long long i = NL_TID; i < nItems; i += NL_STRIDE) {
    const long long c = activeList[i / kSearchParts];
    const int part = (int)(i % kSearchParts);
    const int ii = (int)(c / nAtom), a1 = (int)(c - (long long)ii * nAtom);
    double rr[3];
    int found = 0;
    visitNeighbours(symmetric != 0, part, kSearchParts, ii, a1, coord0, rCellVec, g, binStart, binAtom, binXyz, cutoff2, rr,
                    [&](int a2, double dist2) {
                      found = 1;  // the image is appended before the proximity check (:604-614)
                      if (dist2 < kMinNeighDist2) {
                        if (!(ii == 0 && a1 == a2)) atomicMin(firstError, (ull)c * (ull)nAtom + (ull)a2);
                        return;  // the atom itself in the central cell is not its own neighbour (:619-622)
                      }
                      atomicAdd(nNeighbour + a2, 1);
                    });
    if (found) flag[c] = 1;  // the threads of an image store the same value
  }
}

// sizes[kSizeAllAtom] = number of flagged candidates, sizes[kSizeMaxNeigh] = max nNeighbour
NL_KERNEL void nl_sizes(long long nCand, int nAtom, const int *__restrict__ imgIndex,
                        const int *__restrict__ nNeighbour, int *__restrict__ sizes) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    atomicMax(sizes + kSizeMaxNeigh, nNeighbour[a]);
    if (a == 0) sizes[kSizeAllAtom] = imgIndex[nCand];
  }
}

// pass 2: image arrays (coord, img2CentCell, iCellVec, species; periodic.F90:610-613, :437) and the
// unsorted neighbour rows (slots in arrival order; nl_rank / nl_sort put them into the reference's order)
NL_KERNEL void nl_fill(const int *__restrict__ nActive, const int *__restrict__ activeList, int nAtom,
                       int symmetric, int ld, const double *__restrict__ coord0,
                       const int *__restrict__ species0, const double *__restrict__ rCellVec,
                       const Grid *__restrict__ grid, const int *__restrict__ binStart,
                       const int *__restrict__ binAtom, const double *__restrict__ binXyz, double cutoff2,
                       const int *__restrict__ flag, const int *__restrict__ imgIndex,
                       double *__restrict__ coords, int *__restrict__ species, int *__restrict__ img2CentCell,
                       int *__restrict__ iCellVec, int *__restrict__ cursor, int *__restrict__ iNeighbour,
                       double *__restrict__ neighDist2) {
  const Grid g = *grid;
  const long long nItems = (long long)*nActive * kSearchParts;
  for (long long i = NL_TID; i < nItems; i += NL_STRIDE) {
    const long long c = activeList[i / kSearchParts];
    const int part = (int)(i % kSearchParts);
    if (!flag[c]) continue;
    const int ii = (int)(c / nAtom), a1 = (int)(c - (long long)ii * nAtom);
    const int img = imgIndex[c];
    double rr[3];
    visitNeighbours(symmetric != 0, part, kSearchParts, ii, a1, coord0, rCellVec, g, binStart, binAtom, binXyz, cutoff2, rr,
                    [&](int a2, double dist2) {
                      if (dist2 < kMinNeighDist2) return;
                      const int slot = atomicAdd(cursor + a2, 1) + 1;
                      iNeighbour[slot + (long long)ld * a2] = img + 1;
                      neighDist2[slot + (long long)ld * a2] = dist2;
                    });
    if (part != 0) continue;  // one thread of the image writes its entry of the image arrays
    coords[3 * (long long)img] = rr[0];
    coords[3 * (long long)img + 1] = rr[1];
    coords[3 * (long long)img + 2] = rr[2];
    img2CentCell[img] = a1 + 1;
    iCellVec[img] = ii + 1;
    species[img] = species0[a1];
  }
}

// index_heap_sort (math/sorting.F90:142-199) acting on the records (key[i], val[i]), i = 1..n,
// instead of on an index vector: the algorithm only compares values and moves indices, so moving
// the records gives the same order.  less(kA, vA, kB, vB) is the reference's comparison.
template <class Less>
NL_DEV void heapSortRecords(int n, int *__restrict__ key, double *__restrict__ val, Less less) {
  if (n <= 1) return;
  int il = n / 2 + 1, ir = n;
  for (;;) {
    int kTmp;
    double vTmp;
    if (il > 1) {
      il -= 1;
      kTmp = key[il];
      vTmp = val[il];
    } else {
      kTmp = key[ir];
      vTmp = val[ir];
      key[ir] = key[1];
      val[ir] = val[1];
      ir -= 1;
      if (ir < 1) {
        key[1] = kTmp;
        val[1] = vTmp;
        return;
      }
    }
    int ii = il, ij = 2 * il;
    while (ij <= ir) {
      if (ij < ir && less(key[ij], val[ij], key[ij + 1], val[ij + 1])) ij += 1;
      if (less(kTmp, vTmp, key[ij], val[ij])) {
        key[ii] = key[ij];
        val[ii] = val[ij];
        ii = ij;
        ij = 2 * ij;
      } else {
        ij = ir + 1;
      }
    }
    key[ii] = kTmp;
    val[ii] = vTmp;
  }
}

// search_asc_real_geq (math/binarysearch.F90:63-94): last jj in 1..n with xVal - xx(jj) >= -tol, 0 if
// none; xx1[jj] = xx(jj) (the element before xx(1) is never touched).
NL_DEV int searchAscRealGeq(int n, const double *__restrict__ xx1, double xVal, double tol) {
  int jlower = 0, jupper = n;
  while (jlower < jupper) {
    const int jcurr = jlower + (jupper - jlower + 1) / 2;
    if (NL_SUB(xVal, xx1[jcurr]) >= -tol) jlower = jcurr;
    else jupper = jcurr - 1;
  }
  return jlower;
}

// One thread per row entry: position of the entry (a) in ascending image number -- the order in which
// the reference's loops append, i.e. the INPUT order of its sort -- and (b) in ascending distance.
// If no two distances of a row are within tolSameDist2 of each other, the reference's comparison
// a < b - tol coincides with a < b on that row (a < b excludes b < a - tol, so a < b - tol must hold),
// which is a strict total order: every correct comparison sort, the reference's heap sort included,
// returns the one ascending sequence, and the rank by distance IS the final position.  Rows with a
// near-tie (perfect lattices) are flagged and go through the reference's algorithm in nl_sort; their
// by-distance slots written here are overwritten there.
//   arrive*:  rows in arrival order (nl_fill)         byKey*: rows in ascending image number
//   iNeighbour / neighDist2: rows in ascending distance (final for rows without ties)
// Rows longer than rankMax skip the counting (quadratic in the row length) and take the n log n path of
// nl_sort from the arrival order (flag 2).
NL_KERNEL void nl_rank(int nAtom, int ld, int rankMax, const int *__restrict__ nNeighbour, const int *__restrict__ arriveKey,
                       const double *__restrict__ arriveDist, int *__restrict__ byKeyKey,
                       double *__restrict__ byKeyDist, int *__restrict__ iNeighbour,
                       double *__restrict__ neighDist2, int *__restrict__ rowHasTies) {
  const long long total = (long long)ld * nAtom;
  for (long long e = NL_TID; e < total; e += NL_STRIDE) {
    const int a = (int)(e / ld), k = (int)(e - (long long)a * ld);
    const int n = nNeighbour[a];
    if (k == 0 || k > n) continue;
    if (n > rankMax) {
      if (k == 1) rowHasTies[a] = 2;
      continue;
    }
    const int *__restrict__ keys = arriveKey + (long long)ld * a;
    const double *__restrict__ dist = arriveDist + (long long)ld * a;
    const int myKey = keys[k];
    const double myDist = dist[k];
    const double myDistMinusTol = NL_SUB(myDist, kTolSameDist2);
    int rankKey = 0, rankDist = 0;
    bool tie = false;
    for (int j = 1; j <= n; ++j) {
      if (j == k) continue;
      if (keys[j] < myKey) ++rankKey;
      const double dj = dist[j];
      if (dj < myDistMinusTol) ++rankDist;                          // dj  <  myDist - tol
      else if (!(myDist < NL_SUB(dj, kTolSameDist2))) tie = true;  // neither is smaller by more than tol
    }
    const long long row = (long long)ld * a;
    byKeyKey[row + 1 + rankKey] = myKey;
    byKeyDist[row + 1 + rankKey] = myDist;
    iNeighbour[row + 1 + rankDist] = myKey;
    neighDist2[row + 1 + rankDist] = myDist;
    if (tie) rowHasTies[a] = 1;  // every writer stores the same value
  }
}

// One thread per atom: rows with near-ties through the reference's heap sort (input = the row in
// ascending image number), row 0 = the atom itself (fillNeighbourArrays), nNeighbourSK for the SK
// cutoff, and the number of sparse elements / orbitals of the atom.
NL_KERNEL void nl_sort(int nAtom, int ld, double skCutoff2, const int *__restrict__ nNeighbour,
                       const int *__restrict__ species0, const int *__restrict__ nOrbSpecies,
                       const int *__restrict__ img2CentCell, const int *__restrict__ rowHasTies,
                       const int *__restrict__ arriveKey, const double *__restrict__ arriveDist,
                       int *__restrict__ byKeyKey, double *__restrict__ byKeyDist, int *__restrict__ iNeighbour,
                       double *__restrict__ neighDist2, int *__restrict__ nNeighbourSK,
                       int *__restrict__ atomSize, int *__restrict__ orbCount) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    int *in = iNeighbour + (long long)ld * a;      // in[1..n]
    double *dn = neighDist2 + (long long)ld * a;
    const int n = nNeighbour[a];
    if (rowHasTies[a]) {
      int *kk = byKeyKey + (long long)ld * a;
      double *kd = byKeyDist + (long long)ld * a;
      if (rowHasTies[a] == 2) {  // long row: not ranked; ascending image number by a plain heap sort
        for (int k = 1; k <= n; ++k) {
          kk[k] = arriveKey[(long long)ld * a + k];
          kd[k] = arriveDist[(long long)ld * a + k];
        }
        heapSortRecords(n, kk, kd, [](int kA, double, int kB, double) { return kA < kB; });
      }
      // index_heap_sort(indx, neighDist2(1:nn1, iAtom1), tolSameDist2), periodic.F90:677-680
      heapSortRecords(n, kk, kd, [](int, double vA, int, double vB) { return vA < NL_SUB(vB, kTolSameDist2); });
      for (int k = 1; k <= n; ++k) {
        in[k] = kk[k];
        dn[k] = kd[k];
      }
    }
    in[0] = (int)a + 1;
    dn[0] = 0.0;
    // getNrOfNeighbours: search_asc_real_geq(nNeighbour, neighDist2(1:n), cutoff**2, tolSameDist2)
    const int jlower = searchAscRealGeq(n, dn, skCutoff2, kTolSameDist2);
    nNeighbourSK[a] = jlower;
    const int nOrb1 = nOrbSpecies[species0[a] - 1];
    int sum = 0;
    for (int k = 0; k <= jlower; ++k) sum += nOrb1 * nOrbSpecies[species0[img2CentCell[in[k] - 1] - 1] - 1];
    atomSize[a] = sum;
    orbCount[a] = nOrb1;
  }
}

// getNrOfNeighboursForAll (periodic.F90:996-1020) for another cutoff on the finished list
NL_KERNEL void nl_nr_of_neighbours(int nAtom, int ld, double cutoff2, const int *__restrict__ nNeighbour,
                                   const double *__restrict__ neighDist2, int *__restrict__ out) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE)
    out[a] = searchAscRealGeq(nNeighbour[a], neighDist2 + (long long)ld * a, cutoff2, kTolSameDist2);
}

// getSparseDescriptor (sparse2dense.F90:3651-3660) from the per-atom offsets, and iAtomStart
NL_KERNEL void nl_ipair(int nAtom, int ld, const int *__restrict__ nNeighbourSK,
                        const int *__restrict__ species0, const int *__restrict__ nOrbSpecies,
                        const int *__restrict__ img2CentCell, const int *__restrict__ iNeighbour,
                        const int *__restrict__ atomOffset, const int *__restrict__ orbOffset,
                        int *__restrict__ iPair, int *__restrict__ iAtomStart, int *__restrict__ sizes) {
  for (long long a = NL_TID; a < nAtom; a += NL_STRIDE) {
    const int *in = iNeighbour + (long long)ld * a;
    int *ip = iPair + (long long)ld * a;
    const int nOrb1 = nOrbSpecies[species0[a] - 1];
    int ind = atomOffset[a];
    for (int k = 0; k <= nNeighbourSK[a]; ++k) {
      ip[k] = ind;
      ind += nOrb1 * nOrbSpecies[species0[img2CentCell[in[k] - 1] - 1] - 1];
    }
    iAtomStart[a] = orbOffset[a] + 1;
    if (a == nAtom - 1) {
      iAtomStart[nAtom] = orbOffset[nAtom] + 1;
      sizes[kSizeSparse] = atomOffset[nAtom];
      sizes[kSizeOrb] = orbOffset[nAtom];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Work list and block descriptors of the adopted geometry, built on the device (what set_geometry
// builds on the host from a read-back of the index arrays): one entry per sparse block (iNeigh =
// 0 .. nNeighbourSK(iAtom), iAtom) in the order of the sparse arrays.
// ------------------------------------------------------------------------------------------------

// layout of BlockDesc (block_desc.cuh; the .cu that launches these kernels asserts the match)
struct GeoDesc {
  int ind, ii, jj, n12, iAt1, iAt2f, iVec, iNeigh;
};

// cnt[a] = blocks of atom a (on-site block + nNeighbourSK), cnt[nAtom] = 0: its exclusive scan is the
// first work-list entry of every atom and, at [nAtom], the number of blocks
NL_KERNEL void geo_block_count(int nAtom, const int *__restrict__ nNeighbourSK, int *__restrict__ cnt) {
  for (long long a = NL_TID; a <= nAtom; a += NL_STRIDE) cnt[a] = a < nAtom ? nNeighbourSK[a] + 1 : 0;
}

// one thread per entry (iNeigh, iAtom) of the neighbour rows.  hasDuplicates: two blocks of one atom
// fold onto the same dense block (an image of the atom itself, or two images of one neighbour)
NL_KERNEL void geo_describe(int nAtom, int ld, const int *__restrict__ nNeighbourSK,
                            const int *__restrict__ blockStart, const int *__restrict__ iNeighbour,
                            const int *__restrict__ img2CentCell, const int *__restrict__ iCellVec,
                            const int *__restrict__ iPair, const int *__restrict__ iAtomStart,
                            int *__restrict__ workAll, GeoDesc *__restrict__ desc, int *__restrict__ hasDuplicates) {
  const long long n = (long long)ld * nAtom;
  for (long long t = NL_TID; t < n; t += NL_STRIDE) {
    const int i = (int)(t / ld), k = (int)(t - (long long)i * ld);
    if (k > nNeighbourSK[i]) continue;
    const int w = blockStart[i] + k;
    const int a2 = iNeighbour[t] - 1, f = img2CentCell[a2] - 1;
    workAll[w] = (int)t;
    GeoDesc d;
    d.ind = iPair[t];
    d.ii = iAtomStart[i] - 1;
    d.jj = iAtomStart[f] - 1;
    d.n12 = (iAtomStart[i + 1] - iAtomStart[i]) | ((iAtomStart[f + 1] - iAtomStart[f]) << 8);
    d.iAt1 = i;
    d.iAt2f = f;
    d.iVec = iCellVec[a2] - 1;
    d.iNeigh = k;
    desc[w] = d;
    if (k >= 1) {
      bool dup = f == i;
      const int *row = iNeighbour + (long long)i * ld;
      for (int k2 = 1; k2 < k && !dup; ++k2) dup = img2CentCell[row[k2] - 1] - 1 == f;
      if (dup) atomicMax(hasDuplicates, 1);
    }
  }
}

// groupFirst[g] = the last block that starts at or before element groupSize * g (flat_stream.cuh); the
// blocks are in ascending offset order by construction.  groupFirst is preset to zero.
NL_KERNEL void geo_group_index(int nBlocks, int nGroups, int groupSize, const GeoDesc *__restrict__ desc,
                               int *__restrict__ groupFirst) {
  for (long long w = NL_TID; w < nBlocks; w += NL_STRIDE) {
    const long long g0 = ((long long)desc[w].ind + groupSize - 1) / groupSize;
    const long long g1 = w + 1 < nBlocks ? ((long long)desc[w + 1].ind + groupSize - 1) / groupSize : nGroups;
    for (long long g = g0; g < g1 && g < nGroups; ++g) groupFirst[g] = (int)w;
  }
}

}  // namespace nl
}  // namespace dftbp_b200

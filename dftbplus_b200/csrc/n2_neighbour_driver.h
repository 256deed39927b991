// n2_neighbour_driver.h -- launch sequence of the device neighbour list (n2_neighbour_kernels.cuh),
// written against a small backend interface so that the same sequence runs on the GPU
// (n2_neighbour_list.cu: CUDA launches, cub::DeviceScan) and, for the CPU test suite, on the
// emulation backend of tests/host_eval/host_nl.cpp.
//
// Backend concept:
//   int  alloc(void **p, size_t bytes);            void release(void *p);
//   int  upload(void *dst, const void *src, size_t bytes);     // src: host or device memory
//   int  download(void *dst, const void *src, size_t bytes);   // blocks until the data is there
//   int  fill(void *p, int byteValue, size_t bytes);
//   int  scan(const int *in, int *out, long long n);           // exclusive prefix sum
//   int  launch(kernel, long long nItems, args...);            // grid-stride kernel over nItems
// All return 0 on success.
#pragma once
#include <cmath>
#include <cstddef>

#include "n2_neighbour_kernels.cuh"

namespace dftbp_b200 {
namespace nl {

template <class T>
struct Buf {
  T *p = nullptr;
  size_t cap = 0;
  template <class Backend>
  int ensure(Backend &be, size_t n) {
    if (n <= cap) return 0;
    if (p) be.release(p);
    p = nullptr;
    cap = 0;
    void *q = nullptr;
    if (int rc = be.alloc(&q, n * sizeof(T))) return rc;
    p = static_cast<T *>(q);
    cap = n;
    return 0;
  }
  template <class Backend>
  void free(Backend &be) {
    if (p) be.release(p);
    p = nullptr;
    cap = 0;
  }
};

#define NL_TRY(call)          \
  do {                        \
    int rc_ = (call);         \
    if (rc_ != 0) return rc_; \
  } while (0)

// rows up to this length are ordered by counting (n^2 per row, fully parallel), longer ones by heap sorts
constexpr int kRankMaxDefault = 192;

// Result of one update: sizes, and the device arrays of the reference's neighbour list / sparse
// descriptor in the reference's layout (1-based values, iPair 0-based, ld = mNeigh + 1).
struct NeighbourList {
  int nAtom = 0, nCellVec = 0, nAllAtom = 0, mNeigh = 0, sparseSize = 0, nOrb = 0;
  long long nCand = 0;
  double cutoff = 0.0, skCutoff = 0.0;
  bool valid = false;
  bool symmetric = false;  // symmetric list: not a geometry for the pair kernels (they expect every pair once)
  int errAtom1 = 0, errAtom2 = 0;  // "Atoms i and j too close together" (periodic.F90:645-648); 0 = none

  // inputs kept on the device
  Buf<double> coord0, rCellVec, cellVec;
  Buf<int> species0;
  // uniform grid over the central cell
  Buf<ull> box;  // [0..5] bounding box, [6] first error key
  Buf<Grid> grid;
  Buf<int> binCount, binStart, binAtom;
  Buf<double> binXyz;
  // candidate images
  Buf<int> flag, imgIndex, activeList, nActive;
  // per atom
  Buf<int> nNeighbour, cursor, nNeighbourSK, atomSize, atomOffset, orbCount, orbOffset, iAtomStart;
  Buf<int> sizes;
  // the lists
  Buf<double> coords, neighDist2;
  Buf<int> species, img2CentCell, iCellVec, iNeighbour, iPair;
  // scratch rows of the ordering step: arrival order, ascending image number; rows with near-ties
  Buf<int> arriveKey, byKeyKey, rowHasTies;
  Buf<double> arriveDist, byKeyDist;

  template <class Backend>
  void free(Backend &be) {
    coord0.free(be); rCellVec.free(be); cellVec.free(be); species0.free(be); box.free(be); grid.free(be);
    binCount.free(be); binStart.free(be); binAtom.free(be); binXyz.free(be); flag.free(be);
    imgIndex.free(be); activeList.free(be); nActive.free(be); nNeighbour.free(be); cursor.free(be); nNeighbourSK.free(be); atomSize.free(be);
    atomOffset.free(be); orbCount.free(be); orbOffset.free(be); iAtomStart.free(be); sizes.free(be);
    coords.free(be); neighDist2.free(be); species.free(be); img2CentCell.free(be); iCellVec.free(be);
    iNeighbour.free(be); iPair.free(be); arriveKey.free(be); byKeyKey.free(be); rowHasTies.free(be);
    arriveDist.free(be); byKeyDist.free(be);
    valid = false;
  }
};

// invert33 (invertInOut with determinant33, math/simplealgebra.F90:45-66 and the cofactor form after
// it) and the column norms of latVec(3,3) column-major, for nl_fold.  Host arithmetic in the
// reference's operation order.  Returns false for a singular cell.
inline bool foldSetup(const double *latVec, FoldParams &fp) {
  auto M = [&](int i, int j) { return latVec[(i - 1) + 3 * (j - 1)]; };
  double tmp = M(1, 1) * (M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3));
  tmp = tmp - M(1, 2) * (M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3));
  tmp = tmp + M(1, 3) * (M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2));
  const double det = tmp;
  if (!(det != 0.0)) return false;
  double *inv = fp.inv;
  inv[0] = -M(2, 3) * M(3, 2) + M(2, 2) * M(3, 3);
  inv[1] = M(2, 3) * M(3, 1) - M(2, 1) * M(3, 3);
  inv[2] = -M(2, 2) * M(3, 1) + M(2, 1) * M(3, 2);
  inv[3] = M(1, 3) * M(3, 2) - M(1, 2) * M(3, 3);
  inv[4] = -M(1, 3) * M(3, 1) + M(1, 1) * M(3, 3);
  inv[5] = M(1, 2) * M(3, 1) - M(1, 1) * M(3, 2);
  inv[6] = -M(1, 3) * M(2, 2) + M(1, 2) * M(2, 3);
  inv[7] = M(1, 3) * M(2, 1) - M(1, 1) * M(2, 3);
  inv[8] = -M(1, 2) * M(2, 1) + M(1, 1) * M(2, 2);
  for (int k = 0; k < 9; ++k) {
    inv[k] = inv[k] / det;
    fp.lat[k] = latVec[k];
  }
  for (int j = 1; j <= 3; ++j) fp.vecLen[j - 1] = std::sqrt(M(1, j) * M(1, j) + M(2, j) * M(2, j) + M(3, j) * M(3, j));
  return true;
}

// updateNeighbourListAndSpecies + getNrOfNeighboursForAll(skCutoff) + getSparseDescriptor
// (handleCoordinateChange, dftbplus/main.F90:2562-2578).  coord0(3,nAtom), species0(nAtom),
// cellVec / rCellVec(3,nCellVec) with the zero vector first, nOrbSpecies on the device.
// symmetric: the symmetric list (periodic.F90:580-584).
// Returns 0, or the backend's error code; a too-close pair is reported in nb.errAtom1/2 with
// nb.valid = false and return value 0.
template <class Backend>
int updateNeighbourList(Backend &be, NeighbourList &nb, int nAtom, const double *coord0, const int *species0,
                        const int *dNOrbSpecies, int nCellVec, const double *cellVec, const double *rCellVec,
                        double cutoff, double skCutoff, bool symmetric = false, int rankMax = kRankMaxDefault) {
  nb.valid = false;
  nb.errAtom1 = nb.errAtom2 = 0;
  nb.nAtom = nAtom;
  nb.nCellVec = nCellVec;
  nb.cutoff = cutoff;
  nb.skCutoff = skCutoff;
  nb.symmetric = symmetric;
  const long long nCand = (long long)nCellVec * nAtom;
  nb.nCand = nCand;
  const double cutoff2 = cutoff * cutoff;  // cutoff2 = cutoff**2, periodic.F90:536

  NL_TRY(nb.coord0.ensure(be, (size_t)3 * nAtom));
  NL_TRY(nb.species0.ensure(be, nAtom));
  NL_TRY(nb.rCellVec.ensure(be, (size_t)3 * nCellVec));
  NL_TRY(nb.cellVec.ensure(be, (size_t)3 * nCellVec));
  NL_TRY(be.upload(nb.coord0.p, coord0, sizeof(double) * 3 * nAtom));
  NL_TRY(be.upload(nb.species0.p, species0, sizeof(int) * nAtom));
  NL_TRY(be.upload(nb.rCellVec.p, rCellVec, sizeof(double) * 3 * nCellVec));
  NL_TRY(be.upload(nb.cellVec.p, cellVec, sizeof(double) * 3 * nCellVec));

  // grid over the central cell
  NL_TRY(nb.box.ensure(be, 8));
  NL_TRY(nb.grid.ensure(be, 1));
  NL_TRY(nb.binCount.ensure(be, (size_t)kMaxBins + 1));
  NL_TRY(nb.binStart.ensure(be, (size_t)kMaxBins + 1));
  NL_TRY(nb.binAtom.ensure(be, nAtom));
  NL_TRY(nb.binXyz.ensure(be, (size_t)3 * nAtom));
  NL_TRY(be.fill(nb.box.p, 0xff, 3 * sizeof(ull)));      // minima
  NL_TRY(be.fill(nb.box.p + 3, 0x00, 3 * sizeof(ull)));  // maxima
  NL_TRY(be.fill(nb.box.p + 6, 0xff, sizeof(ull)));      // first error key
  NL_TRY(be.launch(nl_bbox, nAtom, nAtom, (const double *)nb.coord0.p, nb.box.p));
  NL_TRY(be.launch(nl_grid_setup, 1, (const ull *)nb.box.p, cutoff, nb.grid.p));
  NL_TRY(be.fill(nb.binCount.p, 0, sizeof(int) * ((size_t)kMaxBins + 1)));
  NL_TRY(be.launch(nl_bin_count, nAtom, nAtom, (const double *)nb.coord0.p, (const Grid *)nb.grid.p,
                   nb.binCount.p));
  NL_TRY(be.scan(nb.binCount.p, nb.binStart.p, (long long)kMaxBins + 1));
  NL_TRY(be.fill(nb.binCount.p, 0, sizeof(int) * ((size_t)kMaxBins + 1)));  // reused as the fill cursors
  NL_TRY(be.launch(nl_bin_fill, nAtom, nAtom, (const double *)nb.coord0.p, (const Grid *)nb.grid.p,
                   (const int *)nb.binStart.p, nb.binCount.p, nb.binAtom.p, nb.binXyz.p));

  // pass 1 over the candidate images
  NL_TRY(nb.flag.ensure(be, (size_t)nCand + 1));
  NL_TRY(nb.imgIndex.ensure(be, (size_t)nCand + 1));
  NL_TRY(nb.nNeighbour.ensure(be, nAtom));
  NL_TRY(nb.cursor.ensure(be, nAtom));
  NL_TRY(nb.sizes.ensure(be, kSizeCount));
  NL_TRY(nb.activeList.ensure(be, (size_t)nCand));
  NL_TRY(nb.nActive.ensure(be, 1));
  NL_TRY(be.fill(nb.nActive.p, 0, sizeof(int)));
  NL_TRY(be.launch(nl_active, nCand, nCand, nAtom, (const double *)nb.coord0.p, (const double *)nb.rCellVec.p,
                   (const Grid *)nb.grid.p, nb.nActive.p, nb.activeList.p));
  NL_TRY(be.fill(nb.flag.p, 0, sizeof(int) * ((size_t)nCand + 1)));
  NL_TRY(be.fill(nb.nNeighbour.p, 0, sizeof(int) * nAtom));
  NL_TRY(be.fill(nb.sizes.p, 0, sizeof(int) * kSizeCount));
  NL_TRY(be.launch(nl_count, nCand * kSearchParts, (const int *)nb.nActive.p, (const int *)nb.activeList.p, nAtom,
                   symmetric ? 1 : 0, (const double *)nb.coord0.p, (const double *)nb.rCellVec.p,
                   (const Grid *)nb.grid.p, (const int *)nb.binStart.p, (const int *)nb.binAtom.p,
                   (const double *)nb.binXyz.p, cutoff2, nb.flag.p, nb.nNeighbour.p, nb.box.p + 6));
  NL_TRY(be.scan(nb.flag.p, nb.imgIndex.p, nCand + 1));
  NL_TRY(be.launch(nl_sizes, nAtom, nCand, nAtom, (const int *)nb.imgIndex.p, (const int *)nb.nNeighbour.p,
                   nb.sizes.p));
  int hSizes[kSizeCount];
  ull hErr = 0;
  NL_TRY(be.download(hSizes, nb.sizes.p, sizeof(hSizes)));
  NL_TRY(be.download(&hErr, nb.box.p + 6, sizeof(hErr)));
  if (hErr != ~0ull) {
    const ull c = hErr / (ull)nAtom;
    nb.errAtom1 = (int)(c % (ull)nAtom) + 1;
    nb.errAtom2 = (int)(hErr % (ull)nAtom) + 1;
    return 0;
  }
  nb.nAllAtom = hSizes[kSizeAllAtom];
  nb.mNeigh = hSizes[kSizeMaxNeigh];
  const int ld = nb.mNeigh + 1;

  // pass 2: image arrays and unsorted rows
  NL_TRY(nb.coords.ensure(be, (size_t)3 * nb.nAllAtom));
  NL_TRY(nb.species.ensure(be, nb.nAllAtom));
  NL_TRY(nb.img2CentCell.ensure(be, nb.nAllAtom));
  NL_TRY(nb.iCellVec.ensure(be, nb.nAllAtom));
  NL_TRY(nb.iNeighbour.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.neighDist2.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.iPair.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.arriveKey.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.arriveDist.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.byKeyKey.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.byKeyDist.ensure(be, (size_t)ld * nAtom));
  NL_TRY(nb.rowHasTies.ensure(be, nAtom));
  NL_TRY(nb.nNeighbourSK.ensure(be, nAtom));
  NL_TRY(nb.atomSize.ensure(be, (size_t)nAtom + 1));
  NL_TRY(nb.atomOffset.ensure(be, (size_t)nAtom + 1));
  NL_TRY(nb.orbCount.ensure(be, (size_t)nAtom + 1));
  NL_TRY(nb.orbOffset.ensure(be, (size_t)nAtom + 1));
  NL_TRY(nb.iAtomStart.ensure(be, (size_t)nAtom + 1));
  NL_TRY(be.fill(nb.cursor.p, 0, sizeof(int) * nAtom));
  NL_TRY(be.fill(nb.iNeighbour.p, 0, sizeof(int) * ld * nAtom));    // iNeighbour(:,:) = 0, :565
  NL_TRY(be.fill(nb.neighDist2.p, 0, sizeof(double) * ld * nAtom));  // neighDist2(:,:) = 0.0_dp, :566
  NL_TRY(be.fill(nb.iPair.p, 0, sizeof(int) * ld * nAtom));          // iPair(:,:) = 0, sparse2dense.F90:3648
  NL_TRY(be.launch(nl_fill, nCand * kSearchParts, (const int *)nb.nActive.p, (const int *)nb.activeList.p, nAtom,
                   symmetric ? 1 : 0, ld, (const double *)nb.coord0.p, (const int *)nb.species0.p,
                   (const double *)nb.rCellVec.p, (const Grid *)nb.grid.p, (const int *)nb.binStart.p,
                   (const int *)nb.binAtom.p, (const double *)nb.binXyz.p, cutoff2, (const int *)nb.flag.p,
                   (const int *)nb.imgIndex.p, nb.coords.p, nb.species.p, nb.img2CentCell.p, nb.iCellVec.p,
                   nb.cursor.p, nb.arriveKey.p, nb.arriveDist.p));

  // order of the rows, nNeighbourSK, sparse descriptor
  NL_TRY(be.fill(nb.atomSize.p + nAtom, 0, sizeof(int)));
  NL_TRY(be.fill(nb.orbCount.p + nAtom, 0, sizeof(int)));
  NL_TRY(be.fill(nb.rowHasTies.p, 0, sizeof(int) * nAtom));
  NL_TRY(be.launch(nl_rank, (long long)ld * nAtom, nAtom, ld, rankMax, (const int *)nb.nNeighbour.p,
                   (const int *)nb.arriveKey.p, (const double *)nb.arriveDist.p, nb.byKeyKey.p, nb.byKeyDist.p,
                   nb.iNeighbour.p, nb.neighDist2.p, nb.rowHasTies.p));
  NL_TRY(be.launch(nl_sort, nAtom, nAtom, ld, skCutoff * skCutoff, (const int *)nb.nNeighbour.p,
                   (const int *)nb.species0.p, dNOrbSpecies, (const int *)nb.img2CentCell.p,
                   (const int *)nb.rowHasTies.p, (const int *)nb.arriveKey.p, (const double *)nb.arriveDist.p,
                   nb.byKeyKey.p, nb.byKeyDist.p, nb.iNeighbour.p,
                   nb.neighDist2.p, nb.nNeighbourSK.p, nb.atomSize.p, nb.orbCount.p));
  NL_TRY(be.scan(nb.atomSize.p, nb.atomOffset.p, (long long)nAtom + 1));
  NL_TRY(be.scan(nb.orbCount.p, nb.orbOffset.p, (long long)nAtom + 1));
  NL_TRY(be.launch(nl_ipair, nAtom, nAtom, ld, (const int *)nb.nNeighbourSK.p, (const int *)nb.species0.p,
                   dNOrbSpecies, (const int *)nb.img2CentCell.p, (const int *)nb.iNeighbour.p,
                   (const int *)nb.atomOffset.p, (const int *)nb.orbOffset.p, nb.iPair.p, nb.iAtomStart.p,
                   nb.sizes.p));
  NL_TRY(be.download(hSizes, nb.sizes.p, sizeof(hSizes)));
  nb.sparseSize = hSizes[kSizeSparse];
  nb.nOrb = hSizes[kSizeOrb];
  nb.valid = true;
  return 0;
}


// Work list / block descriptors of a list for adopt_neighbour_list, in two steps because the caller
// sizes the arrays of the second from the result of the first:
//   countBlocks:     blockStart(nAtom + 1) on the device and, read back, on the host (hBlockStart)
//   describeBlocks:  workAll(nBlocks), desc(nBlocks), groupFirst(nGroups), flags[0] = hasDuplicates
template <class Backend>
int countBlocks(Backend &be, const NeighbourList &L, int *cnt, int *blockStart, int *hBlockStart) {
  NL_TRY(be.launch(geo_block_count, (long long)L.nAtom + 1, L.nAtom, (const int *)L.nNeighbourSK.p, cnt));
  NL_TRY(be.scan(cnt, blockStart, (long long)L.nAtom + 1));
  NL_TRY(be.download(hBlockStart, blockStart, sizeof(int) * ((size_t)L.nAtom + 1)));
  return 0;
}

template <class Backend>
int describeBlocks(Backend &be, const NeighbourList &L, const int *blockStart, int nBlocks, int groupSize,
                   int *workAll, GeoDesc *desc, int *groupFirst, int *flags, int *hFlags) {
  const int ld = L.mNeigh + 1;
  const int nGroups = (int)(((long long)L.sparseSize + groupSize - 1) / groupSize);
  NL_TRY(be.fill(flags, 0, sizeof(int)));
  NL_TRY(be.fill(groupFirst, 0, sizeof(int) * (size_t)nGroups));
  NL_TRY(be.launch(geo_describe, (long long)ld * L.nAtom, L.nAtom, ld, (const int *)L.nNeighbourSK.p, blockStart,
                   (const int *)L.iNeighbour.p, (const int *)L.img2CentCell.p, (const int *)L.iCellVec.p,
                   (const int *)L.iPair.p, (const int *)L.iAtomStart.p, workAll, desc, flags));
  NL_TRY(be.launch(geo_group_index, nBlocks, nBlocks, nGroups, groupSize, (const GeoDesc *)desc, groupFirst));
  NL_TRY(be.download(hFlags, flags, sizeof(int)));
  return 0;
}

}  // namespace nl
}  // namespace dftbp_b200

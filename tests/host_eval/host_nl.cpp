// host_nl.cpp -- runs the PRODUCT's neighbour-list kernels and launch sequence
// (dftbplus_b200/csrc/n2_neighbour_kernels.cuh + n2_neighbour_driver.h) on the CPU, so that the CPU
// test suite can check them against the oracle without a GPU.  Test infrastructure.
//
// The kernels are grid-stride loops that communicate through atomics only, so executing the
// emulated threads one after the other is one of the interleavings the GPU may produce; the
// launcher runs them in a shuffled order (seeded) to show that the results do not depend on it.
// What this cannot cover: the CUDA backend itself (cub::DeviceScan, copies, launch shapes) -- the
// `-m gpu` parity tests do.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

typedef unsigned long long ull_t;

static long long g_tid = 0, g_stride = 1;
#define NL_TID g_tid
#define NL_STRIDE g_stride

inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline ull_t atomicMin(ull_t *p, ull_t v) { ull_t o = *p; if (v < o) *p = v; return o; }
inline ull_t atomicMax(ull_t *p, ull_t v) { ull_t o = *p; if (v > o) *p = v; return o; }
inline ull_t nlBits(double x) { ull_t u; std::memcpy(&u, &x, 8); return u; }
inline double nlFromBits(ull_t u) { double x; std::memcpy(&x, &u, 8); return x; }
#define NL_BITS(x) nlBits(x)
#define NL_FROM_BITS(u) nlFromBits(u)

#include "../../dftbplus_b200/csrc/n2_neighbour_driver.h"
#include "../../dftbplus_b200/csrc/k4_atom_population.cuh"

using namespace dftbp_b200;

namespace {

struct HostBackend {
  int nThreads;
  std::mt19937 rng;
  long long launches = 0;

  int alloc(void **p, size_t bytes) {
    *p = std::malloc(bytes ? bytes : 1);
    if (!*p) return 5;
    std::memset(*p, 0xA5, bytes);  // poison: nothing may rely on fresh memory being zero
    return 0;
  }
  void release(void *p) { std::free(p); }
  int upload(void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return 0; }
  int download(void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); return 0; }
  int fill(void *p, int v, size_t bytes) { std::memset(p, v, bytes); return 0; }
  int scan(const int *in, int *out, long long n) {
    long long s = 0;
    for (long long i = 0; i < n; ++i) { out[i] = (int)s; s += in[i]; }
    return 0;
  }
  template <class... P, class... A>
  int launch(void (*kernel)(P...), long long nItems, A... args) {
    if (nItems <= 0) return 0;
    std::vector<int> order(nThreads);
    std::iota(order.begin(), order.end(), 0);
    std::shuffle(order.begin(), order.end(), rng);
    g_stride = nThreads;
    for (int t : order) {
      g_tid = t;
      kernel(args...);
    }
    ++launches;
    return 0;
  }
};

nl::NeighbourList g_list;
int g_rankMax = nl::kRankMaxDefault;
int *g_nOrbSpecies = nullptr;

}  // namespace

extern "C" {

// out[0..7] = nAllAtom, mNeigh, sparseSize, nOrb, errAtom1, errAtom2, valid, launches
int host_nl_update(int nAtom, const double *coord0, const int *species0, int nSpecies, const int *nOrbSpecies,
                   int nCellVec, const double *cellVec, const double *rCellVec, double cutoff, double skCutoff,
                   int nThreads, unsigned seed, int symmetric, int *out) {
  HostBackend be{nThreads, std::mt19937(seed)};
  std::free(g_nOrbSpecies);
  g_nOrbSpecies = (int *)std::malloc(sizeof(int) * nSpecies);
  std::memcpy(g_nOrbSpecies, nOrbSpecies, sizeof(int) * nSpecies);
  const int rc = nl::updateNeighbourList(be, g_list, nAtom, coord0, species0, g_nOrbSpecies, nCellVec, cellVec,
                                         rCellVec, cutoff, skCutoff, symmetric != 0, g_rankMax);
  out[0] = g_list.nAllAtom; out[1] = g_list.mNeigh; out[2] = g_list.sparseSize; out[3] = g_list.nOrb;
  out[4] = g_list.errAtom1; out[5] = g_list.errAtom2; out[6] = g_list.valid ? 1 : 0; out[7] = (int)be.launches;
  return rc;
}

void host_nl_get(double *coords, int *species, int *img2CentCell, int *iCellVec, int *nNeighbour, int *iNeighbour,
                 double *neighDist2, int *nNeighbourSK, int *iPair, int *iAtomStart) {
  const nl::NeighbourList &L = g_list;
  const size_t rows = (size_t)(L.mNeigh + 1) * L.nAtom;
  std::memcpy(coords, L.coords.p, sizeof(double) * 3 * L.nAllAtom);
  std::memcpy(species, L.species.p, sizeof(int) * L.nAllAtom);
  std::memcpy(img2CentCell, L.img2CentCell.p, sizeof(int) * L.nAllAtom);
  std::memcpy(iCellVec, L.iCellVec.p, sizeof(int) * L.nAllAtom);
  std::memcpy(nNeighbour, L.nNeighbour.p, sizeof(int) * L.nAtom);
  std::memcpy(iNeighbour, L.iNeighbour.p, sizeof(int) * rows);
  std::memcpy(neighDist2, L.neighDist2.p, sizeof(double) * rows);
  std::memcpy(nNeighbourSK, L.nNeighbourSK.p, sizeof(int) * L.nAtom);
  std::memcpy(iPair, L.iPair.p, sizeof(int) * rows);
  std::memcpy(iAtomStart, L.iAtomStart.p, sizeof(int) * (L.nAtom + 1));
}

// foldCoordsToCell through foldSetup + nl_fold; returns 0, or 1 for a singular lattice
int host_nl_fold(int nAtom, double *coord, const double *latVec, int nThreads, unsigned seed) {
  nl::FoldParams fp;
  if (!nl::foldSetup(latVec, fp)) return 1;
  HostBackend be{nThreads, std::mt19937(seed)};
  return be.launch(nl::nl_fold, nAtom, nAtom, fp, coord);
}

// mullikenPerAtom's final sum and getOnsitePopulation (k4_atom_population.cuh)
void host_pop_atom_sum(int nAtom, int mOrb, const double *qPerOrbital, double *qq, int nThreads, unsigned seed) {
  HostBackend be{nThreads, std::mt19937(seed)};
  be.launch(pop::pop_atom_sum, nAtom, nAtom, mOrb, qPerOrbital, qq);
}
void host_pop_onsite(int nAtom, int ld, const int *iPair, const int *iAtomStart, const double *rho, double *qq,
                     int nThreads, unsigned seed) {
  HostBackend be{nThreads, std::mt19937(seed)};
  be.launch(pop::pop_onsite, nAtom, nAtom, ld, iPair, iAtomStart, rho, qq);
}

// the product's bisection and record heap sort on their own (reference known-answer tests)
int host_search_asc_real_geq(int n, const double *xx, double xVal, double tol) {
  return nl::searchAscRealGeq(n, xx - 1, xVal, tol);
}
// keys(n) = 1..n on entry (the index vector), vals(n) the values; sorted in place with the tolerance comparison
void host_heap_sort_records(int n, int *keys, double *vals, double tol) {
  nl::heapSortRecords(n, keys - 1, vals - 1, [tol](int, double vA, int, double vB) { return vA < vB - tol; });
}

// getNrOfNeighboursForAll for another cutoff on the last list
void host_nl_nr_of_neighbours(double cutoff, int *out, int nThreads, unsigned seed) {
  HostBackend be{nThreads, std::mt19937(seed)};
  be.launch(nl::nl_nr_of_neighbours, g_list.nAtom, g_list.nAtom, g_list.mNeigh + 1, cutoff * cutoff,
            (const int *)g_list.nNeighbour.p, (const double *)g_list.neighDist2.p, out);
}

// work list / block descriptors of the last list as adopt_neighbour_list builds them on the device
// (countBlocks + describeBlocks): hStart(nAtom + 1) -> returns the number of blocks
int host_geo_count(int *hStart, int nThreads, unsigned seed) {
  HostBackend be{nThreads, std::mt19937(seed)};
  if (nl::countBlocks(be, g_list, g_list.atomSize.p, g_list.atomOffset.p, hStart)) return -1;
  return hStart[g_list.nAtom];
}
// desc: 8 ints per block (the BlockDesc layout); returns hasDuplicates
int host_geo_describe(int nBlocks, int groupSize, int *workAll, int *desc, int *groupFirst, int nThreads, unsigned seed) {
  HostBackend be{nThreads, std::mt19937(seed)};
  static_assert(sizeof(nl::GeoDesc) == 32, "GeoDesc is 8 ints");
  const size_t nGroups = ((size_t)g_list.sparseSize + groupSize - 1) / groupSize;
  // exactly sized scratch copies (ASan: an index past the end must fail), poisoned like device memory
  void *w = nullptr, *d = nullptr, *gf = nullptr;
  be.alloc(&w, sizeof(int) * nBlocks); be.alloc(&d, sizeof(nl::GeoDesc) * nBlocks); be.alloc(&gf, sizeof(int) * nGroups);
  int hFlag = -1;
  const int rc = nl::describeBlocks(be, g_list, (const int *)g_list.atomOffset.p, nBlocks, groupSize, (int *)w,
                                    (nl::GeoDesc *)d, (int *)gf, g_list.sizes.p, &hFlag);
  std::memcpy(workAll, w, sizeof(int) * nBlocks);
  std::memcpy(desc, d, sizeof(nl::GeoDesc) * nBlocks);
  std::memcpy(groupFirst, gf, sizeof(int) * nGroups);
  be.release(w); be.release(d); be.release(gf);
  return rc ? -1 : hFlag;
}

// test hook: row length above which the counting is skipped (default nl::kRankMaxDefault)
void host_nl_set_rank_max(int v) { g_rankMax = v > 0 ? v : nl::kRankMaxDefault; }

// rows of the last update that went through the reference's heap sort (near-ties) / all rows
void host_nl_tie_rows(int *out) {
  int t = 0;
  for (int a = 0; a < g_list.nAtom; ++a) t += g_list.rowHasTies.p[a] != 0;
  out[0] = t;
  out[1] = g_list.nAtom;
}

// the grid the last update used: nb[3], lo[3], invW[3]
void host_nl_grid(int *nb, double *lo, double *invW) {
  for (int k = 0; k < 3; ++k) { nb[k] = g_list.grid.p->nb[k]; lo[k] = g_list.grid.p->lo[k]; invW[k] = g_list.grid.p->invW[k]; }
}

}  // extern "C"

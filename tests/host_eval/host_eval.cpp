// host_eval.cpp -- compiles the DEVICE evaluator headers of the library (sk_lane_eval.cuh,
// sk_gaunt_code.cuh: interpolation, harmonics, generated rotation code, Gaunt-contracted weights)
// for the HOST, so that the CPU test suite can check the product's own formulas against the oracle
// without a GPU.  Test infrastructure: under g++ the CUDA qualifiers of host_defines.h expand to
// nothing; the few device intrinsics the headers use are supplied below.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
template <class T>
inline T __ldg(const T *p) { return *p; }

#include "../../dftbplus_b200/csrc/sk_lane_eval.cuh"
#include "../../dftbplus_b200/csrc/sk_dmat_eval.cuh"
#include "../../dftbplus_b200/csrc/sk_table_prep.h"

using namespace dftbp_b200;

namespace {

// SpPairProg for a species pair given as shell lists (what set_orbitals builds, context.cu)
SpPairProg makeProg(int nSh1, const int *sh1, int nSh2, const int *sh2) {
  SpPairProg pg;
  std::memset(&pg, 0, sizeof(pg));
  int lmax1 = 0, lmax2 = 0, ind = 0;
  for (int a = 0; a < nSh1; ++a) {
    pg.nOrb1 += 2 * sh1[a] + 1;
    lmax1 = std::max(lmax1, sh1[a]);
    pg.shells1 |= (unsigned long long)sh1[a] << (4 * a);
    for (int b = 0; b < nSh2; ++b) ind += std::min(sh1[a], sh2[b]) + 1;
  }
  for (int b = 0; b < nSh2; ++b) {
    pg.nOrb2 += 2 * sh2[b] + 1;
    lmax2 = std::max(lmax2, sh2[b]);
    pg.shells2 |= (unsigned long long)sh2[b] << (4 * b);
  }
  pg.nInteg = ind;
  pg.nSh1 = nSh1;
  pg.nSh2 = nSh2;
  auto canon = [](int n, const int *sh) {
    for (int k = 0; k < n; ++k)
      if (sh[k] != k) return -1;
    return n - 1;
  };
  pg.canon1 = canon(nSh1, sh1);
  pg.canon2 = canon(nSh2, sh2);
  pg.nY = (lmax1 + lmax2 + 1) * (lmax1 + lmax2 + 1);
  fillShellProgram(pg);
  return pg;
}

}  // namespace

extern "C" {

// Block (nOrb2 x nOrb1, column-major) of one matrix for a pair at distance r along the unit vector
// d, from table tab(nGrid, nInteg) column-major, prepared exactly like set_sk_table does
// (prepareSkTable: cutoff parameters + tail coefficients).  forceGeneric != 0 takes the run-time
// shell loops even for canonical species (both code paths of skLaneEvalAny).  Returns the number
// of elements, -1 if the table is too short.
int host_eval_block(int nSh1, const int *sh1, int nSh2, const int *sh2, int nGrid, double dist, int method,
                    const double *tab, double r, const double *d, int forceGeneric, double *out) {
  SpPairProg pg = makeProg(nSh1, sh1, nSh2, sh2);
  if (forceGeneric) pg.canon1 = pg.canon2 = -1;
  SkTableDev t;
  std::vector<double> tail, padded(tab, tab + (size_t)nGrid * pg.nInteg);
  if (!prepareSkTable(t, nGrid, pg.nInteg, dist, method, tab, tail)) return -1;
  padded.resize(padded.size() + 8, 0.0);
  t.tab = padded.data(); t.tail = tail.data();
  pg.H = t; pg.S = t;
  skLaneEvalAny<3>(
      pg, t, r, d[0], d[1], d[2], [&](int first, int e, double v) { out[first + e] = v; }, [](int, int) {});
  return pg.nOrb1 * pg.nOrb2;
}

// The same block through the compact D-matrix evaluator of K1 (sk_dmat_eval.cuh: frame, D^1 / D^2,
// shell-pair loops; f shells through the Gaunt code).
int host_eval_block_d(int nSh1, const int *sh1, int nSh2, const int *sh2, int nGrid, double dist, int method,
                      const double *tab, double r, const double *d, int lmaxT, double *out) {
  SpPairProg pg = makeProg(nSh1, sh1, nSh2, sh2);
  SkTableDev t;
  std::vector<double> tail, padded(tab, tab + (size_t)nGrid * pg.nInteg);
  if (!prepareSkTable(t, nGrid, pg.nInteg, dist, method, tab, tail)) return -1;
  padded.resize(padded.size() + 8, 0.0);
  t.tab = padded.data(); t.tail = tail.data();
  pg.H = t; pg.S = t;
  LaneInterp li;
  laneInterpSetup(t, r, li);
  skd::DMat D;
  skd::buildD<true>(d[0], d[1], d[2], D);
  if (lmaxT == 3) {
    double Y[49];
    skg::harmonics<6>(d[0], d[1], d[2], Y);
    skd::evalBlock<3>(pg, t, li, D, Y, out);
  } else if (lmaxT == 2) {
    double Y[1] = {1.0};
    skd::evalBlock<2>(pg, t, li, D, Y, out);
  } else {
    double Y[1] = {1.0};
    skd::evalBlock<1>(pg, t, li, D, Y, out);
  }
  return pg.nOrb1 * pg.nOrb2;
}

// sum_e w_e E_e(d) through the weight-contracted form of the force kernel (skLaneWBuild + skLaneWDot)
double host_eval_wdot(int nSh1, const int *sh1, int nSh2, const int *sh2, int nGrid, double dist, int method,
                      const double *tab, double r, const double *d, int forceGeneric, const double *w) {
  SpPairProg pg = makeProg(nSh1, sh1, nSh2, sh2);
  if (forceGeneric) pg.canon1 = pg.canon2 = -1;
  SkTableDev t;
  std::vector<double> tail, padded(tab, tab + (size_t)nGrid * pg.nInteg);
  if (!prepareSkTable(t, nGrid, pg.nInteg, dist, method, tab, tail)) return 0.0;
  padded.resize(padded.size() + 8, 0.0);
  t.tab = padded.data(); t.tail = tail.data();
  pg.H = t; pg.S = t;
  std::vector<double> row(w, w + pg.nOrb1 * pg.nOrb2);
  skLaneWBuildAny<3>(pg, row.data());
  return skLaneWDotAny<3>(pg, t, r, d[0], d[1], d[2], row.data());
}

}  // extern "C"

// sk_table_prep.h -- host-side preparation of one SK table for the device evaluator
// (sk_lane_eval.cuh): cutoff parameters and the three coefficients of the 5th-order tail per
// integral, formed in the reference's own operation order.  Shared by set_sk_table (context.cu)
// and by the host build of the evaluator that the CPU tests use (tests/host_eval/).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace dftbp_b200 {

constexpr double kDistFudge = 1.0;      // common/accuracy.F90:44
constexpr double kDistFudgeOld = 0.3;   // common/accuracy.F90:48
constexpr double kDeltaR = 1e-5;        // dftb/slakoeqgrid.F90:89

// 8-point Neville tableau on a uniform grid, same operation order as polyInterU2
// (math/interpolation.F90:182-239) so that the tail coefficients below carry the reference's
// own rounding (their second derivative amplifies rounding by 1/deltaR^2, SURVEY.md hard part 1).
inline double nevilleUniform8(const double *xp, const double *yp, double xx) {
  const int nn = 8;
  double cc[8], dd[8], delta[7];
  for (int ii = 1; ii <= nn - 1; ++ii) delta[ii - 1] = 1.0 / (xp[ii] - xp[0]);
  for (int ii = 0; ii < nn; ++ii) cc[ii] = dd[ii] = yp[ii];
  int iCl = (int)std::ceil((xx - xp[0]) * delta[0]);
  iCl = std::min(std::max(iCl, 1), nn);
  double yy = yp[iCl - 1];
  iCl -= 1;
  for (int mm = 1; mm <= nn - 1; ++mm) {
    for (int ii = 1; ii <= nn - mm; ++ii) {
      double r2 = (dd[ii - 1] - cc[ii]) * delta[mm - 1];
      cc[ii - 1] = (xp[ii - 1] - xx) * r2;
      dd[ii - 1] = (xp[ii + mm - 1] - xx) * r2;
    }
    double dyy;
    if (2 * iCl < nn - mm) {
      dyy = cc[iCl];
    } else {
      dyy = dd[iCl - 1];
      iCl -= 1;
    }
    yy = yy + dyy;
  }
  return yy;
}

// y1, y1p, y1pp at the last grid point (SlakoEqGrid_interNew_, slakoeqgrid.F90:224-240)
// poly5ToZero, math/interpolation.F90:31-67: coefficients of xr^3, xr^4, xr^5 in the reference's order
inline void poly5Coefficients(double y0, double y0p, double y0pp, double dx, double *out) {
  const double dx1 = y0p * dx;
  const double dx2 = y0pp * dx * dx;
  out[0] = 10.0 * y0 - 4.0 * dx1 + 0.5 * dx2;
  out[1] = -15.0 * y0 + 7.0 * dx1 - 1.0 * dx2;
  out[2] = 6.0 * y0 - 3.0 * dx1 + 0.5 * dx2;
}

// skEqGridNew tail, slakoeqgrid.F90:222-241
inline void tailCoefficients(int nGrid, int nInteg, double dist, const double *tab, std::vector<double> &out) {
  out.assign(3 * (size_t)nInteg, 0.0);
  if (nGrid < 9) return;
  double xa[8], ya[8];
  for (int ii = 1; ii <= 8; ++ii) xa[ii - 1] = (double)(nGrid - 8 + ii) * dist;
  for (int c = 0; c < nInteg; ++c) {
    for (int ii = 0; ii < 8; ++ii) ya[ii] = tab[(size_t)(nGrid - 8 + ii) + (size_t)nGrid * c];
    double y0 = nevilleUniform8(xa, ya, xa[7] - kDeltaR);
    double y2 = nevilleUniform8(xa, ya, xa[7] + kDeltaR);
    double y1 = ya[7];
    poly5Coefficients(y1, (y2 - y0) / (2.0 * kDeltaR), (y2 + y0 - 2.0 * y1) / (kDeltaR * kDeltaR),
                      -1.0 * kDistFudge, &out[3 * c]);
  }
}

// skEqGridOld extrapolation, slakoeqgrid.F90:295-307: slope and curvature of the free cubic spline
// (math/interpolation.F90:74-120) at the last grid point start the 5th-order tail
inline void tailCoefficientsOld(int nGrid, int nInteg, double dist, const double *tab, double tailDx,
                         std::vector<double> &out) {
  out.assign(3 * (size_t)nInteg, 0.0);
  for (int c = 0; c < nInteg; ++c) {
    const double *col = tab + (size_t)nGrid * c;
    const double y0 = col[nGrid - 3], y1 = col[nGrid - 2], y2 = col[nGrid - 1];
    const double r1 = (y2 - y0) / (2.0 * dist);
    const double r2 = (y2 + y0 - 2.0 * y1) / (dist * dist);
    const double cc = 0.5 * r2, dx1 = 1.0 / dist;
    const double dd = (((y2 - y1) * dx1 - r1) * dx1 - 0.5 * r2) * dx1;
    const double y1p = (3.0 * dd * dist + 2.0 * cc) * dist + r1;
    const double y1pp = 6.0 * dd * dist + 2.0 * cc;
    poly5Coefficients(y2, y1p, y1pp, tailDx, &out[3 * c]);
  }
}

// Fills every field of t except the device pointers; tail receives (3, nInteg) coefficients.
// Returns false for tables too short for the interpolation (slakoeqgrid.F90:207-210, :271-273).
inline bool prepareSkTable(SkTableDev &t, int nGrid, int nInteg, double dist, int method, const double *tab,
                           std::vector<double> &tail) {
  const bool old = method == DFTBP_B200_SK_EQGRID_OLD;
  if (nGrid < (old ? 4 : 9)) return false;
  t.dist = dist; t.invDist = 1.0 / dist;
  t.nGrid = nGrid; t.nInteg = nInteg; t.method = method;
  if (old) {
    // mInd = leng + floor(distFudgeOld / incr); the tail runs over (mInd - leng - 1) * incr
    const int mInd = nGrid + (int)std::floor(kDistFudgeOld / dist);
    const double span = (double)(mInd - nGrid - 1) * dist;
    t.mIndM1 = mInd - 1;
    t.tailX0 = (double)(mInd - 1) * dist;
    t.tailInvDx = -1.0 / span;
    t.rMax = (double)(mInd - 1) * dist;  // first distance that certainly yields zero
    tailCoefficientsOld(nGrid, nInteg, dist, tab, -1.0 * span, tail);
  } else {
    t.rMax = (double)nGrid * dist + kDistFudge;
    t.mIndM1 = 0;
    t.tailX0 = t.rMax;
    t.tailInvDx = -1.0 / kDistFudge;
    tailCoefficients(nGrid, nInteg, dist, tab, tail);
  }
  return true;
}

// Position of the integrals of shell-pair type c = l1 * 4 + l2 in an exchange row ordered by type,
// types with l1, l2 <= lmax only (the device-side twin is skd::caseSlot<LMAX>, sk_dmat_eval.cuh).
inline int caseSlotHost(int c, int lmax) {
  int n = 0;
  for (int cc = 0; cc < c; ++cc) {
    const int l1 = cc >> 2, l2 = cc & 3;
    if (l1 <= lmax && l2 <= lmax) n += (l1 < l2 ? l1 : l2) + 1;
  }
  return n;
}

// Shell-pair program of a species pair (SpPairProg::shPair / integDst / caseOff / integSlot) from its
// shell lists; lmaxAll = highest angular momentum of any species of the run.
// Returns false if the pair has more shell pairs or integrals than the program holds.
inline bool fillShellProgram(SpPairProg &pg, int lmaxAll = 3) {
  for (int c = 0; c < 16; ++c) pg.caseOff[c] = 0xFFFF;
  for (int l = 0; l < 4; ++l) pg.shellOfL1[l] = -1;
  pg.simple = 1;
  for (int sh = 0; sh < pg.nSh1; ++sh) {
    const int l = (int)((pg.shells1 >> (4 * sh)) & 0xF);
    if (l > 3) return false;
    if (pg.shellOfL1[l] >= 0) pg.simple = 0;
    pg.shellOfL1[l] = (signed char)sh;
  }
  {
    int seen = 0;
    for (int sh = 0; sh < pg.nSh2; ++sh) {
      const int l = (int)((pg.shells2 >> (4 * sh)) & 0xF);
      if (l > 3) return false;
      if (seen & (1 << l)) pg.simple = 0;
      seen |= 1 << l;
    }
  }
  int ind = 0, iCol = 0, n = 0;
  for (int sh1 = 0; sh1 < pg.nSh1; ++sh1) {
    const int l1 = (int)((pg.shells1 >> (4 * sh1)) & 0xF);
    int iRow = 0;
    for (int sh2 = 0; sh2 < pg.nSh2; ++sh2) {
      const int l2 = (int)((pg.shells2 >> (4 * sh2)) & 0xF);
      const int lmin = l1 < l2 ? l1 : l2;
      if (n >= kMaxInteg || ind + lmin + 1 > kMaxInteg || l1 > 3 || l2 > 3) return false;
      const int first = iRow + pg.nOrb2 * iCol;
      pg.shPair[n++] = (unsigned short)((first << 4) | (l1 * 4 + l2));
      for (int k = 0; k <= lmin; ++k) pg.integDst[ind + k] = (unsigned short)(first + k);
      if (pg.simple) {
        pg.caseOff[l1 * 4 + l2] = (unsigned short)first;
        for (int k = 0; k <= lmin; ++k) pg.integSlot[ind + k] = (unsigned char)(caseSlotHost(l1 * 4 + l2, lmaxAll) + k);
      }
      ind += lmin + 1;
      iRow += 2 * l2 + 1;
    }
    iCol += 2 * l1 + 1;
  }
  pg.nShPair = n;
  return true;
}

// SlakoEqGrid_getCutoff, slakoeqgrid.F90:162-179
inline double skTableCutoff(int nGrid, double dist, int method) {
  return (double)nGrid * dist + (method == DFTBP_B200_SK_EQGRID_OLD ? kDistFudgeOld : kDistFudge);
}

}  // namespace dftbp_b200

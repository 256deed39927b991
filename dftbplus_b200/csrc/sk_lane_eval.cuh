// sk_lane_eval.cuh -- lane-per-pair evaluation of one diatomic H0 / S block.
//
// Replaces, for one neighbour pair, the body of buildDiatomicBlocks (dftb/nonscc.F90:406-421):
// getSKIntegrals (dftb/slakoeqgrid.F90:183-244: 8-point polynomial on the equidistant grid, 5th
// order tail) followed by rotateH0 (dftb/sk.F90:35-140).  Every THREAD owns one pair; the 32 pairs
// of a warp share the ordered species pair (pair_sort.cu), so the shell loops of rotateH0 and the
// dispatch on the shell-pair type are warp-uniform and the rotation is the generated straight-line
// code of sk_gaunt_code.cuh (coefficients as immediates, harmonics in registers).  Compared with
// the warp-per-pair evaluator (sk_eval.cuh) this removes the table-driven term loops whose
// integer / branch overhead made K1 issue bound (profiles/r01_ncu_k1_warp_per_pair.txt).
//
// The interpolation is the Lagrange form of the polynomial the reference builds with Neville
// (math/interpolation.F90:182-239): the eight weights depend on the distance only, so they are
// formed once per pair and table and every integral is 8 loads + 8 FMAs.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "sk_gaunt_code.cuh"

namespace dftbp_b200 {

struct LaneInterp {
  double w[8];  // Lagrange weights of rows j0 .. j0+7               (mode 0)
  double xr;    // tail abscissa / offset into the spline interval    (mode 1, 3)
  int j0;       // first row (0-based) of the window
  int mode;     // 0 = polynomial, 1 = 5th order tail, 2 = beyond the cutoff -> 0,
                // 3 = free cubic spline over the last grid interval (skEqGridOld only)
};

// SlakoEqGrid_interOld_ (OldSKInterpolation = Yes), slakoeqgrid.F90:248-317, the part that depends
// on r only.  The 3-point polynomial is expressed in the same 8-row window form as the 8-point one
// (five zero weights; the window is shifted so that it never leaves the column), so the integral
// loops below are shared by the two methods.
__device__ __forceinline__ void laneInterpSetupOld(const SkTableDev &t, double r, LaneInterp &li) {
  const int leng = t.nGrid;
  const int ind = (int)floor(r / t.dist);
  if (ind < leng - 1) {
    li.mode = 0;
    int iLast = min(leng, ind + 2);
    iLast = max(iLast, 3);
    const int jf = iLast - 3;                   // first row of the 3-point window, row j sits at (j+1)*dist
    const int jw = max(0, min(jf, leng - 8));   // 8-row window that contains it
    const int off = jf - jw;
    const double u = (r - (double)(jf + 1) * t.dist) * t.invDist;
    const double w0 = 0.5 * ((u - 1.0) * (u - 2.0)), w1 = -(u * (u - 2.0)), w2 = 0.5 * (u * (u - 1.0));
#pragma unroll
    for (int k = 0; k < 8; ++k) li.w[k] = k == off ? w0 : (k == off + 1 ? w1 : (k == off + 2 ? w2 : 0.0));
    li.j0 = jw;
  } else if (ind < leng) {
    li.mode = 3;  // between the penultimate and the last grid point
    li.xr = r - (double)(leng - 1) * t.dist;
  } else if (ind < t.mIndM1) {
    li.mode = 1;  // dr = rr - (mInd - 1) * incr; xr = dr * invdistFudge
    li.xr = (r - t.tailX0) * t.tailInvDx;
  } else {
    li.mode = 2;
  }
}

// SlakoEqGrid_interNew_, slakoeqgrid.F90:183-244, the part that depends on r only
__device__ __forceinline__ void laneInterpSetup(const SkTableDev &t, double r, LaneInterp &li) {
  li.j0 = 0;
  li.xr = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) li.w[j] = 0.0;
  if (t.method == DFTBP_B200_SK_EQGRID_OLD) {  // uniform per table
    laneInterpSetupOld(t, r, li);
    return;
  }
  if (r >= t.rMax) {
    li.mode = 2;
    return;
  }
  const int ind = (int)floor(r / t.dist);
  if (ind < t.nGrid) {
    li.mode = 0;
    int iLast = min(t.nGrid, ind + 4);
    iLast = max(iLast, 8);
    const int j0 = iLast - 8;  // row j sits at (j+1)*dist
    li.j0 = j0;
    const double u = (r - (double)(j0 + 1) * t.dist) * t.invDist;
    const double d0 = u, d1 = u - 1.0, d2 = u - 2.0, d3 = u - 3.0, d4 = u - 4.0, d5 = u - 5.0,
                 d6 = u - 6.0, d7 = u - 7.0;
    const double p1 = d0, p2 = p1 * d1, p3 = p2 * d2, p4 = p3 * d3, p5 = p4 * d4, p6 = p5 * d5,
                 p7 = p6 * d6;
    const double s6 = d7, s5 = s6 * d6, s4 = s5 * d5, s3 = s4 * d4, s2 = s3 * d3, s1 = s2 * d2,
                 s0 = s1 * d1;
    li.w[0] = (-1.0 / 5040.0) * s0;
    li.w[1] = (1.0 / 720.0) * (p1 * s1);
    li.w[2] = (-1.0 / 240.0) * (p2 * s2);
    li.w[3] = (1.0 / 144.0) * (p3 * s3);
    li.w[4] = (-1.0 / 144.0) * (p4 * s4);
    li.w[5] = (1.0 / 240.0) * (p5 * s5);
    li.w[6] = (-1.0 / 720.0) * (p6 * s6);
    li.w[7] = (1.0 / 5040.0) * p7;
  } else {
    li.mode = 1;  // xr = (rr - rMax) * invdx, invdx = -1 / distFudge
    li.xr = (r - t.tailX0) * t.tailInvDx;
  }
}

// value of integral column `col` (uniform across the warp)
__device__ __forceinline__ double laneInterp(const SkTableDev &t, const LaneInterp &li, int col) {
  if (li.mode == 0) {
    const double *__restrict__ y = t.tab + (size_t)col * t.nGrid + li.j0;
    double acc = li.w[0] * __ldg(y);
#pragma unroll
    for (int j = 1; j < 8; ++j) acc = fma(li.w[j], __ldg(y + j), acc);
    return acc;
  }
  if (li.mode == 1) {
    // poly5ToZero (math/interpolation.F90:31-67): the coefficients dd, ee, ff of the three highest
    // powers were formed on the host from (y, y', y'') at the start of the tail, in the reference's
    // operation order (context.cu: tailCoefficients / tailCoefficientsOld)
    const double dd = __ldg(t.tail + 3 * col), ee = __ldg(t.tail + 3 * col + 1), ff = __ldg(t.tail + 3 * col + 2);
    const double xr = li.xr;
    return ((ff * xr + ee) * xr + dd) * xr * xr * xr;
  }
  if (li.mode == 3) {
    // free cubic spline through the last three grid points (slakoeqgrid.F90:285-294,
    // freeCubicSpline math/interpolation.F90:74-120)
    const double *__restrict__ c = t.tab + (size_t)col * t.nGrid + (t.nGrid - 3);
    const double y0 = __ldg(c), y1 = __ldg(c + 1), y2 = __ldg(c + 2);
    const double incr = t.dist;
    const double y1p = (y2 - y0) / (2.0 * incr);
    const double y1pp = (y2 + y0 - 2.0 * y1) / (incr * incr);
    const double dx1 = 1.0 / incr;
    const double cc = 0.5 * y1pp;
    const double d3 = (((y2 - y1) * dx1 - y1p) * dx1 - 0.5 * y1pp) * dx1;
    const double xx = li.xr;
    return ((d3 * xx + cc) * xx + y1p) * xx + y1;
  }
  return 0.0;
}

// Evaluate one matrix (table t = the H or the S table of the warp-uniform species-pair program pg)
// of this lane's pair at distance r along the unit vector (dx, dy, dz), one atom-1 shell ("column
// group") after the other, like the outer loop of rotateH0 (sk.F90:72-137).  The columns of one
// atom-1 shell are a contiguous run of (2 l1 + 1) * nOrb2 elements of the column-major block:
//   emit(first, eLocal, value)  element first + eLocal of the block, eLocal = row + nOrb2 * (col - firstCol)
//                               inside the current group, first = nOrb2 * firstCol
//   endGroup(first, count)      the group is complete; it covers elements [first, first + count)
// Generic shell structure: run-time loops, warp-uniform dispatch on the shell-pair type.
template <int LMAX, class Emit, class EndGroup>
__device__ __forceinline__ void skLaneEval(const SpPairProg &pg, const SkTableDev &t, double r, double dx,
                                           double dy, double dz, Emit &&emit, EndGroup &&endGroup) {
  LaneInterp li;
  laneInterpSetup(t, r, li);
  double Y[(2 * LMAX + 1) * (2 * LMAX + 1)];
  skg::harmonics<2 * LMAX>(dx, dy, dz, Y);
  int ind = 0, iCol = 0;
  const int nSh1 = pg.nSh1, nSh2 = pg.nSh2, n2 = pg.nOrb2;
  const unsigned long long sh1s = pg.shells1, sh2s = pg.shells2;
  for (int sh1 = 0; sh1 < nSh1; ++sh1) {
    const int l1 = (int)((sh1s >> (4 * sh1)) & 0xF);
    int iRow = 0;
    for (int sh2 = 0; sh2 < nSh2; ++sh2) {
      const int l2 = (int)((sh2s >> (4 * sh2)) & 0xF);
      const int lmin = min(l1, l2);
      double v[4];
      const double zero[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (k <= LMAX && k <= lmin) ? laneInterp(t, li, ind + k) : 0.0;
      skg::blockDispatch<true, false, LMAX>(l1, l2, Y, v, zero, [&](int a, int b, double h, double) {
        emit(n2 * iCol, iRow + b + n2 * a, h);
      });
      ind += lmin + 1;
      iRow += 2 * l2 + 1;
    }
    endGroup(n2 * iCol, n2 * (2 * l1 + 1));
    iCol += 2 * l1 + 1;
  }
}

// ---- canonical species pairs: shells 0..LM1 on the atom, 0..LM2 on the neighbour (s, sp, spd, spdf) ---
// Everything about the shell structure is a compile-time constant, so the loops of rotateH0 unroll
// completely: the integrals live in registers with static indices and the element offsets become
// immediates.  The table loads of group g+1 are issued before endGroup(g) runs (the caller's
// write-out), so their latency hides behind it.

template <int N, class F>
__device__ __forceinline__ void staticFor(F &&f) {
  if constexpr (N > 0) {
    staticFor<N - 1>(f);
    f(std::integral_constant<int, N - 1>{});
  }
}

__host__ __device__ constexpr int canonMin(int a, int b) { return a < b ? a : b; }
// first integral of shell pair (l1, l2) when the neighbour has shells 0..LM2 (getFullTable order,
// dftbplus/parser.F90:4096-4127: atom-1 shells outer)
__host__ __device__ constexpr int canonIntegStart(int l1, int l2, int LM2) {
  int n = 0;
  for (int a = 0; a < l1; ++a)
    for (int b = 0; b <= LM2; ++b) n += canonMin(a, b) + 1;
  for (int b = 0; b < l2; ++b) n += canonMin(l1, b) + 1;
  return n;
}

// integrals [START, START + N) of table t (polynomial branch taken once; 8 N independent loads)
template <int START, int N, int NV>
__device__ __forceinline__ void laneInterpRange(const SkTableDev &t, const LaneInterp &li, double (&v)[NV]) {
  static_assert(N <= NV, "group buffer too small");
  if (li.mode == 0) {
    const double *__restrict__ y = t.tab + li.j0 + (size_t)START * t.nGrid;
    const size_t nGrid = (size_t)t.nGrid;
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const double *__restrict__ yc = y + c * nGrid;
      double acc = li.w[0] * __ldg(yc);
#pragma unroll
      for (int j = 1; j < 8; ++j) acc = fma(li.w[j], __ldg(yc + j), acc);
      v[c] = acc;
    }
  } else {
#pragma unroll
    for (int c = 0; c < N; ++c) v[c] = laneInterp(t, li, START + c);
  }
}

template <int LM1, int LM2, class Emit, class EndGroup>
__device__ __forceinline__ void skLaneEvalCanon(const SkTableDev &t, double r, double dx, double dy, double dz,
                                                Emit &&emit, EndGroup &&endGroup) {
  constexpr int LY = LM1 + LM2;
  constexpr int N2 = (LM2 + 1) * (LM2 + 1);
  // integrals of the largest group (atom-1 shell LM1)
  constexpr int NG = canonIntegStart(LM1 + 1, 0, LM2) - canonIntegStart(LM1, 0, LM2);
  LaneInterp li;
  laneInterpSetup(t, r, li);
  double Y[(LY + 1) * (LY + 1)];
  skg::harmonics<LY>(dx, dy, dz, Y);
  double vCur[NG], vNext[NG];
  laneInterpRange<0, canonIntegStart(1, 0, LM2), NG>(t, li, vCur);
  staticFor<LM1 + 1>([&](auto l1c) {
    constexpr int l1 = decltype(l1c)::value;
    constexpr int g0 = canonIntegStart(l1, 0, LM2);
    if constexpr (l1 < LM1)  // prefetch the next group's integrals
      laneInterpRange<canonIntegStart(l1 + 1, 0, LM2), canonIntegStart(l1 + 2, 0, LM2) - canonIntegStart(l1 + 1, 0, LM2), NG>(
          t, li, vNext);
    staticFor<LM2 + 1>([&](auto l2c) {
      constexpr int l2 = decltype(l2c)::value;
      constexpr int lmin = canonMin(l1, l2);
      constexpr int ind = canonIntegStart(l1, l2, LM2) - g0;
      double v4[4] = {0.0, 0.0, 0.0, 0.0};
      const double zero[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k <= lmin; ++k) v4[k] = vCur[ind + k];
      skg::blockT<l1, l2, true, false>(Y, v4, zero, [&](int a, int b, double h, double) {
        emit(N2 * l1 * l1, l2 * l2 + b + N2 * a, h);  // (iRow + b) + nOrb2 * (column inside the group)
      });
    });
    endGroup(N2 * l1 * l1, N2 * (2 * l1 + 1));
    if constexpr (l1 < LM1) {
#pragma unroll
      for (int c = 0; c < NG; ++c) vCur[c] = vNext[c];
    }
  });
}

// Warp-uniform dispatch: canonical species pairs up to spd go to the unrolled evaluator, anything
// else (f shells, repeated or unordered shells) to the generic loops.
template <int LMAX, class Emit, class EndGroup>
__device__ __forceinline__ void skLaneEvalAny(const SpPairProg &pg, const SkTableDev &t, double r, double dx,
                                              double dy, double dz, Emit &&emit, EndGroup &&endGroup) {
  const int canon = (pg.canon1 >= 0 && pg.canon2 >= 0 && pg.canon1 <= 2 && pg.canon2 <= 2)
                        ? pg.canon1 * 3 + pg.canon2 : -1;
  switch (canon) {
    case 0: skLaneEvalCanon<0, 0>(t, r, dx, dy, dz, emit, endGroup); break;
    case 1: if constexpr (LMAX >= 1) skLaneEvalCanon<0, 1>(t, r, dx, dy, dz, emit, endGroup); break;
    case 2: if constexpr (LMAX >= 2) skLaneEvalCanon<0, 2>(t, r, dx, dy, dz, emit, endGroup); break;
    case 3: if constexpr (LMAX >= 1) skLaneEvalCanon<1, 0>(t, r, dx, dy, dz, emit, endGroup); break;
    case 4: if constexpr (LMAX >= 1) skLaneEvalCanon<1, 1>(t, r, dx, dy, dz, emit, endGroup); break;
    case 5: if constexpr (LMAX >= 2) skLaneEvalCanon<1, 2>(t, r, dx, dy, dz, emit, endGroup); break;
    case 6: if constexpr (LMAX >= 2) skLaneEvalCanon<2, 0>(t, r, dx, dy, dz, emit, endGroup); break;
    case 7: if constexpr (LMAX >= 2) skLaneEvalCanon<2, 1>(t, r, dx, dy, dz, emit, endGroup); break;
    case 8: if constexpr (LMAX >= 2) skLaneEvalCanon<2, 2>(t, r, dx, dy, dz, emit, endGroup); break;
    default: skLaneEval<LMAX>(pg, t, r, dx, dy, dz, emit, endGroup);
  }
}


// ---- weight-contracted form (K5) ---------------------------------------------------------------------
// A contraction sum_e w_e E_e(d) of the block with fixed weights w (rho, E-rho, ...) is
//   sum over shell pairs, k:  A_k(r) * sum_M W_(k,M) S_(L_k M)(d),   W_(k,M) = sum_(a,b) G_k(a,b;M) w(a,b),
// and W does not depend on the bond vector.  The force kernel evaluates the block at six displaced
// bond vectors per pair (getFirstDerivFiniteDiff, dftb/nonscc.F90:428-484): it folds the weights
// into W once (skLaneWBuild, in place in the lane's row: W of a shell pair is never longer than
// the sub-block it replaces) and every displaced evaluation is then interpolation + harmonics + a
// (lmin + 1)(2 lmax + 1)-term dot product per shell pair instead of a full rotation.

// row[e] = weight of block element e (column-major nOrb2 x nOrb1) -> W, in place; generic shells
template <int LMAX>
__device__ __forceinline__ void skLaneWBuild(const SpPairProg &pg, double *row) {
  int iCol = 0;
  const int nSh1 = pg.nSh1, nSh2 = pg.nSh2, n2 = pg.nOrb2;
  const unsigned long long sh1s = pg.shells1, sh2s = pg.shells2;
  for (int sh1 = 0; sh1 < nSh1; ++sh1) {
    const int l1 = (int)((sh1s >> (4 * sh1)) & 0xF);
    int iRow = 0;
    for (int sh2 = 0; sh2 < nSh2; ++sh2) {
      const int l2 = (int)((sh2s >> (4 * sh2)) & 0xF);
      double *sub = row + iRow + n2 * iCol;
      skg::wbuildDispatch<LMAX>(
          l1, l2, [&](int a, int b) { return sub[b + n2 * a]; }, [&](int a, int b, double v) { sub[b + n2 * a] = v; });
      iRow += 2 * l2 + 1;
    }
    iCol += 2 * l1 + 1;
  }
}

// sum_e w_e E_e for table t at distance r along (dx, dy, dz); row holds W (skLaneWBuild)
template <int LMAX>
__device__ __forceinline__ double skLaneWDot(const SpPairProg &pg, const SkTableDev &t, double r, double dx,
                                             double dy, double dz, const double *row) {
  LaneInterp li;
  laneInterpSetup(t, r, li);
  double Y[(2 * LMAX + 1) * (2 * LMAX + 1)];
  skg::harmonics<2 * LMAX>(dx, dy, dz, Y);
  int ind = 0, iCol = 0;
  const int nSh1 = pg.nSh1, nSh2 = pg.nSh2, n2 = pg.nOrb2;
  const unsigned long long sh1s = pg.shells1, sh2s = pg.shells2;
  double acc = 0.0;
  for (int sh1 = 0; sh1 < nSh1; ++sh1) {
    const int l1 = (int)((sh1s >> (4 * sh1)) & 0xF);
    int iRow = 0;
    for (int sh2 = 0; sh2 < nSh2; ++sh2) {
      const int l2 = (int)((sh2s >> (4 * sh2)) & 0xF);
      const int lmin = min(l1, l2);
      double v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (k <= LMAX && k <= lmin) ? laneInterp(t, li, ind + k) : 0.0;
      const double *sub = row + iRow + n2 * iCol;
      acc += skg::wdotDispatch<LMAX>(l1, l2, Y, v, [&](int a, int b) { return sub[b + n2 * a]; });
      ind += lmin + 1;
      iRow += 2 * l2 + 1;
    }
    iCol += 2 * l1 + 1;
  }
  return acc;
}

template <int LM1, int LM2>
__device__ __forceinline__ void skLaneWBuildCanon(double *row) {
  constexpr int N2 = (LM2 + 1) * (LM2 + 1);
  staticFor<LM1 + 1>([&](auto l1c) {
    constexpr int l1 = decltype(l1c)::value;
    staticFor<LM2 + 1>([&](auto l2c) {
      constexpr int l2 = decltype(l2c)::value;
      double *sub = row + l2 * l2 + N2 * l1 * l1;
      skg::wbuildT<l1, l2>([&](int a, int b) { return sub[b + N2 * a]; },
                           [&](int a, int b, double v) { sub[b + N2 * a] = v; });
    });
  });
}

template <int LM1, int LM2>
__device__ __forceinline__ double skLaneWDotCanon(const SkTableDev &t, double r, double dx, double dy, double dz,
                                                  const double *row) {
  constexpr int LY = LM1 + LM2;
  constexpr int N2 = (LM2 + 1) * (LM2 + 1);
  constexpr int NI = canonIntegStart(LM1 + 1, 0, LM2);  // all integrals of the species pair
  LaneInterp li;
  laneInterpSetup(t, r, li);
  double v[NI];
  laneInterpRange<0, NI, NI>(t, li, v);
  double Y[(LY + 1) * (LY + 1)];
  skg::harmonics<LY>(dx, dy, dz, Y);
  double acc = 0.0;
  staticFor<LM1 + 1>([&](auto l1c) {
    constexpr int l1 = decltype(l1c)::value;
    staticFor<LM2 + 1>([&](auto l2c) {
      constexpr int l2 = decltype(l2c)::value;
      double a[4];
      skg::acoefT<l1, l2>(v + canonIntegStart(l1, l2, LM2), a);
      const double *sub = row + l2 * l2 + N2 * l1 * l1;
      acc += skg::wdotT<l1, l2>(Y, a, [&](int aa, int b) { return sub[b + N2 * aa]; });
    });
  });
  return acc;
}

__device__ __forceinline__ int canonCase(const SpPairProg &pg) {
  return (pg.canon1 >= 0 && pg.canon2 >= 0 && pg.canon1 <= 2 && pg.canon2 <= 2) ? pg.canon1 * 3 + pg.canon2 : -1;
}

template <int LMAX>
__device__ __forceinline__ void skLaneWBuildAny(const SpPairProg &pg, double *row) {
  switch (canonCase(pg)) {
    case 0: skLaneWBuildCanon<0, 0>(row); break;
    case 1: if constexpr (LMAX >= 1) skLaneWBuildCanon<0, 1>(row); break;
    case 2: if constexpr (LMAX >= 2) skLaneWBuildCanon<0, 2>(row); break;
    case 3: if constexpr (LMAX >= 1) skLaneWBuildCanon<1, 0>(row); break;
    case 4: if constexpr (LMAX >= 1) skLaneWBuildCanon<1, 1>(row); break;
    case 5: if constexpr (LMAX >= 2) skLaneWBuildCanon<1, 2>(row); break;
    case 6: if constexpr (LMAX >= 2) skLaneWBuildCanon<2, 0>(row); break;
    case 7: if constexpr (LMAX >= 2) skLaneWBuildCanon<2, 1>(row); break;
    case 8: if constexpr (LMAX >= 2) skLaneWBuildCanon<2, 2>(row); break;
    default: skLaneWBuild<LMAX>(pg, row);
  }
}

template <int LMAX>
__device__ __forceinline__ double skLaneWDotAny(const SpPairProg &pg, const SkTableDev &t, double r, double dx,
                                                double dy, double dz, const double *row) {
  switch (canonCase(pg)) {
    case 0: return skLaneWDotCanon<0, 0>(t, r, dx, dy, dz, row);
    case 1: if constexpr (LMAX >= 1) return skLaneWDotCanon<0, 1>(t, r, dx, dy, dz, row); break;
    case 2: if constexpr (LMAX >= 2) return skLaneWDotCanon<0, 2>(t, r, dx, dy, dz, row); break;
    case 3: if constexpr (LMAX >= 1) return skLaneWDotCanon<1, 0>(t, r, dx, dy, dz, row); break;
    case 4: if constexpr (LMAX >= 1) return skLaneWDotCanon<1, 1>(t, r, dx, dy, dz, row); break;
    case 5: if constexpr (LMAX >= 2) return skLaneWDotCanon<1, 2>(t, r, dx, dy, dz, row); break;
    case 6: if constexpr (LMAX >= 2) return skLaneWDotCanon<2, 0>(t, r, dx, dy, dz, row); break;
    case 7: if constexpr (LMAX >= 2) return skLaneWDotCanon<2, 1>(t, r, dx, dy, dz, row); break;
    case 8: if constexpr (LMAX >= 2) return skLaneWDotCanon<2, 2>(t, r, dx, dy, dz, row); break;
    default: break;
  }
  return skLaneWDot<LMAX>(pg, t, r, dx, dy, dz, row);
}

}  // namespace dftbp_b200

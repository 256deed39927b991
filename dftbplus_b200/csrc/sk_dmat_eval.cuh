// sk_dmat_eval.cuh -- compact lane-per-pair evaluator of one diatomic H0 / S block (s, p, d shells).
//
// rotateH0 (dftb/sk.F90:35-140) in its defining form: in the bond frame (u, v, w = bond direction)
// the two-centre block of a shell pair (l1, l2) is diagonal in m with the Slater-Koster integrals
// V_|m| on it (sigma, pi, delta), so in the fixed frame
//
//   block(b, a) = sgn * sum_m  V_|m| * D^l2(b, m) * D^l1(a, m),      m = -lmin .. lmin,
//
// where D^l(:, m) are the coefficients of the bond-frame orbital (l, m) in the fixed-frame tesseral
// basis (type/orbitals.F90:59-64: p = y,z,x; d = xy,yz,z2,xz,x2-y2) and sgn = (-1)^(l1+l2) for
// l1 > l2 (sk.F90:128-133).  D^1 is the frame itself, D^2 is bilinear in the frame vectors (a d
// orbital is a traceless quadratic form), the sum over +-m does not depend on the choice of u, v.
//
// Why this form in the kernel (round 2): the Gaunt-expanded straight-line code of sk_gaunt_code.cuh
// needs ~1000 FP64 instructions and 38 KB of SASS per spd-spd block, which streamed through the
// 32 KB instruction cache once per warp and item (profiles/r01g_ncu_k1_pair_build.txt: 2.65
// "no instruction" stalls per issue).  Here a block costs 328 FMAs, the nine shell-pair bodies
// together are a few KB, and D (34 doubles) is shared by the H0 and the S block of the pair.
// Shell pairs with an f shell keep the Gaunt code (callers pass the harmonics).
#pragma once
#include "sk_lane_eval.cuh"

namespace dftbp_b200 {
namespace skd {

// D^1 and D^2 column-wise: d1[m][i], d2[m][i]; m = 0 sigma, 1 / 2 the pi pair (u, v), 3 / 4 the
// delta pair (uv, uu - vv); i = fixed-frame orbital in the reference's order.
struct DMat {
  double d1[3][3];
  double d2[5][5];
};

// coefficients of the quadratic form sqrt(3) (a.r)(b.r), a.b = 0, in the d basis
__device__ __forceinline__ void dBilinear(const double (&a)[3], const double (&b)[3], double (&c)[5]) {
  const double s3 = 1.7320508075688772;
  c[0] = fma(a[0], b[1], a[1] * b[0]);   // xy
  c[1] = fma(a[1], b[2], a[2] * b[1]);   // yz
  c[2] = s3 * (a[2] * b[2]);             // z2
  c[3] = fma(a[0], b[2], a[2] * b[0]);   // xz
  c[4] = fma(a[0], b[0], -(a[1] * b[1]));  // x2-y2
}

// Orthonormal frame with w = (x, y, z) (branch-free construction, valid at both poles), D^1 and,
// if WITH_D, D^2.
template <bool WITH_D>
__device__ __forceinline__ void buildD(double x, double y, double z, DMat &D) {
  const double sg = z >= 0.0 ? 1.0 : -1.0;
  const double aa = -1.0 / (sg + z), bb = x * y * aa;
  const double u[3] = {fma(sg * x, x * aa, 1.0), sg * bb, -sg * x};
  const double v[3] = {bb, fma(y, y * aa, sg), -y};
  const double w[3] = {x, y, z};
  // p order (y, z, x)
  D.d1[0][0] = w[1]; D.d1[0][1] = w[2]; D.d1[0][2] = w[0];
  D.d1[1][0] = u[1]; D.d1[1][1] = u[2]; D.d1[1][2] = u[0];
  D.d1[2][0] = v[1]; D.d1[2][1] = v[2]; D.d1[2][2] = v[0];
  if constexpr (WITH_D) {
    const double s3 = 1.7320508075688772, hs3 = 0.8660254037844386;
    // sigma: (3 (w.r)^2 - r^2) / 2
    D.d2[0][0] = s3 * (w[0] * w[1]);
    D.d2[0][1] = s3 * (w[1] * w[2]);
    D.d2[0][2] = fma(1.5 * w[2], w[2], -0.5);
    D.d2[0][3] = s3 * (w[0] * w[2]);
    D.d2[0][4] = hs3 * fma(w[0], w[0], -(w[1] * w[1]));
    dBilinear(u, w, D.d2[1]);  // pi: sqrt(3) (u.r)(w.r), sqrt(3) (v.r)(w.r)
    dBilinear(v, w, D.d2[2]);
    dBilinear(u, v, D.d2[3]);  // delta: sqrt(3) (u.r)(v.r)
    // delta: sqrt(3)/2 ((u.r)^2 - (v.r)^2)
    D.d2[4][0] = fma(u[0], u[1], -(v[0] * v[1]));
    D.d2[4][1] = fma(u[1], u[2], -(v[1] * v[2]));
    D.d2[4][2] = hs3 * fma(u[2], u[2], -(v[2] * v[2]));
    D.d2[4][3] = fma(u[0], u[2], -(v[0] * v[2]));
    D.d2[4][4] = 0.5 * (fma(u[0], u[0], -(v[0] * v[0])) - fma(u[1], u[1], -(v[1] * v[1])));
  }
}

template <int L>
__device__ __forceinline__ double dEntry(const DMat &D, int m, int i) {
  if constexpr (L == 0) return 1.0;
  else if constexpr (L == 1) return D.d1[m][i];
  else return D.d2[m][i];
}

// Position of the integrals of shell-pair type c = l1 * 4 + l2 when the integrals of a pair are
// ordered by type (types with l1, l2 <= LMAX only); caseSlot(16) = their number.  Host twin:
// caseSlotHost (sk_table_prep.h).
template <int LMAX>
__host__ __device__ constexpr int caseSlot(int c) {
  int n = 0;
  for (int cc = 0; cc < c; ++cc) {
    const int l1 = cc >> 2, l2 = cc & 3;
    if (l1 <= LMAX && l2 <= LMAX) n += (l1 < l2 ? l1 : l2) + 1;
  }
  return n;
}

// One shell-pair block into dst (column-major, leading dimension n2): dst[b + n2 a], a = orbital of the
// atom's shell (l1), b = orbital of the neighbour's shell (l2); v[0..lmin] = the integrals of the
// shell pair in the order of getSKIntegrals (sigma, pi, delta).
template <int L1, int L2>
__device__ __forceinline__ void rotBlockV(const DMat &D, const double *v, double *dst, int n2) {
  constexpr int N1 = 2 * L1 + 1, N2 = 2 * L2 + 1, LMIN = L1 < L2 ? L1 : L2, NM = 2 * LMIN + 1;
  constexpr double sgn = (L1 > L2 && ((L1 + L2) & 1)) ? -1.0 : 1.0;
  double vm[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) vm[m] = sgn * v[(m + 1) >> 1];
#pragma unroll
  for (int a = 0; a < N1; ++a) {
    double c[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) c[m] = vm[m] * dEntry<L1>(D, m, a);
    double *col = dst + n2 * a;
#pragma unroll
    for (int b = 0; b < N2; ++b) {
      double e = c[0] * dEntry<L2>(D, 0, b);
#pragma unroll
      for (int m = 1; m < NM; ++m) e = fma(c[m], dEntry<L2>(D, m, b), e);
      col[b] = e;
    }
  }
}

// Same, with the integrals parked in dst[0..lmin] by the interpolation loop of evalBlock.
template <int L1, int L2>
__device__ __forceinline__ void rotBlock(const DMat &D, double *dst, int n2) {
  constexpr int N1 = 2 * L1 + 1, N2 = 2 * L2 + 1, LMIN = L1 < L2 ? L1 : L2, NM = 2 * LMIN + 1;
  constexpr double sgn = (L1 > L2 && ((L1 + L2) & 1)) ? -1.0 : 1.0;
  double v[LMIN + 1], vm[NM];
#pragma unroll
  for (int k = 0; k <= LMIN; ++k) v[k] = dst[k];
#pragma unroll
  for (int m = 0; m < NM; ++m) vm[m] = sgn * v[(m + 1) >> 1];
#pragma unroll
  for (int a = 0; a < N1; ++a) {
    double c[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) c[m] = vm[m] * dEntry<L1>(D, m, a);
    double *col = dst + n2 * a;
#pragma unroll
    for (int b = 0; b < N2; ++b) {
      double e = c[0] * dEntry<L2>(D, 0, b);
#pragma unroll
      for (int m = 1; m < NM; ++m) e = fma(c[m], dEntry<L2>(D, m, b), e);
      col[b] = e;
    }
  }
}

// The whole block of one matrix of this lane's pair into row (column-major nOrb2 x nOrb1).
//  1. all integrals of the species pair in one flat loop (every table load of the block is in
//     flight before the first use; round 2 profile of the per-shell-pair form: one exposed L2 round
//     trip per shell pair was the top stall), each parked inside the sub-block it belongs to;
//  2. the rotation, driven by the shell-pair program of the species pair (warp-uniform: the 32
//     pairs of a warp share the species pair), compile-time bodies per shell-pair type.
// Y: harmonics for the Gaunt code, used for f shells only.
template <int LMAX, int NY>
__device__ __forceinline__ void evalBlock(const SpPairProg &pg, const SkTableDev &t, const LaneInterp &li, const DMat &D,
                                          const double (&Y)[NY], double *row) {
  const int nInteg = pg.nInteg, n2 = pg.nOrb2;
  if (li.mode == 0) {
    const double *__restrict__ y = t.tab + li.j0;
    const size_t nGrid = (size_t)t.nGrid;
#pragma unroll 2
    for (int k = 0; k < nInteg; ++k) {
      const double *__restrict__ yc = y + k * nGrid;
      double acc = li.w[0] * __ldg(yc);
#pragma unroll
      for (int j = 1; j < 8; ++j) acc = fma(li.w[j], __ldg(yc + j), acc);
      row[pg.integDst[k]] = acc;
    }
  } else {
    for (int k = 0; k < nInteg; ++k) row[pg.integDst[k]] = laneInterp(t, li, k);
  }
  const int nShPair = pg.nShPair;
  for (int i = 0; i < nShPair; ++i) {
    const int e = pg.shPair[i];
    double *dst = row + (e >> 4);
    switch (e & 15) {
      case 0: rotBlock<0, 0>(D, dst, n2); break;
      case 1: if constexpr (LMAX >= 1) rotBlock<0, 1>(D, dst, n2); break;
      case 2: if constexpr (LMAX >= 2) rotBlock<0, 2>(D, dst, n2); break;
      case 4: if constexpr (LMAX >= 1) rotBlock<1, 0>(D, dst, n2); break;
      case 5: if constexpr (LMAX >= 1) rotBlock<1, 1>(D, dst, n2); break;
      case 6: if constexpr (LMAX >= 2) rotBlock<1, 2>(D, dst, n2); break;
      case 8: if constexpr (LMAX >= 2) rotBlock<2, 0>(D, dst, n2); break;
      case 9: if constexpr (LMAX >= 2) rotBlock<2, 1>(D, dst, n2); break;
      case 10: if constexpr (LMAX >= 2) rotBlock<2, 2>(D, dst, n2); break;
      default:
        if constexpr (LMAX >= 3) {  // a shell pair with an f shell: generated Gaunt code
          const int l1 = (e >> 2) & 3, l2 = e & 3, lmin = min(l1, l2);
          double v4[4] = {0.0, 0.0, 0.0, 0.0};
          const double zero[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k <= lmin) v4[k] = dst[k];
          skg::blockDispatch<true, false, LMAX>(l1, l2, Y, v4, zero,
                                                [&](int a, int b, double h, double) { dst[b + n2 * a] = h; });
        }
        break;
    }
  }
}

}  // namespace skd
}  // namespace dftbp_b200

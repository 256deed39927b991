// sk_eval.cuh -- warp-cooperative evaluation of one diatomic H0 / S block.
//
// Replaces, for one neighbour pair, the body of buildDiatomicBlocks (dftb/nonscc.F90:406-421):
// getSKIntegrals (dftb/slakoeqgrid.F90:183-244, 8-point polynomial on the equidistant grid, 5th
// order tail) followed by rotateH0 (dftb/sk.F90:35-140).  One warp owns one pair:
//   phase 1  lane i < 2*nInteg interpolates one column of the H or the S table
//            (Lagrange form of the same 8-point polynomial the reference builds with Neville);
//   phase 2  lane i < nY evaluates one tesseral harmonic S_LM of the bond direction;
//   phase 3  lane i < 2*nInteg forms one expansion coefficient A_L of a shell pair;
//   phase 4  lane e evaluates output elements e, e+32, ... as short Gaunt term lists, so the
//            stores of a block are contiguous across the warp (tools/gen_sk_gaunt.py).
// No per-shell-pair branching, s/p/d/f share the code path, H and S share phases 2 and 4.
#pragma once
#include "common.cuh"

namespace dftbp_b200 {

struct YDescDev {
  double c[4];
  int m, par, isSin, pad;
};

struct __align__(16) WarpScratch {
  double Y[49];
  double pad_;
  double V[2 * kMaxInteg];
  double A[2 * kMaxInteg];
};

// value of column `col` of table t at distance r (SlakoEqGrid_interNew_, slakoeqgrid.F90:183-244)
__device__ __forceinline__ double skInterpColumn(const SkTableDev &t, int col, double r) {
  if (r >= t.rMax) return 0.0;
  const int ind = (int)floor(r / t.dist);
  const double *__restrict__ colp = t.tab + (size_t)col * t.nGrid;
  if (ind < t.nGrid) {
    int iLast = min(t.nGrid, ind + 4);
    iLast = max(iLast, 8);
    const int j0 = iLast - 8;  // 0-based first row of the window; row j sits at (j+1)*dist
    const double u = (r - (double)(j0 + 1) * t.dist) * t.invDist;
    const double d0 = u, d1 = u - 1.0, d2 = u - 2.0, d3 = u - 3.0, d4 = u - 4.0, d5 = u - 5.0,
                 d6 = u - 6.0, d7 = u - 7.0;
    // prefix / suffix products of (u - k)
    const double p1 = d0, p2 = p1 * d1, p3 = p2 * d2, p4 = p3 * d3, p5 = p4 * d4, p6 = p5 * d5,
                 p7 = p6 * d6;
    const double s6 = d7, s5 = s6 * d6, s4 = s5 * d5, s3 = s4 * d4, s2 = s3 * d3, s1 = s2 * d2,
                 s0 = s1 * d1;
    const double *__restrict__ y = colp + j0;
    double acc;
    acc = (-1.0 / 5040.0) * s0 * __ldg(y + 0);
    acc = fma((1.0 / 720.0) * (p1 * s1), __ldg(y + 1), acc);
    acc = fma((-1.0 / 240.0) * (p2 * s2), __ldg(y + 2), acc);
    acc = fma((1.0 / 144.0) * (p3 * s3), __ldg(y + 3), acc);
    acc = fma((-1.0 / 144.0) * (p4 * s4), __ldg(y + 4), acc);
    acc = fma((1.0 / 240.0) * (p5 * s5), __ldg(y + 5), acc);
    acc = fma((-1.0 / 720.0) * (p6 * s6), __ldg(y + 6), acc);
    acc = fma((1.0 / 5040.0) * p7, __ldg(y + 7), acc);
    return acc;
  }
  // 5th-order tail to zero over distFudge = 1 (poly5ToZero, math/interpolation.F90:31-67, with
  // dx = -1, invdx = -1); y1, y1p, y1pp were prepared on the host in the reference's own
  // operation order (context.cu: tailCoefficients).
  const double y1 = __ldg(t.tail + 3 * col), y1p = __ldg(t.tail + 3 * col + 1),
               y1pp = __ldg(t.tail + 3 * col + 2);
  const double dx1 = -y1p, dx2 = y1pp;
  const double dd = 10.0 * y1 - 4.0 * dx1 + 0.5 * dx2;
  const double ee = -15.0 * y1 + 7.0 * dx1 - dx2;
  const double ff = 6.0 * y1 - 3.0 * dx1 + 0.5 * dx2;
  const double xr = -(r - t.rMax);
  return ((ff * xr + ee) * xr + dd) * xr * xr * xr;
}

// S_LM(x, y, z) on the unit sphere for harmonic index i = L*L + L + M
__device__ __forceinline__ double tesseralLM(const YDescDev *__restrict__ yd, int i, double x,
                                             double y, double z) {
  const double2 c01 = __ldg(reinterpret_cast<const double2 *>(yd[i].c));
  const double2 c23 = __ldg(reinterpret_cast<const double2 *>(yd[i].c) + 1);
  const int4 q = __ldg(reinterpret_cast<const int4 *>(&yd[i].m));
  const double z2 = z * z;
  double poly = fma(z2, fma(z2, fma(z2, c23.y, c23.x), c01.y), c01.x);
  if (q.y) poly *= z;
  double re = 1.0, im = 0.0, br = x, bi = y;
  int mm = q.x;
#pragma unroll
  for (int bit = 0; bit < 3; ++bit) {
    if (mm & 1) {
      const double nr = re * br - im * bi;
      im = re * bi + im * br;
      re = nr;
    }
    const double b2 = br * br - bi * bi;
    bi = 2.0 * br * bi;
    br = b2;
    mm >>= 1;
  }
  return poly * (q.z ? im : re);
}

struct SkDeviceTables {
  const SpPairProg *progs;
  const ADesc *aDesc;
  const int *elemStart;
  const DevTerm *terms;
  const YDescDev *yDesc;
};

// Evaluate the block of pair program pg at distance r along unit vector (dx, dy, dz) and call
// f(e, h, s) for every element e = row + nOrb2*col owned by this lane.  All 32 lanes must call.
template <bool DO_H, bool DO_S, class F>
__device__ __forceinline__ void skEvalBlock(const SkDeviceTables &tb, const SpPairProg &pg, double r,
                                            double dx, double dy, double dz, WarpScratch &ws,
                                            int lane, F &&f) {
  const int nI = pg.nInteg;
  // phase 1: integrals.  slots [0, nI) = H columns, [kMaxInteg, kMaxInteg + nI) = S columns
  for (int idx = lane; idx < 2 * nI; idx += 32) {
    const bool isS = idx >= nI;
    const int col = isS ? idx - nI : idx;
    if ((isS && DO_S) || (!isS && DO_H))
      ws.V[(isS ? kMaxInteg : 0) + col] = skInterpColumn(isS ? pg.S : pg.H, col, r);
  }
  // phase 2: harmonics of the direction
  for (int i = lane; i < pg.nY; i += 32) ws.Y[i] = tesseralLM(tb.yDesc, i, dx, dy, dz);
  __syncwarp();
  // phase 3: A_L per shell pair
  for (int idx = lane; idx < 2 * nI; idx += 32) {
    const bool isS = idx >= nI;
    const int s = isS ? idx - nI : idx;
    if ((isS && DO_S) || (!isS && DO_H)) {
      const ADesc *ad = tb.aDesc + pg.aOff + s;
      const double2 t01 = __ldg(reinterpret_cast<const double2 *>(ad->t));
      const double2 t23 = __ldg(reinterpret_cast<const double2 *>(ad->t) + 1);
      const int2 bc = __ldg(reinterpret_cast<const int2 *>(&ad->vBase));
      const double *v = ws.V + (isS ? kMaxInteg : 0) + bc.x;
      double a = t01.x * v[0];
      if (bc.y > 1) a = fma(t01.y, v[1], a);
      if (bc.y > 2) a = fma(t23.x, v[2], a);
      if (bc.y > 3) a = fma(t23.y, v[3], a);
      ws.A[(isS ? kMaxInteg : 0) + s] = a;
    }
  }
  __syncwarp();
  // phase 4: elements
  const int nElem = pg.nOrb1 * pg.nOrb2;
  const int *__restrict__ es = tb.elemStart + pg.elemOff;
  for (int e = lane; e < nElem; e += 32) {
    const int t0 = __ldg(es + e), t1 = __ldg(es + e + 1);
    double h = 0.0, s = 0.0;
    for (int t = t0; t < t1; ++t) {
      const int4 raw = __ldg(reinterpret_cast<const int4 *>(tb.terms + t));
      const double coef = __hiloint2double(raw.y, raw.x);
      const double cy = coef * ws.Y[raw.z];
      if (DO_H) h = fma(cy, ws.A[raw.w], h);
      if (DO_S) s = fma(cy, ws.A[kMaxInteg + raw.w], s);
    }
    f(e, h, s);
  }
  __syncwarp();  // scratch may be reused by the caller's next evaluation
}

}  // namespace dftbp_b200

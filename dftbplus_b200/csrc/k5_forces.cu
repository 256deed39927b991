// k5_forces.cu -- derivative_shift (dftb/forces.F90:27-38, 119-217, 548-669) with the reference's
// finite-difference derivator (getFirstDerivFiniteDiff, dftb/nonscc.F90:428-484).
//
// One warp per off-site pair (iAtom1 /= iAtom2f).  The reference materialises H' and S'
// (mOrb x mOrb x 3 each, 12 full SK evaluations per pair) and then contracts them; here the six
// displaced blocks are contracted on the fly in registers -- no dH/dR ever exists in memory:
//
//   g(j) = 2 [ sum rho.H'_j - sum Erho.S'_j + sum_spin sum rho_s . 1/2 (S'_j V1_s + V2_s S'_j) ]
//        = 2 sum_e [ wH_e H'_j,e + wS_e S'_j,e ],
//   wH = rho_1,  wS = -Erho + sum_spin 1/2 (rho_s V1_s^T + V2_s^T rho_s)        (independent of j)
//   X'_j = (X(v - delta e_j) - X(v + delta e_j)) / (2 delta)                      (nonscc.F90:471-481)
//
// so g(j) = (1/delta) sum_e [ wH_e (H_e(-) - H_e(+)) + wS_e (S_e(-) - S_e(+)) ], one warp
// reduction per Cartesian direction, then deriv(:,iAtom1) += g, deriv(:,iAtom2f) -= g.
// FP64 CUDA-core bound (about 6 block evaluations of sk_eval.cuh per pair), not HBM bound.
#include "sk_eval.cuh"

using namespace dftbp_b200;

namespace {

constexpr int kMaxElemPerLane = (kMaxOrb * kMaxOrb + 31) / 32;  // 8

struct K5Params {
  SkDeviceTables tb;
  const int *work;
  long long nWork;
  int ld, nSpecies, mOrb, nAtom, nSpin, sparseSize;
  const double *coords;
  const int *species, *iNeighbour, *iPair, *img2CentCell;
  const double *DM, *EDM, *shift;
  double delta;
  double *deriv;
};

template <int NE>  // elements per lane: ceil(mOrb^2 / 32)
__global__ void __launch_bounds__(32 * kWarpsPerCta)
k5_derivative_shift(const K5Params p) {
  __shared__ WarpScratch scratch[kWarpsPerCta];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  WarpScratch &ws = scratch[wIn];
  const long long nWarps = (long long)gridDim.x * kWarpsPerCta;
  for (long long w = (long long)blockIdx.x * kWarpsPerCta + wIn; w < p.nWork; w += nWarps) {
    const int flat = __ldg(p.work + w);
    const int iAt1 = flat / p.ld, iNeigh = flat - iAt1 * p.ld;
    if (iNeigh == 0) continue;  // loop from 1 as no contribution from the atom itself
    const int iAt2 = __ldg(p.iNeighbour + flat) - 1;
    const int iAt2f = __ldg(p.img2CentCell + iAt2) - 1;
    if (iAt1 == iAt2f) continue;
    const int sp1 = __ldg(p.species + iAt1) - 1, sp2 = __ldg(p.species + iAt2) - 1;
    const int ind = __ldg(p.iPair + flat);
    const SpPairProg &pg = p.tb.progs[sp1 + p.nSpecies * sp2];
    const int n1 = pg.nOrb1, n2 = pg.nOrb2, nElem = n1 * n2;

    // contraction weights of this lane's elements
    double wH[NE], wS[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int e = lane + 32 * i;
      wH[i] = 0.0;
      wS[i] = 0.0;
      if (e < nElem) {
        wH[i] = __ldg(p.DM + ind + e);
        double s = -__ldg(p.EDM + ind + e);
        if (p.shift) {
          const int a = e / n2, b = e - a * n2;  // element (row b, col a) of the n2 x n1 block
          const int mo2 = p.mOrb * p.mOrb;
          for (int iSpin = 0; iSpin < p.nSpin; ++iSpin) {
            const double *__restrict__ rho = p.DM + (size_t)p.sparseSize * iSpin + ind;
            const double *__restrict__ V1 = p.shift + (size_t)mo2 * (iAt1 + (size_t)p.nAtom * iSpin);
            const double *__restrict__ V2 = p.shift + (size_t)mo2 * (iAt2f + (size_t)p.nAtom * iSpin);
            double x = 0.0;
            // (rho V1^T)(b,a) = sum_c rho(b,c) V1(a,c);  (V2^T rho)(b,a) = sum_c V2(c,b) rho(c,a)
            for (int c = 0; c < n1; ++c) x = fma(__ldg(rho + b + n2 * c), __ldg(V1 + a + p.mOrb * c), x);
            for (int c = 0; c < n2; ++c) x = fma(__ldg(V2 + c + p.mOrb * b), __ldg(rho + c + n2 * a), x);
            s = fma(0.5, x, s);
          }
        }
        wS[i] = s;
      }
    }

    const double v0 = __ldg(p.coords + 3 * iAt2) - __ldg(p.coords + 3 * iAt1);
    const double v1 = __ldg(p.coords + 3 * iAt2 + 1) - __ldg(p.coords + 3 * iAt1 + 1);
    const double v2 = __ldg(p.coords + 3 * iAt2 + 2) - __ldg(p.coords + 3 * iAt1 + 2);
    double g[3];
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
      double acc = 0.0;
#pragma unroll
      for (int ii = 1; ii <= 2; ++ii) {
        // vect(jj) = vect(jj) - real(2*ii-3) * deltaXDiff : ii=1 -> +delta, ii=2 -> -delta
        const double sgn = (ii == 1) ? 1.0 : -1.0;
        double d0 = v0, d1 = v1, d2 = v2;
        if (jj == 0) d0 = v0 + sgn * p.delta;
        if (jj == 1) d1 = v1 + sgn * p.delta;
        if (jj == 2) d2 = v2 + sgn * p.delta;
        const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double wgt = (ii == 1) ? -1.0 : 1.0;  // (tmp(ii=2) - tmp(ii=1))
        skEvalBlock<true, true>(p.tb, pg, dist, d0 / dist, d1 / dist, d2 / dist, ws, lane,
                                [&](int e, double h, double s) {
                                  const int i = e >> 5;
#pragma unroll
                                  for (int k = 0; k < NE; ++k)
                                    if (k == i) acc = fma(wgt, fma(wH[k], h, wS[k] * s), acc);
                                });
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      g[jj] = acc / p.delta;
    }
    if (lane < 3) {
      const double gv = lane == 0 ? g[0] : (lane == 1 ? g[1] : g[2]);
      atomicAdd(p.deriv + lane + 3 * iAt1, gv);
      atomicAdd(p.deriv + lane + 3 * iAt2f, -gv);
    }
  }
}

}  // namespace

extern "C" int dftbp_b200_derivative_shift(dftbp_b200_ctx *ctx, double *deriv, const double *DM, int nSpin,
                                           const double *EDM, const double *shift, int diffType,
                                           double delta) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "derivative_shift: no geometry set");
  DFTBP_CHECK(ctx, deriv && DM && EDM, "derivative_shift: null argument");
  DFTBP_CHECK(ctx, nSpin == 1 || nSpin == 2 || nSpin == 4, "derivative_shift: nSpin must be 1, 2 or 4");
  if (diffType == DFTBP_B200_DIFF_RICHARDSON)
    return fail(ctx, DFTBP_B200_ENOTIMPL,
                "Differentiation = Richardson (nonscc.F90:488-600) is not built yet; use FiniteDiff");
  DFTBP_CHECK(ctx, diffType == DFTBP_B200_DIFF_FINITE, "derivative_shift: invalid differentiator type");
  DFTBP_TRY(syncPrograms(ctx));
  K5Params p;
  p.tb = {ctx->dProgs, ctx->dADesc, ctx->dElemStart, ctx->dTerms, (const YDescDev *)ctx->dYDesc};
  p.work = g.workLoad; p.nWork = g.nWorkLoad; p.ld = g.mNeigh + 1; p.nSpecies = ctx->nSpecies;
  p.mOrb = ctx->mOrb; p.nAtom = g.nAtom; p.nSpin = nSpin; p.sparseSize = g.sparseSize;
  p.coords = g.coords; p.species = g.species; p.iNeighbour = g.iNeighbour; p.iPair = g.iPair;
  p.img2CentCell = g.img2CentCell;
  // nonscc.F90:43: deltaXDiff = epsilon(1.0_dp)**0.25_dp
  p.delta = delta > 0.0 ? delta : 1.220703125e-4;
  // the non-SCC routine takes DM(:) only; the block routine DM(:, nSpin)
  const int nDm = shift ? nSpin : 1;
  DFTBP_TRY(devIn(ctx, kSlotA, DM, (size_t)g.sparseSize * nDm, &p.DM));
  DFTBP_TRY(devIn(ctx, kSlotB, EDM, (size_t)g.sparseSize, &p.EDM));
  p.shift = nullptr;
  if (shift)
    DFTBP_TRY(devIn(ctx, kSlotC, shift, (size_t)ctx->mOrb * ctx->mOrb * g.nAtom * nSpin, &p.shift));
  DFTBP_TRY(devOut(ctx, kSlotD, deriv, (size_t)3 * g.nAtom, false, &p.deriv));
  DFTBP_CUDA(ctx, cudaMemsetAsync(p.deriv, 0, sizeof(double) * 3 * g.nAtom, ctx->stream));
  if (p.nWork > 0) {
    ScopedTimer timer(ctx, kTimerForceCalc);
    const int grid = gridFor(ctx, p.nWork, kWarpsPerCta, 8 * 4);
    const int ne = (ctx->mOrb * ctx->mOrb + 31) / 32;
    if (ne <= 1) k5_derivative_shift<1><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
    else if (ne <= 3) k5_derivative_shift<3><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
    else k5_derivative_shift<kMaxElemPerLane><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
    ctx->launches++;
  }
  return finish(ctx);
}

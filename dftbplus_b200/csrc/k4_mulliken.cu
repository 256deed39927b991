// k4_mulliken.cu -- mullikenPerOrbital / mullikenPerBlock (dftb/populations.F90:120-187, 306-375).
//
// HBM-streaming: 16 bytes read per sparse element (over, rho), output is tiny (mOrb x nAtom).
// One warp per sparse block.  The element products are staged in the warp's shared-memory tile,
// then lane a forms the column sum for orbital a of iAtom1 and lane b the row sum for orbital b
// of the folded neighbour (segmented sums inside the warp), so only n1 + n2 global reductions
// leave the warp per block instead of 2*n1*n2.  qq ACCUMULATES like the reference.
#include "common.cuh"

using namespace dftbp_b200;

namespace {

struct K4Params {
  const int *work;
  long long nWork;
  int ld, mOrb;
  const int *iNeighbour, *iPair, *img2CentCell, *iAtomStart;
};

__global__ void __launch_bounds__(256) k4_mulliken_orbital(const K4Params p, const double *__restrict__ over,
                                                           const double *__restrict__ rho,
                                                           double *__restrict__ qq) {
  __shared__ double tile[8][kMaxOrb * kMaxOrb];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  double *t = tile[wIn];
  const long long nWarps = (long long)gridDim.x * 8;
  for (long long w = (long long)blockIdx.x * 8 + wIn; w < p.nWork; w += nWarps) {
    const int flat = __ldg(p.work + w);
    const int iAt1 = flat / p.ld;
    const int iAt2f = __ldg(p.img2CentCell + __ldg(p.iNeighbour + flat) - 1) - 1;
    const int n1 = __ldg(p.iAtomStart + iAt1 + 1) - __ldg(p.iAtomStart + iAt1);
    const int n2 = __ldg(p.iAtomStart + iAt2f + 1) - __ldg(p.iAtomStart + iAt2f);
    const int ind = __ldg(p.iPair + flat);
    const int nElem = n1 * n2;
    for (int e = lane; e < nElem; e += 32) t[e] = __ldg(over + ind + e) * __ldg(rho + ind + e);
    __syncwarp();
    if (lane < n1) {  // sum(sqrTmp, dim=1) -> qq(1:nOrb1, iAtom1)
      double s = 0.0;
      for (int b = 0; b < n2; ++b) s += t[b + n2 * lane];
      atomicAdd(qq + lane + p.mOrb * iAt1, s);
    }
    if (iAt1 != iAt2f && lane >= 16 && lane - 16 < n2) {  // sum(sqrTmp, dim=2) -> qq(1:nOrb2, iAtom2f)
      const int b = lane - 16;
      double s = 0.0;
      for (int a = 0; a < n1; ++a) s += t[b + n2 * a];
      atomicAdd(qq + b + p.mOrb * iAt2f, s);
    }
    __syncwarp();
  }
}

// qq(:,:,iAtom1) += 0.5*(rho^T S + S^T rho);  qq(:,:,iAtom2f) += 0.5*(rho S^T + S rho^T)
__global__ void __launch_bounds__(256) k4_mulliken_block(const K4Params p, const double *__restrict__ over,
                                                         const double *__restrict__ rho,
                                                         double *__restrict__ qq) {
  __shared__ double tS[8][kMaxOrb * kMaxOrb], tP[8][kMaxOrb * kMaxOrb];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  double *S = tS[wIn], *P = tP[wIn];
  const long long nWarps = (long long)gridDim.x * 8;
  for (long long w = (long long)blockIdx.x * 8 + wIn; w < p.nWork; w += nWarps) {
    const int flat = __ldg(p.work + w);
    const int iAt1 = flat / p.ld;
    const int iAt2f = __ldg(p.img2CentCell + __ldg(p.iNeighbour + flat) - 1) - 1;
    const int n1 = __ldg(p.iAtomStart + iAt1 + 1) - __ldg(p.iAtomStart + iAt1);
    const int n2 = __ldg(p.iAtomStart + iAt2f + 1) - __ldg(p.iAtomStart + iAt2f);
    const int ind = __ldg(p.iPair + flat);
    for (int e = lane; e < n1 * n2; e += 32) {
      S[e] = __ldg(over + ind + e);
      P[e] = __ldg(rho + ind + e);
    }
    __syncwarp();
    double *q1 = qq + (size_t)p.mOrb * p.mOrb * iAt1;
    for (int e = lane; e < n1 * n1; e += 32) {
      const int a = e / n1, b = e - a * n1;  // q1(b, a)
      double s = 0.0;
      for (int c = 0; c < n2; ++c) s += P[c + n2 * b] * S[c + n2 * a] + S[c + n2 * b] * P[c + n2 * a];
      atomicAdd(q1 + b + p.mOrb * a, 0.5 * s);
    }
    if (iAt1 != iAt2f) {
      double *q2 = qq + (size_t)p.mOrb * p.mOrb * iAt2f;
      for (int e = lane; e < n2 * n2; e += 32) {
        const int a = e / n2, b = e - a * n2;  // q2(b, a)
        double s = 0.0;
        for (int c = 0; c < n1; ++c) s += P[b + n2 * c] * S[a + n2 * c] + S[b + n2 * c] * P[a + n2 * c];
        atomicAdd(q2 + b + p.mOrb * a, 0.5 * s);
      }
    }
    __syncwarp();
  }
}

int mullikenImpl(dftbp_b200_ctx *ctx, double *qq, const double *over, const double *rho, bool block) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "mulliken: no geometry set");
  DFTBP_CHECK(ctx, qq && over && rho, "mulliken: null argument");
  const size_t nQ = (size_t)ctx->mOrb * (block ? ctx->mOrb : 1) * g.nAtom;
  const double *dOver, *dRho;
  double *dQ;
  DFTBP_TRY(devIn(ctx, kSlotA, over, (size_t)g.sparseSize, &dOver));
  DFTBP_TRY(devIn(ctx, kSlotB, rho, (size_t)g.sparseSize, &dRho));
  DFTBP_TRY(devOut(ctx, kSlotC, qq, nQ, true, &dQ));
  K4Params p;
  p.work = g.workLoad; p.nWork = g.nWorkLoad; p.ld = g.mNeigh + 1; p.mOrb = ctx->mOrb;
  p.iNeighbour = g.iNeighbour; p.iPair = g.iPair; p.img2CentCell = g.img2CentCell;
  p.iAtomStart = g.iAtomStart;
  if (p.nWork > 0) {
    ScopedTimer timer(ctx, kTimerMulliken);
    const int grid = gridFor(ctx, p.nWork, 8, 8);
    if (block) k4_mulliken_block<<<grid, 256, 0, ctx->stream>>>(p, dOver, dRho, dQ);
    else k4_mulliken_orbital<<<grid, 256, 0, ctx->stream>>>(p, dOver, dRho, dQ);
    ctx->launches++;
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_mulliken_orbital(dftbp_b200_ctx *ctx, double *qq, const double *over, const double *rho) {
  return mullikenImpl(ctx, qq, over, rho, false);
}

int dftbp_b200_mulliken_block(dftbp_b200_ctx *ctx, double *qq, const double *over, const double *rho) {
  return mullikenImpl(ctx, qq, over, rho, true);
}

}  // extern "C"

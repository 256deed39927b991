// k4_mulliken.cu -- mullikenPerOrbital / mullikenPerBlock (dftb/populations.F90:120-187, 306-375).
//
// HBM-streaming: 16 bytes read per sparse element (over, rho), output is tiny (mOrb x nAtom).
// A warp walks through a chunk of consecutive sparse blocks (block_desc.cuh: no dependent index
// loads in front of the data).  The element products of a block are staged in the warp's
// shared-memory tile, then lane a forms the column sum for orbital a of iAtom1 and lane 16 + b the
// row sum for orbital b of the folded neighbour (segmented sums inside the warp).  Consecutive
// blocks belong to the same iAtom1 (the sparse array is atom-major), so the column sums stay in a
// register across the chunk and leave the warp only when the atom changes: about n2 global
// reductions per block instead of 2*n1*n2.  qq ACCUMULATES like the reference.
#include "block_desc.cuh"

using namespace dftbp_b200;

namespace {

struct K4Params {
  const BlockDesc *desc;
  long long nWork;
  int mOrb;
};

__global__ void __launch_bounds__(256) k4_mulliken_orbital(const K4Params p, const double *__restrict__ over,
                                                           const double *__restrict__ rho,
                                                           double *__restrict__ qq) {
  __shared__ double tile[8][kMaxOrb * kMaxOrb];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  double *t = tile[wIn];
  double colAcc = 0.0;  // lane a: sum(sqrTmp, dim=1)(a) accumulated over the blocks of curAtom
  int curAtom = -1, curN1 = 0;
  auto flush = [&]() {
    if (curAtom >= 0 && lane < curN1) atomicAdd(qq + lane + p.mOrb * curAtom, colAcc);
    colAcc = 0.0;
  };
  forEachBlock(
      p.desc, p.nWork, [&] { curAtom = -1; colAcc = 0.0; },
      [&](const BlockDesc &b) {
        const int n1 = b.n12 & 0xFF, n2 = b.n12 >> 8, nElem = n1 * n2;
        if (b.iAt1 != curAtom) {
          flush();
          curAtom = b.iAt1;
          curN1 = n1;
        }
        {
          double o[8], r[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // all loads of the block in flight at once
            const int e = lane + 32 * k;
            o[k] = e < nElem ? __ldg(over + b.ind + e) : 0.0;
            r[k] = e < nElem ? __ldg(rho + b.ind + e) : 0.0;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = lane + 32 * k;
            if (e < nElem) t[e] = o[k] * r[k];
          }
        }
        __syncwarp();
        if (lane < n1) {  // sum(sqrTmp, dim=1) -> qq(1:nOrb1, iAtom1)
          double s = 0.0;
          for (int r = 0; r < n2; ++r) s += t[r + n2 * lane];
          colAcc += s;
        }
        if (b.iAt1 != b.iAt2f && lane >= 16 && lane - 16 < n2) {  // sum(sqrTmp, dim=2) -> qq(1:nOrb2, iAtom2f)
          const int r = lane - 16;
          double s = 0.0;
          for (int a = 0; a < n1; ++a) s += t[r + n2 * a];
          atomicAdd(qq + r + p.mOrb * b.iAt2f, s);
        }
        __syncwarp();
      },
      [&] { flush(); curAtom = -1; });
}

// qq(:,:,iAtom1) += 0.5*(rho^T S + S^T rho);  qq(:,:,iAtom2f) += 0.5*(rho S^T + S rho^T)
__global__ void __launch_bounds__(256) k4_mulliken_block(const K4Params p, const double *__restrict__ over,
                                                         const double *__restrict__ rho,
                                                         double *__restrict__ qq) {
  __shared__ double tS[8][kMaxOrb * kMaxOrb], tP[8][kMaxOrb * kMaxOrb];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  double *S = tS[wIn], *P = tP[wIn];
  forEachBlock(
      p.desc, p.nWork, [] {},
      [&](const BlockDesc &b) {
        const int n1 = b.n12 & 0xFF, n2 = b.n12 >> 8;
        for (int e = lane; e < n1 * n2; e += 32) {
          S[e] = __ldg(over + b.ind + e);
          P[e] = __ldg(rho + b.ind + e);
        }
        __syncwarp();
        double *q1 = qq + (size_t)p.mOrb * p.mOrb * b.iAt1;
        for (int e = lane; e < n1 * n1; e += 32) {
          const int a = e / n1, c1 = e - a * n1;  // q1(c1, a)
          double s = 0.0;
          for (int c = 0; c < n2; ++c) s += P[c + n2 * c1] * S[c + n2 * a] + S[c + n2 * c1] * P[c + n2 * a];
          atomicAdd(q1 + c1 + p.mOrb * a, 0.5 * s);
        }
        if (b.iAt1 != b.iAt2f) {
          double *q2 = qq + (size_t)p.mOrb * p.mOrb * b.iAt2f;
          for (int e = lane; e < n2 * n2; e += 32) {
            const int a = e / n2, c2 = e - a * n2;  // q2(c2, a)
            double s = 0.0;
            for (int c = 0; c < n1; ++c) s += P[c2 + n2 * c] * S[a + n2 * c] + S[c2 + n2 * c] * P[a + n2 * c];
            atomicAdd(q2 + c2 + p.mOrb * a, 0.5 * s);
          }
        }
        __syncwarp();
      },
      [] {});
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
  p.desc = g.descLoad; p.nWork = g.nWorkLoad; p.mOrb = ctx->mOrb;
  if (p.nWork > 0) {
    ScopedTimer timer(ctx, kTimerMulliken);
    const int grid = gridFor(ctx, (p.nWork + kBlocksPerChunk - 1) / kBlocksPerChunk, 8, 8);
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

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
#include <cstdlib>

#include "block_desc.cuh"
#include "k4_atom_population.cuh"

using namespace dftbp_b200;

namespace {

struct K4Params {
  const BlockDesc *desc;
  long long nWork;
  int mOrb;
};

// EPL = elements per lane and block (3: blocks up to 96 elements, s/p/d; 8: f shells), BATCH = blocks
// whose loads are issued together (BATCH * EPL * 2 independent loads per lane in flight)
template <int EPL, int BATCH>
__global__ void __launch_bounds__(256) k4_mulliken_orbital(const K4Params p, const double *__restrict__ over,
                                                           const double *__restrict__ rho,
                                                           double *__restrict__ qq) {
  extern __shared__ double smem[];
  constexpr int kTile = 32 * EPL;  // doubles per staged block
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  double *tile = smem + (size_t)wIn * BATCH * kTile;
  double colAcc = 0.0;  // lane a: sum(sqrTmp, dim=1)(a) accumulated over the blocks of curAtom
  int curAtom = -1, curN1 = 0;
  auto flush = [&]() {
    if (curAtom >= 0 && lane < curN1) atomicAdd(qq + lane + p.mOrb * curAtom, colAcc);
    colAcc = 0.0;
  };
  forEachBlockBatch<BATCH>(p.desc, p.nWork, [&](const BlockDesc(&b)[BATCH], int count) {
    // ---- element products of all blocks of the batch -> shared tile (coalesced, all loads in flight)
    {
      double o[BATCH][EPL], r[BATCH][EPL];
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int nElem = q < count ? (b[q].n12 & 0xFF) * (b[q].n12 >> 8) : 0;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const int e = lane + 32 * k;
          o[q][k] = e < nElem ? __ldg(over + b[q].ind + e) : 0.0;
          r[q][k] = e < nElem ? __ldg(rho + b[q].ind + e) : 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < BATCH; ++q)
#pragma unroll
        for (int k = 0; k < EPL; ++k) tile[q * kTile + lane + 32 * k] = o[q][k] * r[q][k];
    }
    __syncwarp();
    // ---- segmented sums, block by block: lane a -> column a, lane 16 + b -> row b
    for (int q = 0; q < count; ++q) {
      const int n1 = b[q].n12 & 0xFF, n2 = b[q].n12 >> 8;
      const double *t = tile + q * kTile;
      if (b[q].iAt1 != curAtom) {
        flush();
        curAtom = b[q].iAt1;
        curN1 = n1;
      }
      if (lane < n1) {  // sum(sqrTmp, dim=1) -> qq(1:nOrb1, iAtom1)
        double s = 0.0;
        for (int rr = 0; rr < n2; ++rr) s += t[rr + n2 * lane];
        colAcc += s;
      }
      if (b[q].iAt1 != b[q].iAt2f && lane >= 16 && lane - 16 < n2) {  // sum(sqrTmp, dim=2) -> qq(1:nOrb2, iAtom2f)
        const int rr = lane - 16;
        double s = 0.0;
        for (int a = 0; a < n1; ++a) s += t[rr + n2 * a];
        atomicAdd(qq + rr + p.mOrb * b[q].iAt2f, s);
      }
    }
    __syncwarp();
  });
  flush();
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
    const int grid = gridFor(ctx, (p.nWork + kBlocksPerChunk - 1) / kBlocksPerChunk, 8, tuneInt("DFTBP_B200_K4_CTAS", 8));
    if (block) {
      k4_mulliken_block<<<grid, 256, 0, ctx->stream>>>(p, dOver, dRho, dQ);
    } else if (ctx->mOrb <= 9) {  // 8 warps x BATCH blocks x 96 doubles
      const char *env = getenv("DFTBP_B200_K4_BATCH");  // tuning switch (tools/gpu_probe.py)
      const int batch = env ? atoi(env) : 2;
      if (batch == 1) k4_mulliken_orbital<3, 1><<<grid, 256, sizeof(double) * 8 * 1 * 96, ctx->stream>>>(p, dOver, dRho, dQ);
      else if (batch == 4) k4_mulliken_orbital<3, 4><<<grid, 256, sizeof(double) * 8 * 4 * 96, ctx->stream>>>(p, dOver, dRho, dQ);
      else k4_mulliken_orbital<3, 2><<<grid, 256, sizeof(double) * 8 * 2 * 96, ctx->stream>>>(p, dOver, dRho, dQ);
    } else {
      const size_t smemBytes = sizeof(double) * 8 * 2 * 256;
      k4_mulliken_orbital<8, 2><<<grid, 256, smemBytes, ctx->stream>>>(p, dOver, dRho, dQ);
    }
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

int dftbp_b200_mulliken_atom(dftbp_b200_ctx *ctx, double *qq, const double *over, const double *rho) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "mulliken_atom: no geometry set");
  DFTBP_CHECK(ctx, qq && over && rho, "mulliken_atom: null argument");
  // qPerOrbital(:,:) = 0; call mullikenPerOrbital(...); qq(:) = qq + sum(qPerOrbital, dim=1)   (:103-110)
  void *tmp;
  const size_t nOrbQ = (size_t)ctx->mOrb * g.nAtom;
  DFTBP_TRY(stagingGet(ctx, kSlotD, sizeof(double) * nOrbQ, &tmp));
  DFTBP_CUDA(ctx, cudaMemsetAsync(tmp, 0, sizeof(double) * nOrbQ, ctx->stream));
  DFTBP_TRY(mullikenImpl(ctx, static_cast<double *>(tmp), over, rho, false));  // device output: no read-back
  double *dQ;
  DFTBP_TRY(devOut(ctx, kSlotE, qq, (size_t)g.nAtom, true, &dQ));
  pop::pop_atom_sum<<<(g.nAtom + 127) / 128, 128, 0, ctx->stream>>>(g.nAtom, ctx->mOrb, static_cast<const double *>(tmp), dQ);
  ctx->launches++;
  return finish(ctx);
}

int dftbp_b200_onsite_population(dftbp_b200_ctx *ctx, const double *rho, double *qq) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "onsite_population: no geometry set");
  DFTBP_CHECK(ctx, qq && rho, "onsite_population: null argument");
  const double *dRho;
  double *dQ;
  DFTBP_TRY(devIn(ctx, kSlotA, rho, (size_t)g.sparseSize, &dRho));
  DFTBP_TRY(devOut(ctx, kSlotC, qq, (size_t)g.nAtom, false, &dQ));
  pop::pop_onsite<<<(g.nAtom + 127) / 128, 128, 0, ctx->stream>>>(g.nAtom, g.mNeigh + 1, g.iPair, g.iAtomStart, dRho, dQ);
  ctx->launches++;
  return finish(ctx);
}

}  // extern "C"

// k2_sparse_dense.cu -- unpackHS / packHS (dftb/sparse2dense.F90:185-321, 820-918, 1024-1096).
//
// HBM-streaming kernels.  A warp walks through a chunk of consecutive sparse blocks (block_desc.cuh: the
// iNeighbour -> img2CentCell -> iAtomStart chain is resolved once per geometry, so no dependent
// index loads sit in front of the data), lane e <-> element e of the column-major block: the sparse
// side is read / written as contiguous runs, the dense side as nOrb2-long column segments.  The
// loads of a block are issued together before the first use.  The Bloch phase
// exp(+-i 2 pi k.cellVec) is one sincospi per block.
//
// unpack: the dense matrix is zero-filled first (the reference does square(:,:) = 0).  Blocks of
// one atom that fold onto the same dense block (several periodic images of one neighbour, or a
// self image on top of the on-site block; small cells only) are detected at set_geometry time;
// only then are the adds atomic, otherwise they are plain stores.
// pack: pure gather + read-modify-write of the sparse array (it ACCUMULATES, like the reference);
// the on-site block is symmetrised / hermitised from its lower triangle (:1084-1088, :906-910).
#include <cstdlib>

#include "block_desc.cuh"

using namespace dftbp_b200;

namespace {

struct K2Params {
  const BlockDesc *desc;
  long long nWork;
  int ldSq;
  const double *cellVec;
  double k0, k1, k2, kWeight;
};

constexpr int kUnroll = 3;  // 96 elements per pass: one pass for anything up to spd-spd (81)

// exp(i * sign * 2 pi k . cellVec)
__device__ __forceinline__ void blochPhase(const K2Params &p, int iVec, double sign, double *re, double *im) {
  const double kr = p.k0 * __ldg(p.cellVec + 3 * iVec) + p.k1 * __ldg(p.cellVec + 3 * iVec + 1) +
                    p.k2 * __ldg(p.cellVec + 3 * iVec + 2);
  sincospi(2.0 * sign * kr, im, re);
}

template <bool CPLX, bool ATOMIC>
__global__ void __launch_bounds__(256) k2_unpack(const K2Params p, const double *__restrict__ orig,
                                                 double *__restrict__ square) {
  const int lane = threadIdx.x & 31;
  forEachBlock(
      p.desc, p.nWork, [] {},
      [&](const BlockDesc &b) {
        double pr = 1.0, pi = 0.0;
        if (CPLX) blochPhase(p, b.iVec, 1.0, &pr, &pi);
        const int n2 = b.n12 >> 8, nElem = (b.n12 & 0xFF) * n2;
        const unsigned M2 = (1048576u + (unsigned)n2 - 1u) / (unsigned)n2;  // e / n2 = e * M2 >> 20 (e < 256)
        for (int e0 = lane; e0 < nElem; e0 += 32 * kUnroll) {
          double v[kUnroll];
#pragma unroll
          for (int k = 0; k < kUnroll; ++k) {
            const int e = e0 + 32 * k;
            v[k] = e < nElem ? __ldg(orig + b.ind + e) : 0.0;
          }
#pragma unroll
          for (int k = 0; k < kUnroll; ++k) {
            const int e = e0 + 32 * k;
            if (e >= nElem) break;
            const int a = (int)(((unsigned)e * M2) >> 20), r = e - a * n2;
            const size_t pos = (size_t)(b.jj + r) + (size_t)p.ldSq * (b.ii + a);
            if (CPLX) {
              if (ATOMIC) {
                atomicAdd(square + 2 * pos, pr * v[k]);
                atomicAdd(square + 2 * pos + 1, pi * v[k]);
              } else {
                reinterpret_cast<double2 *>(square)[pos] = make_double2(pr * v[k], pi * v[k]);
              }
            } else {
              if (ATOMIC) atomicAdd(square + pos, v[k]);
              else square[pos] = v[k];
            }
          }
        }
      },
      [] {});
}

// kPackBatch blocks per pass: the dense gather and the sparse read of all of them are in flight
// before the first store -- one block alone is only six loads per lane (profiles/
// r01d_ncu_all_kernels.txt: 14 long-scoreboard stall cycles per issue), but the registers of a
// large batch cost more occupancy than the batch wins (4 blocks: 100 registers, 2x slower).
template <bool CPLX, int kPackBatch>
__global__ void __launch_bounds__(256) k2_pack(const K2Params p, double *__restrict__ primitive,
                                               const double *__restrict__ square) {
  const int lane = threadIdx.x & 31;
  forEachBlockBatch<kPackBatch>(p.desc, p.nWork, [&](const BlockDesc(&b)[kPackBatch], int count) {
    int maxElem = 0;
#pragma unroll
    for (int q = 0; q < kPackBatch; ++q)
      if (q < count) maxElem = max(maxElem, (b[q].n12 & 0xFF) * (b[q].n12 >> 8));
    for (int e0 = lane; e0 < maxElem; e0 += 32 * kUnroll) {
      double val[kPackBatch][kUnroll], old[kPackBatch][kUnroll];
#pragma unroll
      for (int q = 0; q < kPackBatch; ++q) {
        const bool onsite = b[q].iAt1 == b[q].iAt2f;
        const int n2 = b[q].n12 >> 8, nElem = q < count ? (b[q].n12 & 0xFF) * n2 : 0;
        const unsigned M2 = (1048576u + (unsigned)n2 - 1u) / (unsigned)n2;  // e / n2 = e * M2 >> 20 (e < 256)
        double pr = 1.0, pi = 0.0;
        if (CPLX) blochPhase(p, b[q].iVec, -1.0, &pr, &pi);
        const double *__restrict__ prim = primitive + b[q].ind;
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
          const int e = e0 + 32 * k;
          val[q][k] = 0.0;
          old[q][k] = 0.0;
          if (e < nElem) {
            const int a = (int)(((unsigned)e * M2) >> 20), r = e - a * n2;
            // on-site: element above the diagonal (row < col) is taken from the lower triangle
            const bool mirror = onsite && r < a;
            const size_t pos = mirror ? (size_t)(b[q].jj + a) + (size_t)p.ldSq * (b[q].ii + r)
                                      : (size_t)(b[q].jj + r) + (size_t)p.ldSq * (b[q].ii + a);
            if (CPLX) {
              const double2 z = __ldg(reinterpret_cast<const double2 *>(square) + pos);
              const double zi = mirror ? -z.y : z.y;
              val[q][k] = p.kWeight * (pr * z.x - pi * zi);
            } else {
              val[q][k] = __ldg(square + pos);
            }
            old[q][k] = prim[e];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kPackBatch; ++q) {
        const int nElem = q < count ? (b[q].n12 & 0xFF) * (b[q].n12 >> 8) : 0;
        double *__restrict__ prim = primitive + b[q].ind;
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
          const int e = e0 + 32 * k;
          if (e < nElem) prim[e] = old[q][k] + val[q][k];
        }
      }
    }
  });
}

K2Params makeParams(dftbp_b200_ctx *ctx, int ldSq, const double *k, double kWeight) {
  const Geometry &g = ctx->geo;
  K2Params p;
  p.desc = g.descAll; p.nWork = g.nWorkAll; p.ldSq = ldSq; p.cellVec = g.cellVec;
  p.k0 = k ? k[0] : 0.0; p.k1 = k ? k[1] : 0.0; p.k2 = k ? k[2] : 0.0; p.kWeight = kWeight;
  return p;
}

// one chunk of blocks per warp and pass; enough CTAs to fill the machine, 8 resident per SM
int gridForChunks(const dftbp_b200_ctx *ctx, long long nBlocks) {
  return gridFor(ctx, (nBlocks + kBlocksPerChunk - 1) / kBlocksPerChunk, 8, tuneInt("DFTBP_B200_K2_CTAS", 8));
}

int unpackImpl(dftbp_b200_ctx *ctx, const double *orig, const double *kPoint, void *square, int ld, bool cplx) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "unpackHS: no geometry set");
  DFTBP_CHECK(ctx, orig && square && ld >= g.nOrb, "unpackHS: bad arguments");
  const size_t nSq = (size_t)ld * g.nOrb * (cplx ? 2 : 1);
  const double *dOrig;
  double *dSq;
  DFTBP_TRY(devIn(ctx, kSlotA, orig, (size_t)g.sparseSize, &dOrig));
  DFTBP_TRY(devOut(ctx, kSlotB, (double *)square, nSq, false, &dSq));
  {
  ScopedTimer timer(ctx, kTimerSparseToDense);
  DFTBP_CUDA(ctx, cudaMemsetAsync(dSq, 0, nSq * sizeof(double), ctx->stream));
  const K2Params p = makeParams(ctx, ld, kPoint, 1.0);
  const int grid = gridForChunks(ctx, p.nWork);
  if (cplx) {
    if (g.hasDuplicates) k2_unpack<true, true><<<grid, 256, 0, ctx->stream>>>(p, dOrig, dSq);
    else k2_unpack<true, false><<<grid, 256, 0, ctx->stream>>>(p, dOrig, dSq);
  } else {
    if (g.hasDuplicates) k2_unpack<false, true><<<grid, 256, 0, ctx->stream>>>(p, dOrig, dSq);
    else k2_unpack<false, false><<<grid, 256, 0, ctx->stream>>>(p, dOrig, dSq);
  }
  ctx->launches += 1;  // the scatter kernel (the zero-fill is a cudaMemsetAsync)
  }
  return finish(ctx);
}

int packImpl(dftbp_b200_ctx *ctx, double *primitive, const void *square, int ld, const double *kPoint,
             double kWeight, bool cplx, int atomFirst = 1, int atomLast = 0) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "packHS: no geometry set");
  DFTBP_CHECK(ctx, primitive && square && ld >= g.nOrb, "packHS: bad arguments");
  if (atomLast <= 0) atomLast = g.nAtom;
  DFTBP_CHECK(ctx, atomFirst >= 1 && atomFirst <= atomLast && atomLast <= g.nAtom, "packHS: bad atom range");
  const size_t nSq = (size_t)ld * g.nOrb * (cplx ? 2 : 1);
  double *dPrim;
  const double *dSq;
  DFTBP_TRY(devOut(ctx, kSlotA, primitive, (size_t)g.sparseSize, true, &dPrim));
  DFTBP_TRY(devIn(ctx, kSlotB, (const double *)square, nSq, &dSq));
  {
    ScopedTimer timer(ctx, kTimerDenseToSparse);
    K2Params p = makeParams(ctx, ld, kPoint, kWeight);
    // only the blocks of atoms atomFirst..atomLast (= dense columns of those atoms): the work list is atom-major
    p.desc = g.descAll + ctx->blockStartOfAtom[atomFirst - 1];
    p.nWork = ctx->blockStartOfAtom[atomLast] - ctx->blockStartOfAtom[atomFirst - 1];
    const int grid = gridForChunks(ctx, p.nWork);
    const char *env = getenv("DFTBP_B200_K2_BATCH");  // tuning switch (tools/gpu_probe.py)
    const int batch = env ? atoi(env) : 1;
    if (cplx) {
      if (batch == 1) k2_pack<true, 1><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
      else k2_pack<true, 2><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
    } else {
      if (batch == 1) k2_pack<false, 1><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
      else if (batch == 4) k2_pack<false, 4><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
      else k2_pack<false, 2><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
    }
    ctx->launches++;
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_unpack_hs_real(dftbp_b200_ctx *ctx, const double *orig, double *square, int ld) {
  return unpackImpl(ctx, orig, nullptr, square, ld, false);
}

int dftbp_b200_unpack_hs_cplx(dftbp_b200_ctx *ctx, const double *orig, const double kPoint[3],
                              void *square, int ld) {
  if (ctx && !kPoint) return fail(ctx, DFTBP_B200_EINVAL, "unpack_hs_cplx: kPoint is null");
  return unpackImpl(ctx, orig, kPoint, square, ld, true);
}

int dftbp_b200_pack_hs_real(dftbp_b200_ctx *ctx, double *primitive, const double *square, int ld) {
  return packImpl(ctx, primitive, square, ld, nullptr, 1.0, false);
}

int dftbp_b200_pack_hs_cplx(dftbp_b200_ctx *ctx, double *primitive, const void *square, int ld,
                            const double kPoint[3], double kWeight) {
  if (ctx && !kPoint) return fail(ctx, DFTBP_B200_EINVAL, "pack_hs_cplx: kPoint is null");
  return packImpl(ctx, primitive, square, ld, kPoint, kWeight, true);
}

int dftbp_b200_pack_hs_real_atoms(dftbp_b200_ctx *ctx, double *primitive, const double *square, int ld,
                                  int atomFirst, int atomLast) {
  return packImpl(ctx, primitive, square, ld, nullptr, 1.0, false, atomFirst, atomLast);
}

int dftbp_b200_pack_hs_cplx_atoms(dftbp_b200_ctx *ctx, double *primitive, const void *square, int ld,
                                  const double kPoint[3], double kWeight, int atomFirst, int atomLast) {
  if (ctx && !kPoint) return fail(ctx, DFTBP_B200_EINVAL, "pack_hs_cplx: kPoint is null");
  return packImpl(ctx, primitive, square, ld, kPoint, kWeight, true, atomFirst, atomLast);
}

}  // extern "C"

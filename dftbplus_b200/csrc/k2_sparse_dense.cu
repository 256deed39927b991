// k2_sparse_dense.cu -- unpackHS / packHS (dftb/sparse2dense.F90:185-321, 820-918, 1024-1096).
//
// HBM-streaming kernels, one warp per sparse block, lane e <-> element e of the column-major
// block: the sparse side is read / written as contiguous runs, the dense side as nOrb2-long
// column segments.  The Bloch phase exp(+-i 2 pi k.cellVec) is one sincospi per block.
//
// unpack: the dense matrix is zero-filled first (the reference does square(:,:) = 0).  Blocks of
// one atom that fold onto the same dense block (several periodic images of one neighbour, or a
// self image on top of the on-site block; small cells only) are detected at set_geometry time;
// only then are the adds atomic, otherwise they are plain stores.
// pack: pure gather + read-modify-write of the sparse array (it ACCUMULATES, like the reference);
// the on-site block is symmetrised / hermitised from its lower triangle (:1084-1088, :906-910).
#include "common.cuh"

using namespace dftbp_b200;

namespace {

struct K2Params {
  const int *work;
  long long nWork;
  int ld, ldSq;
  const int *iNeighbour, *iPair, *img2CentCell, *iAtomStart, *iCellVec;
  const double *cellVec;
  double k0, k1, k2, kWeight;
};

struct BlockInfo {
  int iAt1, iAt2f, ii, jj, n1, n2, ind, iVec;
};

__device__ __forceinline__ BlockInfo blockInfo(const K2Params &p, long long w) {
  BlockInfo b;
  const int flat = __ldg(p.work + w);
  b.iAt1 = flat / p.ld;
  const int iAt2 = __ldg(p.iNeighbour + flat) - 1;
  b.iAt2f = __ldg(p.img2CentCell + iAt2) - 1;
  b.ii = __ldg(p.iAtomStart + b.iAt1) - 1;
  b.n1 = __ldg(p.iAtomStart + b.iAt1 + 1) - 1 - b.ii;
  b.jj = __ldg(p.iAtomStart + b.iAt2f) - 1;
  b.n2 = __ldg(p.iAtomStart + b.iAt2f + 1) - 1 - b.jj;
  b.ind = __ldg(p.iPair + flat);
  b.iVec = __ldg(p.iCellVec + iAt2) - 1;
  return b;
}

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
  const long long nWarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < p.nWork; w += nWarps) {
    const BlockInfo b = blockInfo(p, w);
    double pr = 1.0, pi = 0.0;
    if (CPLX) blochPhase(p, b.iVec, 1.0, &pr, &pi);
    const int nElem = b.n1 * b.n2;
    for (int e = lane; e < nElem; e += 32) {
      const int a = e / b.n2, r = e - a * b.n2;
      const double v = __ldg(orig + b.ind + e);
      const size_t pos = (size_t)(b.jj + r) + (size_t)p.ldSq * (b.ii + a);
      if (CPLX) {
        if (ATOMIC) {
          atomicAdd(square + 2 * pos, pr * v);
          atomicAdd(square + 2 * pos + 1, pi * v);
        } else {
          reinterpret_cast<double2 *>(square)[pos] = make_double2(pr * v, pi * v);
        }
      } else {
        if (ATOMIC) atomicAdd(square + pos, v);
        else square[pos] = v;
      }
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) k2_pack(const K2Params p, double *__restrict__ primitive,
                                               const double *__restrict__ square) {
  const int lane = threadIdx.x & 31;
  const long long nWarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < p.nWork; w += nWarps) {
    const BlockInfo b = blockInfo(p, w);
    double pr = 1.0, pi = 0.0;
    if (CPLX) blochPhase(p, b.iVec, -1.0, &pr, &pi);
    const bool onsite = b.iAt1 == b.iAt2f;
    const int nElem = b.n1 * b.n2;
    for (int e = lane; e < nElem; e += 32) {
      const int a = e / b.n2, r = e - a * b.n2;
      // on-site: element above the diagonal (row < col) is taken from the lower triangle
      const bool mirror = onsite && r < a;
      const size_t pos = mirror ? (size_t)(b.jj + a) + (size_t)p.ldSq * (b.ii + r)
                                : (size_t)(b.jj + r) + (size_t)p.ldSq * (b.ii + a);
      double val;
      if (CPLX) {
        const double2 z = __ldg(reinterpret_cast<const double2 *>(square) + pos);
        const double zi = mirror ? -z.y : z.y;
        val = p.kWeight * (pr * z.x - pi * zi);
      } else {
        val = __ldg(square + pos);
      }
      primitive[b.ind + e] += val;
    }
  }
}

K2Params makeParams(dftbp_b200_ctx *ctx, int ldSq, const double *k, double kWeight) {
  const Geometry &g = ctx->geo;
  K2Params p;
  p.work = g.workAll; p.nWork = g.nWorkAll; p.ld = g.mNeigh + 1; p.ldSq = ldSq;
  p.iNeighbour = g.iNeighbour; p.iPair = g.iPair; p.img2CentCell = g.img2CentCell;
  p.iAtomStart = g.iAtomStart; p.iCellVec = g.iCellVec; p.cellVec = g.cellVec;
  p.k0 = k ? k[0] : 0.0; p.k1 = k ? k[1] : 0.0; p.k2 = k ? k[2] : 0.0; p.kWeight = kWeight;
  return p;
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
  const int grid = gridFor(ctx, p.nWork, 8, 8);
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
             double kWeight, bool cplx) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "packHS: no geometry set");
  DFTBP_CHECK(ctx, primitive && square && ld >= g.nOrb, "packHS: bad arguments");
  const size_t nSq = (size_t)ld * g.nOrb * (cplx ? 2 : 1);
  double *dPrim;
  const double *dSq;
  DFTBP_TRY(devOut(ctx, kSlotA, primitive, (size_t)g.sparseSize, true, &dPrim));
  DFTBP_TRY(devIn(ctx, kSlotB, (const double *)square, nSq, &dSq));
  {
    ScopedTimer timer(ctx, kTimerDenseToSparse);
    const K2Params p = makeParams(ctx, ld, kPoint, kWeight);
    const int grid = gridFor(ctx, p.nWork, 8, 8);
    if (cplx) k2_pack<true><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
    else k2_pack<false><<<grid, 256, 0, ctx->stream>>>(p, dPrim, dSq);
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

}  // extern "C"

// block_desc.cuh -- flat per-block descriptors for the streaming kernels (K2 unpack / pack, K4
// Mulliken, K5 force weights).
//
// The reference reaches a sparse block through a chain of look-ups: iNeighbour(iNeigh, iAtom) ->
// img2CentCell -> iAtomStart / iCellVec, plus iPair(iNeigh, iAtom) (dftb/sparse2dense.F90:236-258).
// On a GPU that is four dependent global loads in front of every block: the streaming kernels were
// latency bound on it (profiles/r01c_ncu_k4_mulliken_orbital.txt: 36 % of the HBM roofline).  The
// chain depends on the geometry only, so set_geometry resolves it once on the host into one
// 32-byte record per block, in work-list order (= the order of the sparse array).  A warp then
// loads 32 records with one coalesced request and walks through them with shuffles.
#pragma once
#include "common.cuh"

namespace dftbp_b200 {

struct __align__(16) BlockDesc {
  int ind;    // iPair(iNeigh, iAtom): 0-based offset of the block in every sparse array
  int ii;     // iAtomStart(iAtom1) - 1: first dense column
  int jj;     // iAtomStart(iAtom2f) - 1: first dense row
  int n12;    // nOrb1 | nOrb2 << 8
  int iAt1;   // 0-based central atom
  int iAt2f;  // 0-based folded neighbour, img2CentCell(iNeighbour(iNeigh, iAtom)) - 1
  int iVec;   // iCellVec(iAtom2) - 1
  int iNeigh; // 0 = on-site block
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc must be two 16-byte words");

// Sparse elements [*eLo, *eHi) covered by the entries [wLo, wHi) of the load work list (a share of the
// pairs, set_pair_partition).  Host-built geometries keep the offsets in ctx->loadInd; a geometry whose
// descriptors were built on the device (adopt_neighbour_list) reads the two offsets from descLoad, once
// per (geometry, share).
inline int loadElementRange(dftbp_b200_ctx *ctx, long long wLo, long long wHi, int *eLo, int *eHi) {
  const Geometry &g = ctx->geo;
  if ((long long)ctx->loadInd.size() == g.nWorkLoad + 1) {
    *eLo = ctx->loadInd[wLo];
    *eHi = ctx->loadInd[wHi];
    return 0;
  }
  if (ctx->rangeLo != wLo || ctx->rangeHi != wHi) {
    auto offsetOf = [&](long long w, int *out) -> int {
      if (w <= 0) { *out = 0; return 0; }
      if (w >= g.nWorkLoad) { *out = g.sparseSize; return 0; }
      DFTBP_CUDA(ctx, cudaMemcpyAsync(out, &g.descLoad[w].ind, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      return 0;
    };
    DFTBP_TRY(offsetOf(wLo, &ctx->rangeELo));
    DFTBP_TRY(offsetOf(wHi, &ctx->rangeEHi));
    ctx->rangeLo = wLo;
    ctx->rangeHi = wHi;
  }
  *eLo = ctx->rangeELo;
  *eHi = ctx->rangeEHi;
  return 0;
}

#ifdef __CUDACC__
// f(const BlockDesc &) is called by all 32 lanes of a warp for every block of the list, in list
// order inside chunks of kBlocksPerChunk consecutive blocks; chunks go round-robin over the warps
// of the grid.  chunkBegin() / chunkEnd() bracket a chunk (register accumulators carried across its
// blocks).  Eight blocks per chunk: one 256-byte descriptor request per chunk, and still about two
// chunks per resident warp at 10^5 blocks (32 per chunk would leave half of the warp slots empty).
constexpr int kBlocksPerChunk = 8;

template <class Begin, class Body, class End>
__device__ __forceinline__ void forEachBlock(const BlockDesc *__restrict__ desc, long long nBlocks, Begin &&chunkBegin,
                                             Body &&f, End &&chunkEnd) {
  const int lane = threadIdx.x & 31;
  const long long nWarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nChunks = (nBlocks + kBlocksPerChunk - 1) / kBlocksPerChunk;
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nChunks; c += nWarps) {
    const long long i = c * kBlocksPerChunk + lane;
    int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
    if (lane < kBlocksPerChunk && i < nBlocks) {
      const int4 *src = reinterpret_cast<const int4 *>(desc + i);
      lo = __ldg(src);
      hi = __ldg(src + 1);
    }
    const int cnt = (int)min((long long)kBlocksPerChunk, nBlocks - c * kBlocksPerChunk);
    chunkBegin();
    for (int j = 0; j < cnt; ++j) {
      BlockDesc b;
      b.ind = __shfl_sync(0xffffffffu, lo.x, j);
      b.ii = __shfl_sync(0xffffffffu, lo.y, j);
      b.jj = __shfl_sync(0xffffffffu, lo.z, j);
      b.n12 = __shfl_sync(0xffffffffu, lo.w, j);
      b.iAt1 = __shfl_sync(0xffffffffu, hi.x, j);
      b.iAt2f = __shfl_sync(0xffffffffu, hi.y, j);
      b.iVec = __shfl_sync(0xffffffffu, hi.z, j);
      b.iNeigh = __shfl_sync(0xffffffffu, hi.w, j);
      f(b);
    }
    chunkEnd();
  }
}

// Same walk, kBatch blocks at a time: f(const BlockDesc (&b)[kBatch], int count) gets count <= kBatch
// consecutive blocks of a chunk, so that a streaming kernel can issue the loads of all of them
// before the first use (memory-level parallelism: one block is only 2-3 loads per lane).
template <int kBatch, class Body>
__device__ __forceinline__ void forEachBlockBatch(const BlockDesc *__restrict__ desc, long long nBlocks, Body &&f) {
  static_assert(kBlocksPerChunk % kBatch == 0, "batch must divide the chunk");
  const int lane = threadIdx.x & 31;
  const long long nWarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nChunks = (nBlocks + kBlocksPerChunk - 1) / kBlocksPerChunk;
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nChunks; c += nWarps) {
    const long long i = c * kBlocksPerChunk + lane;
    int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
    if (lane < kBlocksPerChunk && i < nBlocks) {
      const int4 *src = reinterpret_cast<const int4 *>(desc + i);
      lo = __ldg(src);
      hi = __ldg(src + 1);
    }
    const int cnt = (int)min((long long)kBlocksPerChunk, nBlocks - c * kBlocksPerChunk);
    for (int j0 = 0; j0 < cnt; j0 += kBatch) {
      BlockDesc b[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int j = min(j0 + k, cnt - 1);
        b[k].ind = __shfl_sync(0xffffffffu, lo.x, j);
        b[k].ii = __shfl_sync(0xffffffffu, lo.y, j);
        b[k].jj = __shfl_sync(0xffffffffu, lo.z, j);
        b[k].n12 = __shfl_sync(0xffffffffu, lo.w, j);
        b[k].iAt1 = __shfl_sync(0xffffffffu, hi.x, j);
        b[k].iAt2f = __shfl_sync(0xffffffffu, hi.y, j);
        b[k].iVec = __shfl_sync(0xffffffffu, hi.z, j);
        b[k].iNeigh = __shfl_sync(0xffffffffu, hi.w, j);
      }
      f(b, min(kBatch, cnt - j0));
    }
  }
}
// forEachBlockBatch with one more int per block from a parallel array (aux[i] belongs to desc[i]):
// f(const BlockDesc (&b)[kBatch], const int (&aux)[kBatch], int count)
template <int kBatch, class Body>
__device__ __forceinline__ void forEachBlockBatchAux(const BlockDesc *__restrict__ desc, const int *__restrict__ aux,
                                                     long long nBlocks, Body &&f) {
  static_assert(kBlocksPerChunk % kBatch == 0, "batch must divide the chunk");
  const int lane = threadIdx.x & 31;
  const long long nWarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nChunks = (nBlocks + kBlocksPerChunk - 1) / kBlocksPerChunk;
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nChunks; c += nWarps) {
    const long long i = c * kBlocksPerChunk + lane;
    int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
    int ax = 0;
    if (lane < kBlocksPerChunk && i < nBlocks) {
      const int4 *src = reinterpret_cast<const int4 *>(desc + i);
      lo = __ldg(src);
      hi = __ldg(src + 1);
      ax = __ldg(aux + i);
    }
    const int cnt = (int)min((long long)kBlocksPerChunk, nBlocks - c * kBlocksPerChunk);
    for (int j0 = 0; j0 < cnt; j0 += kBatch) {
      BlockDesc b[kBatch];
      int a[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int j = min(j0 + k, cnt - 1);
        b[k].ind = __shfl_sync(0xffffffffu, lo.x, j);
        b[k].ii = __shfl_sync(0xffffffffu, lo.y, j);
        b[k].jj = __shfl_sync(0xffffffffu, lo.z, j);
        b[k].n12 = __shfl_sync(0xffffffffu, lo.w, j);
        b[k].iAt1 = __shfl_sync(0xffffffffu, hi.x, j);
        b[k].iAt2f = __shfl_sync(0xffffffffu, hi.y, j);
        b[k].iVec = __shfl_sync(0xffffffffu, hi.z, j);
        b[k].iNeigh = __shfl_sync(0xffffffffu, hi.w, j);
        a[k] = __shfl_sync(0xffffffffu, ax, j);
      }
      f(b, a, min(kBatch, cnt - j0));
    }
  }
}
#endif


}  // namespace dftbp_b200

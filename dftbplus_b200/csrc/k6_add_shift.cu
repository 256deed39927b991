// k6_add_shift.cu -- addShift_atom / addShift_block (dftb/shift.F90:41-127, 132-228): the SCC
// contribution to the sparse Hamiltonian, every SCC iteration right before unpackHS
// (getSccHamiltonian, dftbplus/main.F90:1449):
//
//   ham(block, iSpin) += 1/2 * over(block) * (shift(iAt1, iSpin) + shift(iAt2f, iSpin))      (atom)
//   ham(block, iSpin) += 1/2 * (S . V(:,:,iAt1,iSpin) + V(:,:,iAt2f,iSpin) . S)               (block)
//
// HBM-streaming, same shape as the Mulliken kernel: a warp walks through chunks of consecutive
// sparse blocks (block_desc.cuh), lane <-> element, 8 (1 + 2 nSpin) algorithmic bytes per element
// (over read once, ham read-modify-write per spin).  Block shifts of plain SCC / collinear spin are
// diagonal (hamiltonian.F90 expands atom and shell potentials onto the block diagonal,
// shift.F90:232-306): a device-side flag selects the element-wise form
// 1/2 S(b,a) (V1(a,a) + V2(b,b)) or the per-block matrix products, without a host round trip.
// isInputZero: the reference zeroes ham and loops over this rank's workload-balanced atoms
// (distributeRangeWithWorkload; assembleChunks sums the ranks); otherwise every rank does all atoms.
#include "block_desc.cuh"
#include "flat_stream.cuh"

using namespace dftbp_b200;

namespace {

struct K6Params {
  const BlockDesc *desc;
  long long nWork;
  int mOrb, nAtom, nSpin, sparseSize;
  const double *over, *shift;
  const int *shiftFull;  // device flag (block variant): 1 = some block has off-diagonal elements
  const double *diag;    // (mOrb, nAtom, nSpin): diagonals of the block shifts, compact (k6_shift_structure)
  double *ham;
};

// flag = 1 if any off-diagonal element of any shift block is non-zero (one warp per block); the
// diagonals go to diag(mOrb, nBlocks), compact: the element-wise kernels read V(a,a) of the two atoms
// of every sparse element, and in the blocks every value sits in a sector of its own
__global__ void __launch_bounds__(256) k6_shift_structure(const double *__restrict__ shift, int mOrb, int nBlocks,
                                                          int *flag, double *__restrict__ diag) {
  const int lane = threadIdx.x & 31, mo2 = mOrb * mOrb;
  bool any = false;
  for (int blk = blockIdx.x * 8 + (threadIdx.x >> 5); blk < nBlocks; blk += gridDim.x * 8) {
    const double *__restrict__ v = shift + (size_t)mo2 * blk;
    for (int e = lane; e < mo2; e += 32) {
      const int a = e / mOrb, b = e - a * mOrb;
      const double x = __ldg(v + e);
      if (a != b && x != 0.0) any = true;
      if (a == b) diag[(size_t)mOrb * blk + a] = x;
    }
  }
  if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr(flag, 1);
}

// MODE 0: atom shifts shift(nAtom, nSpin); 1: diagonal block shifts; 2: full block shifts.
// ZERO: isInputZero -- ham is written, not accumulated: no read of ham and (on one rank, where the
// work list covers every block) no zero fill either.
template <int MODE, int NSPIN, bool ZERO>
__device__ __forceinline__ void addShiftBlocks(const K6Params &p) {
  const int lane = threadIdx.x & 31;
  const int mo2 = p.mOrb * p.mOrb, dstep = p.mOrb + 1;
  forEachBlock(
      p.desc, p.nWork, [] {},
      [&](const BlockDesc &blk) {
        const int n1 = blk.n12 & 0xFF, n2 = blk.n12 >> 8, nElem = n1 * n2, ind = blk.ind;
        const unsigned M2 = (1048576u + (unsigned)n2 - 1u) / (unsigned)n2;  // e / n2 = e * M2 >> 20 (e < 256)
        const double *__restrict__ S = p.over + ind;
        for (int e0 = lane; e0 < nElem; e0 += 96) {
          double s[3], h[3][NSPIN], x[3][NSPIN];
          int e[3];
          bool on[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {  // three elements per lane and pass, all loads independent
            e[k] = e0 + 32 * k;
            on[k] = e[k] < nElem;
            if (!on[k]) e[k] = 0;
            const int a = (int)(((unsigned)e[k] * M2) >> 20), b = e[k] - a * n2;  // element (row b, col a)
            s[k] = __ldg(S + e[k]);
#pragma unroll
            for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
              h[k][iSpin] = ZERO ? 0.0 : p.ham[(size_t)p.sparseSize * iSpin + ind + e[k]];
              if (MODE == 0) {
                x[k][iSpin] = s[k] * 0.5 * (__ldg(p.shift + blk.iAt1 + (size_t)p.nAtom * iSpin) +
                                            __ldg(p.shift + blk.iAt2f + (size_t)p.nAtom * iSpin));
              } else {
                const double *__restrict__ V1 = p.shift + (size_t)mo2 * (blk.iAt1 + (size_t)p.nAtom * iSpin);
                const double *__restrict__ V2 = p.shift + (size_t)mo2 * (blk.iAt2f + (size_t)p.nAtom * iSpin);
                if (MODE == 1) {
                  x[k][iSpin] = 0.5 * (s[k] * (__ldg(V1 + a * dstep) + __ldg(V2 + b * dstep)));
                } else {
                  // (S V1)(b,a) = sum_c S(b,c) V1(c,a);  (V2 S)(b,a) = sum_c V2(b,c) S(c,a)
                  double u = 0.0, v = 0.0;
                  for (int c = 0; c < n1; ++c) u = fma(__ldg(S + b + n2 * c), __ldg(V1 + c + p.mOrb * a), u);
                  for (int c = 0; c < n2; ++c) v = fma(__ldg(V2 + b + p.mOrb * c), __ldg(S + c + n2 * a), v);
                  x[k][iSpin] = 0.5 * (u + v);
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int iSpin = 0; iSpin < NSPIN; ++iSpin)
              if (on[k]) p.ham[(size_t)p.sparseSize * iSpin + ind + e[k]] = h[k][iSpin] + x[k][iSpin];
        }
      },
      [] {});
}

template <int MODE, int NSPIN, bool ZERO>
__global__ void __launch_bounds__(256) k6_add_shift(const K6Params p) {
  if (MODE == 1 && __ldg(p.shiftFull) != 0) return;  // the MODE 2 launch handles it
  if (MODE == 2 && __ldg(p.shiftFull) == 0) return;
  addShiftBlocks<MODE, NSPIN, ZERO>(p);
}

// Atom shifts and diagonal block shifts through the span pipeline of flat_stream.cuh: `over` (and,
// when accumulating, the NSPIN channels of `ham`) arrive by TMA bulk copies; lanes on consecutive
// elements of the sparse arrays whatever block they belong to.
template <int MODE, int NSPIN, bool ZERO>
__global__ void __launch_bounds__(32 * flat::warpsFor(ZERO ? 1 : 1 + NSPIN), 1)
    k6_add_shift_flat(const K6Params p, const int *__restrict__ groupFirst, int fullElsewhere, int *hint) {
  // what this call found, for the launches of the next one (pinned host word, mapped: one posted store)
  if (MODE == 1 && blockIdx.x == 0 && threadIdx.x == 0)
    *reinterpret_cast<volatile int *>(hint) = __ldg(p.shiftFull) != 0 ? 1 : 0;
  if (MODE == 1 && __ldg(p.shiftFull) != 0) {  // full shift blocks: per-block matrix products, warp per block
    if (!fullElsewhere) addShiftBlocks<2, NSPIN, ZERO>(p);  // (else: the dedicated kernel has been queued too)
    return;
  }
  constexpr int NIN = ZERO ? 1 : 1 + NSPIN, WARPS = flat::warpsFor(NIN);
  extern __shared__ __align__(128) unsigned char flatRaw[];
  flat::Smem<NIN, WARPS> &sm = *reinterpret_cast<flat::Smem<NIN, WARPS> *>(flatRaw);
  __shared__ unsigned m2[kMaxOrb + 1];
  flat::initDivTable(m2);
  const double *src[NIN];
  src[0] = p.over;
  if constexpr (!ZERO) {
#pragma unroll
    for (int iSpin = 0; iSpin < NSPIN; ++iSpin) src[1 + iSpin] = p.ham + (size_t)p.sparseSize * iSpin;
  }
  const int mOrb = p.mOrb;
  const int spinStride = (MODE == 0 ? 1 : mOrb) * p.nAtom;
  flat::streamSpans<NIN, WARPS>(
      sm, src, p.sparseSize, p.desc, (int)p.nWork, groupFirst, 0, 0x7fffffff, [&](const flat::Elem<NIN>(&el)[flat::kBatch]) {
        double v1[flat::kBatch][NSPIN], v2[flat::kBatch][NSPIN];
#pragma unroll
        for (int u = 0; u < flat::kBatch; ++u) {
          const double *__restrict__ V1, *__restrict__ V2;
          if (MODE == 0) {
            V1 = p.shift + el[u].hi.x;
            V2 = p.shift + el[u].hi.y;
          } else {
            const int n2 = el[u].lo.w >> 8, e = el[u].e;
            const int a = (int)(((unsigned)e * m2[n2]) >> 20), b = e - a * n2;  // element (row b, col a)
            V1 = p.diag + (mOrb * el[u].hi.x + a);
            V2 = p.diag + (mOrb * el[u].hi.y + b);
          }
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
            v1[u][iSpin] = __ldg(V1 + (size_t)spinStride * iSpin);
            v2[u][iSpin] = __ldg(V2 + (size_t)spinStride * iSpin);
          }
        }
#pragma unroll
        for (int u = 0; u < flat::kBatch; ++u) {
          const double s = el[u].v[0];
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
            const double xs = MODE == 0 ? s * 0.5 * (v1[u][iSpin] + v2[u][iSpin]) : 0.5 * (s * (v1[u][iSpin] + v2[u][iSpin]));
            double h = 0.0;
            if constexpr (!ZERO) h = el[u].v[1 + iSpin];
            if (el[u].ok) p.ham[(size_t)p.sparseSize * iSpin + el[u].x] = h + xs;
          }
        }
      });
}

template <int MODE, int NSPIN, bool ZERO>
int launchFlatT(dftbp_b200_ctx *ctx, const K6Params &p, const int *groupFirst, int fullElsewhere) {
  constexpr int NIN = ZERO ? 1 : 1 + NSPIN, WARPS = flat::warpsFor(NIN);
  const size_t smemBytes = sizeof(flat::Smem<NIN, WARPS>);
  static LaunchInfo info;
  DFTBP_TRY(launchInfo(ctx, info, k6_add_shift_flat<MODE, NSPIN, ZERO>, 32 * WARPS, smemBytes, nullptr));
  const long long nSpans = ((long long)p.sparseSize + flat::kSpan - 1) / flat::kSpan;
  const int grid = (int)std::max<long long>(1, std::min<long long>((nSpans + WARPS - 1) / WARPS, ctx->smCount));
  k6_add_shift_flat<MODE, NSPIN, ZERO><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p, groupFirst, fullElsewhere, ctx->hShiftHint);
  return 0;
}

template <int MODE>
int launchFlat(dftbp_b200_ctx *ctx, const K6Params &p, bool zero, const int *groupFirst, int fullElsewhere = 0) {
  const int fe = fullElsewhere;
  if (p.nSpin == 1) return zero ? launchFlatT<MODE, 1, true>(ctx, p, groupFirst, fe) : launchFlatT<MODE, 1, false>(ctx, p, groupFirst, fe);
  if (p.nSpin == 2) return zero ? launchFlatT<MODE, 2, true>(ctx, p, groupFirst, fe) : launchFlatT<MODE, 2, false>(ctx, p, groupFirst, fe);
  return zero ? launchFlatT<MODE, 4, true>(ctx, p, groupFirst, fe) : launchFlatT<MODE, 4, false>(ctx, p, groupFirst, fe);
}

template <int MODE, bool ZERO>
void launchModeZ(const K6Params &p, int grid, cudaStream_t st) {
  if (p.nSpin == 1) k6_add_shift<MODE, 1, ZERO><<<grid, 256, 0, st>>>(p);
  else if (p.nSpin == 2) k6_add_shift<MODE, 2, ZERO><<<grid, 256, 0, st>>>(p);
  else k6_add_shift<MODE, 4, ZERO><<<grid, 256, 0, st>>>(p);
}

template <int MODE>
void launchMode(const K6Params &p, bool zero, int grid, cudaStream_t st) {
  if (zero) launchModeZ<MODE, true>(p, grid, st);
  else launchModeZ<MODE, false>(p, grid, st);
}

int addShiftImpl(dftbp_b200_ctx *ctx, double *ham, const double *over, const double *shift, int nSpin,
                 int isInputZero, bool block) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "add_shift: no geometry set");
  DFTBP_CHECK(ctx, ham && over && shift, "add_shift: null argument");
  DFTBP_CHECK(ctx, nSpin == 1 || nSpin == 2 || nSpin == 4, "add_shift: nSpin must be 1, 2 or 4");
  K6Params p;
  p.mOrb = ctx->mOrb; p.nAtom = g.nAtom; p.nSpin = nSpin; p.sparseSize = g.sparseSize;
  const size_t nShift = (block ? (size_t)ctx->mOrb * ctx->mOrb : 1) * g.nAtom * nSpin;
  DFTBP_TRY(devOut(ctx, kSlotA, ham, (size_t)g.sparseSize * nSpin, !isInputZero, &p.ham));
  DFTBP_TRY(devIn(ctx, kSlotB, over, (size_t)g.sparseSize, &p.over));
  DFTBP_TRY(devIn(ctx, kSlotC, shift, nShift, &p.shift));
  const int *groupFirst;
  if (isInputZero) {
    // several ranks: the blocks of the other ranks' atoms stay zero (assembleChunks sums the arrays)
    if (ctx->nRanks > 1 || !g.coversAll)
      DFTBP_CUDA(ctx, cudaMemsetAsync(p.ham, 0, sizeof(double) * (size_t)g.sparseSize * nSpin, ctx->stream));
    p.desc = g.descLoad; p.nWork = g.nWorkLoad;
    groupFirst = g.groupFirstLoad;
  } else {
    p.desc = g.descAll; p.nWork = g.nWorkAll;
    groupFirst = g.groupFirstAll;
  }
  // sorted lists of not too small blocks: the TMA span pipeline (flat_stream.cuh)
  const bool useFlat = groupFirst != nullptr && nShift < ((size_t)1 << 31) && tuneInt("DFTBP_B200_K6_FLAT", 1) != 0;
  p.shiftFull = ctx->dInfo + 1;
  p.diag = nullptr;
  if (p.nWork > 0) {
    ScopedTimer timer(ctx, kTimerAddShift);
    const int grid = gridFor(ctx, (p.nWork + kBlocksPerChunk - 1) / kBlocksPerChunk, 8, tuneInt("DFTBP_B200_K6_CTAS", 4));
    if (!block) {
      if (useFlat) DFTBP_TRY(launchFlat<0>(ctx, p, isInputZero != 0, groupFirst));
      else launchMode<0>(p, isInputZero != 0, grid, ctx->stream);
      ctx->launches++;
    } else {
      int *flag = ctx->dInfo + 1;
      DFTBP_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
      double *dDiag;
      DFTBP_TRY(shiftDiagBuffer(ctx, (long long)ctx->mOrb * g.nAtom * nSpin, &dDiag));
      p.diag = dDiag;
      k6_shift_structure<<<std::min(ctx->smCount * 8, (g.nAtom * nSpin + 7) / 8), 256, 0, ctx->stream>>>(
          p.shift, ctx->mOrb, g.nAtom * nSpin, flag, dDiag);
      // both forms are launched; the one that does not match the device flag returns at once
      if (useFlat) {
        // one launch: the kernel takes the full-block path itself when the device flag says so
        const bool fullLastTime = *ctx->hShiftHint == 1;
        DFTBP_TRY(launchFlat<1>(ctx, p, isInputZero != 0, groupFirst, fullLastTime ? 1 : 0));
        if (fullLastTime) {
          launchMode<2>(p, isInputZero != 0, grid, ctx->stream);
          ctx->launches++;
        }
        ctx->launches += 2;
      } else {
        // both forms are launched; the one that does not match the device flag returns at once
        launchMode<1>(p, isInputZero != 0, grid, ctx->stream);
        launchMode<2>(p, isInputZero != 0, grid, ctx->stream);
        ctx->launches += 3;
      }
    }
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_add_shift_atom(dftbp_b200_ctx *ctx, double *ham, const double *over, const double *shift, int nSpin,
                              int isInputZero) {
  return addShiftImpl(ctx, ham, over, shift, nSpin, isInputZero, false);
}

int dftbp_b200_add_shift_block(dftbp_b200_ctx *ctx, double *ham, const double *over, const double *shift, int nSpin,
                               int isInputZero) {
  return addShiftImpl(ctx, ham, over, shift, nSpin, isInputZero, true);
}

}  // extern "C"

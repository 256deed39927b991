// k1_pair_build.cu -- buildH0 / buildS (dftb/nonscc.F90:66-181, 364-424) as one sm_100a kernel.
//
// Lane-per-pair: every thread evaluates the diatomic block of one neighbour pair with the
// straight-line evaluator of sk_lane_eval.cuh; the 32 pairs of a warp come from the device-built
// schedule of pair_sort.cu (one species pair, neighbouring distances).  A thread's block is
// scattered over 648-byte strides in the sparse array, so the warp transposes through shared
// memory: each lane writes its elements into its own row of a [32][pitch] staging tile (odd
// pitch: conflict free), then the warp streams the 32 blocks out one after the other with lanes
// running over the contiguous elements -- every store instruction covers one 256-byte run.  H0
// and S of a pair share a row (H at [0, nElem), S at [nElem, 2 nElem)) so small blocks (sp-sp: 16
// elements) still fill the warp.  On-site blocks (iNeigh = 0; diagonal = on-site energies or 1,
// nonscc.F90:109-116 / :166-173) are written by a second grid-stride loop, one warp per atom.
// Every element of this rank's segment is written exactly once: no memset, no atomics.
//
// Roofline: HBM-write bound for d-block systems, 16*n1*n2 + 8 algorithmic bytes per pair
// (H0 + S fused); DESIGN.md section "K1".
#include "sk_lane_eval.cuh"

using namespace dftbp_b200;

namespace {

struct K1Params {
  const SpPairProg *progs;
  const int *items;   // sorted pair schedule (pair_sort.cu)
  const int *total;   // its padded length
  int ld, nSpecies, mOrb, pitch;
  int firstAtom, lastAtom;  // on-site blocks of this rank
  const double *coords;
  const int *species;
  const int *iNeighbour;
  const int *iPair;
  const double *selfE;  // (mOrb, nSpecies)
  double *ham;
  double *over;
};

// TMA bulk copy shared -> global (cp.async.bulk, SASS UBLKCP): both addresses 16-byte aligned,
// size a multiple of 16.
__device__ __forceinline__ void bulkStore(double *gdst, const double *ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

// Stage row of one lane and one matrix: element e of the block lives at row[e + par], par = the
// parity (in doubles) of the block's global address, so that shared and global element addresses
// agree modulo 16 and the 16-byte aligned interior of the block can leave as ONE bulk copy; the
// at most one leading and one trailing element go out as plain stores.
__device__ __forceinline__ void flushRow(double *gdst, const double *row, int par, int nElem) {
  const int nBulk = (nElem - par) & ~1;
  if (nBulk > 0) bulkStore(gdst + par, row + 2 * par, 8u * (unsigned)nBulk);
  if (par) gdst[0] = row[1];
  if (par + nBulk < nElem) gdst[nElem - 1] = row[nElem - 1 + par];
}

template <int LMAX, bool DO_H, bool DO_S>
__global__ void __launch_bounds__(256) k1_pair_build(const K1Params p) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  // this lane's staging rows: H at rowH, S at rowS (each 16-byte aligned, p.pitch doubles apart)
  double *rowH = smem + ((size_t)wIn * 32 + lane) * ((DO_H && DO_S) ? 2 : 1) * p.pitch;
  double *rowS = rowH + ((DO_H && DO_S) ? p.pitch : 0);
  const int nChunks = __ldg(p.total) >> 5;
  for (int chunk = blockIdx.x * nW + wIn; chunk < nChunks; chunk += gridDim.x * nW) {
    int flat = __ldg(p.items + 32 * chunk + lane);
    const bool valid = flat >= 0;
    const int flat0 = __shfl_sync(0xffffffffu, flat, 0);  // lane 0 of a chunk is never padding
    if (!valid) flat = flat0;
    const int iAt1 = flat / p.ld;
    const int iAt2 = __ldg(p.iNeighbour + flat) - 1;
    const int sp1 = __shfl_sync(0xffffffffu, __ldg(p.species + iAt1), 0) - 1;
    const int sp2 = __shfl_sync(0xffffffffu, __ldg(p.species + iAt2), 0) - 1;
    const int ind = __ldg(p.iPair + flat);
    const SpPairProg pg = p.progs[sp1 + p.nSpecies * sp2];
    const int n2 = pg.nOrb2, nElem = pg.nOrb1 * n2;
    double *gH = DO_H ? p.ham + ind : nullptr;
    double *gS = DO_S ? p.over + ind : nullptr;
    const int parH = (int)((reinterpret_cast<size_t>(gH) >> 3) & 1);
    const int parS = (int)((reinterpret_cast<size_t>(gS) >> 3) & 1);
    // vect = coords(:,iAt2) - coords(:,iAt1); dist = sqrt(sum(vect**2)); vect = vect/dist
    const double vx = __ldg(p.coords + 3 * iAt2) - __ldg(p.coords + 3 * iAt1);
    const double vy = __ldg(p.coords + 3 * iAt2 + 1) - __ldg(p.coords + 3 * iAt1 + 1);
    const double vz = __ldg(p.coords + 3 * iAt2 + 2) - __ldg(p.coords + 3 * iAt1 + 2);
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const double rinv = 1.0 / dist;
    // the previous chunk's bulk copies must have read their rows before they are overwritten
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    double *wH = rowH + parH, *wS = rowS + parS;
    const double ux = vx * rinv, uy = vy * rinv, uz = vz * rinv;
    auto put = [&](int e, double h, double s) {
      if (DO_H) wH[e] = h;
      if (DO_S) wS[e] = s;
    };
    // canonical species pairs (s / sp / spd): fully unrolled evaluator; anything else: generic loops
    const int canon = (pg.canon1 >= 0 && pg.canon2 >= 0 && pg.canon1 <= 2 && pg.canon2 <= 2)
                          ? pg.canon1 * 3 + pg.canon2 : -1;
    switch (canon) {
      case 0: skLaneEvalCanon<DO_H, DO_S, 0, 0>(pg, dist, ux, uy, uz, put); break;
      case 1: if constexpr (LMAX >= 1) skLaneEvalCanon<DO_H, DO_S, 0, 1>(pg, dist, ux, uy, uz, put); break;
      case 2: if constexpr (LMAX >= 2) skLaneEvalCanon<DO_H, DO_S, 0, 2>(pg, dist, ux, uy, uz, put); break;
      case 3: if constexpr (LMAX >= 1) skLaneEvalCanon<DO_H, DO_S, 1, 0>(pg, dist, ux, uy, uz, put); break;
      case 4: if constexpr (LMAX >= 1) skLaneEvalCanon<DO_H, DO_S, 1, 1>(pg, dist, ux, uy, uz, put); break;
      case 5: if constexpr (LMAX >= 2) skLaneEvalCanon<DO_H, DO_S, 1, 2>(pg, dist, ux, uy, uz, put); break;
      case 6: if constexpr (LMAX >= 2) skLaneEvalCanon<DO_H, DO_S, 2, 0>(pg, dist, ux, uy, uz, put); break;
      case 7: if constexpr (LMAX >= 2) skLaneEvalCanon<DO_H, DO_S, 2, 1>(pg, dist, ux, uy, uz, put); break;
      case 8: if constexpr (LMAX >= 2) skLaneEvalCanon<DO_H, DO_S, 2, 2>(pg, dist, ux, uy, uz, put); break;
      default:
        skLaneEval<DO_H, DO_S, LMAX>(
            pg, dist, ux, uy, uz, [&](int row, int col, double h, double s) { put(row + n2 * col, h, s); },
            [](int, int) {});
    }
    // generic-proxy writes of this thread -> visible to the async proxy, then one bulk copy per matrix
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (valid) {
      if (DO_H) flushRow(gH, rowH, parH, nElem);
      if (DO_S) flushRow(gS, rowS, parS, nElem);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  // on-site blocks
  for (int iAt = p.firstAtom + blockIdx.x * nW + wIn; iAt <= p.lastAtom; iAt += gridDim.x * nW) {
    const int sp = __ldg(p.species + iAt) - 1;
    const int n = p.progs[sp + p.nSpecies * sp].nOrb1, nElem = n * n;
    const int ind = __ldg(p.iPair + (size_t)p.ld * iAt);
    for (int e = lane; e < nElem; e += 32) {
      const int a = e / n, b = e - a * n;
      if (DO_H) p.ham[ind + e] = (a == b) ? __ldg(p.selfE + a + p.mOrb * sp) : 0.0;
      if (DO_S) p.over[ind + e] = (a == b) ? 1.0 : 0.0;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// selfegy(mShell, nSpecies) -> per-orbital on-site energies (mOrb, nSpecies)
__global__ void k1_expand_selfegy(const double *selfegy, const int *orbShell, const int *nOrbSpecies,
                                  int mShell, int mOrb, int nSpecies, double *selfE) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mOrb * nSpecies) return;
  const int sp = i / mOrb, o = i - sp * mOrb;
  selfE[i] = o < nOrbSpecies[sp] ? selfegy[orbShell[i] + mShell * sp] : 0.0;
}

template <int LMAX>
int launchTyped(dftbp_b200_ctx *ctx, const K1Params &p, int grid, int warps, size_t smemBytes, bool h, bool s) {
  auto go = [&](auto kern) -> int {
    DFTBP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    kern<<<grid, 32 * warps, smemBytes, ctx->stream>>>(p);
    return 0;
  };
  if (h && s) return go(k1_pair_build<LMAX, true, true>);
  if (h) return go(k1_pair_build<LMAX, true, false>);
  return go(k1_pair_build<LMAX, false, true>);
}

int launchBuild(dftbp_b200_ctx *ctx, const double *selfegy, double *ham, double *over) {
  if (!ctx) return DFTBP_B200_EINVAL;
  Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "build: no geometry set");
  DFTBP_TRY(syncPrograms(ctx));
  const SortedPairs *sched;
  DFTBP_TRY(ensureSortedPairs(ctx, false, &sched));
  K1Params p;
  p.progs = ctx->dProgs; p.items = sched->items; p.total = sched->total;
  p.ld = g.mNeigh + 1; p.nSpecies = ctx->nSpecies; p.mOrb = ctx->mOrb;
  p.firstAtom = g.firstAtomMine; p.lastAtom = g.lastAtomMine;
  p.coords = g.coords; p.species = g.species; p.iNeighbour = g.iNeighbour;
  p.iPair = g.iPair; p.selfE = ctx->dSelfE; p.ham = nullptr; p.over = nullptr;
  const bool partial = ctx->nRanks > 1;  // other ranks' segments stay zero (assembleChunks sums)
  if (ham) {
    DFTBP_CHECK(ctx, selfegy != nullptr, "build_h0: selfegy is null");
    const double *dSelf;
    DFTBP_TRY(devIn(ctx, kSlotC, selfegy, (size_t)ctx->mShell * ctx->nSpecies, &dSelf));
    const int n = ctx->mOrb * ctx->nSpecies;
    k1_expand_selfegy<<<(n + 127) / 128, 128, 0, ctx->stream>>>(dSelf, ctx->dOrbShell, ctx->dNOrbSpecies,
                                                               ctx->mShell, ctx->mOrb, ctx->nSpecies,
                                                               ctx->dSelfE);
    ctx->launches++;
    DFTBP_TRY(devOut(ctx, kSlotA, ham, (size_t)g.sparseSize, false, &p.ham));
    if (partial) DFTBP_CUDA(ctx, cudaMemsetAsync(p.ham, 0, sizeof(double) * g.sparseSize, ctx->stream));
  }
  if (over) {
    DFTBP_TRY(devOut(ctx, kSlotB, over, (size_t)g.sparseSize, false, &p.over));
    if (partial) DFTBP_CUDA(ctx, cudaMemsetAsync(p.over, 0, sizeof(double) * g.sparseSize, ctx->stream));
  }
  if (g.nWorkMine > 0) {
    ScopedTimer timer(ctx, kTimerSparseH0S);
    // staging rows: per lane and matrix mOrb^2 elements + 1 (parity shift), 16-byte aligned rows whose
    // stride in 16-byte units is odd (spreads the lanes' same-element writes over the banks)
    p.pitch = (ctx->mOrb * ctx->mOrb + 2) & ~1;
    if (((p.pitch >> 1) & 1) == 0) p.pitch += 2;
    const size_t perWarp = sizeof(double) * 32 * (size_t)p.pitch * ((ham && over) ? 2 : 1);
    int warps = (int)std::min<size_t>(8, ctx->smemOptin / perWarp);
    DFTBP_CHECK(ctx, warps >= 1, "build: staging tile does not fit into shared memory");
    const size_t smemBytes = perWarp * warps;
    const int ctasPerSm = std::max<int>(1, (int)std::min<size_t>(ctx->smemPerSm / (smemBytes + 1024), 2048 / (32 * warps)));
    const long long chunks = (g.nWorkMine + 31) / 32 + ctx->nSpecies * ctx->nSpecies;
    const int grid = (int)std::max<long long>(1, std::min<long long>((chunks + warps - 1) / warps,
                                                                     (long long)ctx->smCount * ctasPerSm));
    int rc;
    if (ctx->lMaxAll <= 1) rc = launchTyped<1>(ctx, p, grid, warps, smemBytes, ham != nullptr, over != nullptr);
    else if (ctx->lMaxAll == 2) rc = launchTyped<2>(ctx, p, grid, warps, smemBytes, ham != nullptr, over != nullptr);
    else rc = launchTyped<3>(ctx, p, grid, warps, smemBytes, ham != nullptr, over != nullptr);
    if (rc) return rc;
    ctx->launches++;
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_build_h0(dftbp_b200_ctx *ctx, const double *selfegy, double *ham) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ham != nullptr, "build_h0: ham is null");
  return launchBuild(ctx, selfegy, ham, nullptr);
}

int dftbp_b200_build_s(dftbp_b200_ctx *ctx, double *over) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, over != nullptr, "build_s: over is null");
  return launchBuild(ctx, nullptr, nullptr, over);
}

int dftbp_b200_build_h0_s(dftbp_b200_ctx *ctx, const double *selfegy, double *ham, double *over) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ham != nullptr && over != nullptr, "build_h0_s: null output");
  return launchBuild(ctx, selfegy, ham, over);
}

}  // extern "C"

// k1_pair_build.cu -- buildH0 / buildS (dftb/nonscc.F90:66-181, 364-424) as one sm_100a kernel.
//
// One warp per sparse block (iAtom, iNeigh).  iNeigh = 0 writes the on-site block (diagonal =
// on-site energies or 1, nonscc.F90:109-116 / :166-173); iNeigh >= 1 evaluates the diatomic block
// (sk_eval.cuh) and streams it out.  Lane e owns elements e, e+32, ... of the column-major block,
// so each store instruction of a warp covers one contiguous 256-byte run of the sparse array.
// Every element of this rank's segment is written exactly once: no memset, no atomics.
//
// Roofline: HBM-write bound for d-block systems, 16*n1*n2 + 8 algorithmic bytes per pair
// (H0 + S fused); near the FP64 ridge for sp-sp (DESIGN.md section "K1").
#include "sk_eval.cuh"

using namespace dftbp_b200;

namespace {

struct K1Params {
  SkDeviceTables tb;
  const int *work;
  long long nWork;
  int ld, nSpecies, mOrb;
  const double *coords;
  const int *species;
  const int *iNeighbour;
  const int *iPair;
  const double *selfE;  // (mOrb, nSpecies)
  double *ham;
  double *over;
};

template <bool DO_H, bool DO_S>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
k1_pair_build(const K1Params p) {
  __shared__ WarpScratch scratch[kWarpsPerCta];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5;
  WarpScratch &ws = scratch[wIn];
  const long long nWarps = (long long)gridDim.x * kWarpsPerCta;
  for (long long w = (long long)blockIdx.x * kWarpsPerCta + wIn; w < p.nWork; w += nWarps) {
    const int flat = __ldg(p.work + w);
    const int iAt1 = flat / p.ld, iNeigh = flat - iAt1 * p.ld;
    const int iAt2 = __ldg(p.iNeighbour + flat) - 1;
    const int sp1 = __ldg(p.species + iAt1) - 1, sp2 = __ldg(p.species + iAt2) - 1;
    const int ind = __ldg(p.iPair + flat);
    const SpPairProg &pg = p.tb.progs[sp1 + p.nSpecies * sp2];
    if (iNeigh == 0) {
      const int n = pg.nOrb1, nElem = n * n;
      for (int e = lane; e < nElem; e += 32) {
        const int a = e / n, b = e - a * n;
        if (DO_H) p.ham[ind + e] = (a == b) ? __ldg(p.selfE + a + p.mOrb * sp1) : 0.0;
        if (DO_S) p.over[ind + e] = (a == b) ? 1.0 : 0.0;
      }
      continue;
    }
    // vect = coords(:,iAt2) - coords(:,iAt1); dist = sqrt(sum(vect**2)); vect = vect/dist
    const double vx = __ldg(p.coords + 3 * iAt2) - __ldg(p.coords + 3 * iAt1);
    const double vy = __ldg(p.coords + 3 * iAt2 + 1) - __ldg(p.coords + 3 * iAt1 + 1);
    const double vz = __ldg(p.coords + 3 * iAt2 + 2) - __ldg(p.coords + 3 * iAt1 + 2);
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    double *__restrict__ hOut = p.ham + ind;
    double *__restrict__ sOut = p.over + ind;
    skEvalBlock<DO_H, DO_S>(p.tb, pg, dist, vx / dist, vy / dist, vz / dist, ws, lane,
                            [&](int e, double h, double s) {
                              if (DO_H) hOut[e] = h;
                              if (DO_S) sOut[e] = s;
                            });
  }
}

// selfegy(mShell, nSpecies) -> per-orbital on-site energies (mOrb, nSpecies)
__global__ void k1_expand_selfegy(const double *selfegy, const int *orbShell, const int *nOrbSpecies,
                                  int mShell, int mOrb, int nSpecies, double *selfE) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mOrb * nSpecies) return;
  const int sp = i / mOrb, o = i - sp * mOrb;
  selfE[i] = o < nOrbSpecies[sp] ? selfegy[orbShell[i] + mShell * sp] : 0.0;
}

int launchBuild(dftbp_b200_ctx *ctx, const double *selfegy, double *ham, double *over) {
  if (!ctx) return DFTBP_B200_EINVAL;
  Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "build: no geometry set");
  DFTBP_TRY(syncPrograms(ctx));
  K1Params p;
  p.tb = {ctx->dProgs, ctx->dADesc, ctx->dElemStart, ctx->dTerms, (const YDescDev *)ctx->dYDesc};
  p.work = g.workMine; p.nWork = g.nWorkMine; p.ld = g.mNeigh + 1; p.nSpecies = ctx->nSpecies;
  p.mOrb = ctx->mOrb; p.coords = g.coords; p.species = g.species; p.iNeighbour = g.iNeighbour;
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
  if (p.nWork > 0) {
    ScopedTimer timer(ctx, kTimerSparseH0S);
    const int grid = gridFor(ctx, p.nWork, kWarpsPerCta, 8 * 4);
    if (ham && over) k1_pair_build<true, true><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
    else if (ham) k1_pair_build<true, false><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
    else k1_pair_build<false, true><<<grid, 32 * kWarpsPerCta, 0, ctx->stream>>>(p);
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

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
#include <cstdlib>

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

// One work item = (chunk of 32 sorted pairs, matrix): H0 and S of a chunk are separate items; the
// kernel is latency bound, not FP64 bound, so the harmonics recomputed for the second matrix are
// free, while the smaller register and shared-memory footprint doubles the resident warps.
//
// Write-out, one atom-1 shell ("column group") at a time: the (2 l1 + 1) * nOrb2 elements of a group
// are contiguous in the column-major block.  Lane q stages its group in row q of the warp's
// [32][pitch] tile (odd pitch: the lanes' writes of one element hit 32 different banks); then the
// warp streams the 32 runs out with consecutive lanes on consecutive elements -- flat index
// idx = lane + 32 k over the 32 * count staged elements, (row, element) = (idx / count, idx % count)
// by a multiply-shift -- so all 32 lanes store in every step and every store instruction covers at
// most two contiguous runs of the sparse array.  (A TMA bulk copy per block was tried: UBLKCP takes
// uniform-register operands, so the 32 copies of a warp serialise in a vote loop -- slower.)
//
// The index loads of an item (schedule entry -> iNeighbour / iPair -> coordinates, species) are a
// chain of dependent global loads; they are software-pipelined two items ahead.
template <int LMAX>
__global__ void __launch_bounds__(256, 2) k1_pair_build(const K1Params p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *stage = smem + (size_t)wIn * 32 * p.pitch;
  double *row = stage + (size_t)lane * p.pitch;
  const int nChunks = __ldg(p.total) >> 5;
  const int nMat = (p.ham && p.over) ? 2 : 1;
  const int nItems = nChunks * nMat;
  const int stride = gridDim.x * nW;
  auto loadFlat = [&](int it) -> int {
    if (it >= nItems) return -1;
    int f = __ldg(p.items + 32 * (nMat == 2 ? it >> 1 : it) + lane);
    const int f0 = __shfl_sync(0xffffffffu, f, 0);  // lane 0 of a chunk is never padding
    return f >= 0 ? f : -2 - f0;                    // padding lanes shadow lane 0's pair: encoded < -1
  };
  auto flatOf = [](int f) { return f >= 0 ? f : -2 - f; };
  int item = blockIdx.x * nW + wIn;
  // pipeline registers: C = current item, N = next, NN = the one after
  int fC = loadFlat(item), fN = loadFlat(item + stride);
  int indC = 0;
  double c1x = 0, c1y = 0, c1z = 0, c2x = 0, c2y = 0, c2z = 0;
  int s1C = 1, s2C = 1;
  if (fC != -1) {
    const int f = flatOf(fC), iAt1 = f / p.ld;
    const int iAt2 = __ldg(p.iNeighbour + f) - 1;
    indC = __ldg(p.iPair + f);
    c1x = __ldg(p.coords + 3 * iAt1); c1y = __ldg(p.coords + 3 * iAt1 + 1); c1z = __ldg(p.coords + 3 * iAt1 + 2);
    c2x = __ldg(p.coords + 3 * iAt2); c2y = __ldg(p.coords + 3 * iAt2 + 1); c2z = __ldg(p.coords + 3 * iAt2 + 2);
    s1C = __ldg(p.species + iAt1);
    s2C = __ldg(p.species + iAt2);
  }
  for (; item < nItems; item += stride) {
    // stage 1 of item + 2 strides, stage 2 of item + stride
    const int fNN = loadFlat(item + 2 * stride);
    int iAt2N = 0, indN = 0, iAt1N = 0;
    if (fN != -1) {
      const int f = flatOf(fN);
      iAt1N = f / p.ld;
      iAt2N = __ldg(p.iNeighbour + f) - 1;
      indN = __ldg(p.iPair + f);
    }
    // ---- current item
    const bool isS = nMat == 2 ? (item & 1) : (p.ham == nullptr);
    const bool valid = fC >= 0;
    const int sp1 = __shfl_sync(0xffffffffu, s1C, 0) - 1, sp2 = __shfl_sync(0xffffffffu, s2C, 0) - 1;
    const SpPairProg &pg = p.progs[sp1 + p.nSpecies * sp2];
    const SkTableDev tab = isS ? pg.S : pg.H;
    const int nElem = pg.nOrb1 * pg.nOrb2;
    double *__restrict__ outBase = isS ? p.over : p.ham;
    // vect = coords(:,iAt2) - coords(:,iAt1); dist = sqrt(sum(vect**2)); vect = vect/dist
    const double vx = c2x - c1x, vy = c2y - c1y, vz = c2z - c1z;
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const double rinv = 1.0 / dist;
    const int desc = valid ? indC : -1;
    skLaneEvalAny<LMAX>(
        pg, tab, dist, vx * rinv, vy * rinv, vz * rinv, [&](int, int e, double v) { row[e] = v; },
        [&](int first, int count) {
          __syncwarp();
          const unsigned M = (1048576u + (unsigned)count - 1u) / (unsigned)count;  // idx / count = idx * M >> 20
          double *__restrict__ dst = outBase + first;
#pragma unroll 4
          for (int k = 0; k < count; ++k) {
            const unsigned idx = (unsigned)(lane + 32 * k);
            const unsigned q = (idx * M) >> 20;
            const unsigned e = idx - q * (unsigned)count;
            const int base = __shfl_sync(0xffffffffu, desc, (int)q);
            const double v = stage[q * p.pitch + e];
            if (base >= 0) dst[base + e] = v;
          }
          __syncwarp();
        });
    // stage 3 of item + stride, then rotate the pipeline
    if (fN != -1) {
      c1x = __ldg(p.coords + 3 * iAt1N); c1y = __ldg(p.coords + 3 * iAt1N + 1); c1z = __ldg(p.coords + 3 * iAt1N + 2);
      c2x = __ldg(p.coords + 3 * iAt2N); c2y = __ldg(p.coords + 3 * iAt2N + 1); c2z = __ldg(p.coords + 3 * iAt2N + 2);
      s1C = __ldg(p.species + iAt1N);
      s2C = __ldg(p.species + iAt2N);
    }
    fC = fN; fN = fNN; indC = indN;
  }
  // on-site blocks
  for (int iAt = p.firstAtom + blockIdx.x * nW + wIn; iAt <= p.lastAtom; iAt += gridDim.x * nW) {
    const int sp = __ldg(p.species + iAt) - 1;
    const int n = p.progs[sp + p.nSpecies * sp].nOrb1, nElem = n * n;
    const int ind = __ldg(p.iPair + (size_t)p.ld * iAt);
    for (int e = lane; e < nElem; e += 32) {
      const int a = e / n, b = e - a * n;
      if (p.ham) p.ham[ind + e] = (a == b) ? __ldg(p.selfE + a + p.mOrb * sp) : 0.0;
      if (p.over) p.over[ind + e] = (a == b) ? 1.0 : 0.0;
    }
  }
}

// ---- column-staged variant: H0 and S in one pass (canonical s / sp / spd species, one grid per pair) ----
// One work item = one chunk of 32 sorted pairs, both matrices.  Every lane evaluates its pair column
// by column (skLaneEvalCanonCols) and parks one column (nOrb2 contiguous elements) per matrix in its
// row of the warp's [32][PITCH] tiles; the warp then streams the 32 columns out with the flat index
// idx = lane + 32 k over the 32 * nOrb2 staged elements, (pair, element) = (idx / nOrb2, idx % nOrb2)
// with nOrb2 a compile-time constant.  Compared with k1_pair_build: the harmonics, interpolation
// weights and Gaunt sums are shared by H0 and S, the index arithmetic of the write-out is paid once
// for both matrices, and the tile shrinks from 32 x 45 to 2 x 32 x 9 doubles per warp, which leaves
// the SK tables an L1 of > 150 KB per SM.
constexpr int kNoPair = -(1 << 30);

template <int LM1, int LM2, bool DO_H, bool DO_S>
__device__ __forceinline__ void k1ColsItem(const SpPairProg &pg, double r, double dx, double dy, double dz, int lane,
                                           int desc, double *__restrict__ stageH, double *__restrict__ stageS,
                                           double *__restrict__ ham, double *__restrict__ over) {
  constexpr int N2 = (LM2 + 1) * (LM2 + 1);
  constexpr int PITCH = N2 | 1;  // odd: the lanes' writes of one element hit 32 different banks
  double *rowH = stageH + lane * PITCH, *rowS = stageS + lane * PITCH;
  // element (pair q, e) goes to base_q + N2 * col + e = (base_q - q * N2) + idx + N2 * col
  const int dq = desc >= 0 ? desc - lane * N2 : kNoPair;
  skLaneEvalCanonCols<LM1, LM2, DO_H, DO_S>(
      pg.H, pg.S, r, dx, dy, dz,
      [&](int e, double h, double sv) {
        if (DO_H) rowH[e] = h;
        if (DO_S) rowS[e] = sv;
      },
      [&](int col) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < N2; ++k) {
          const int idx = lane + 32 * k;
          const int q = idx / N2;
          const int dv = __shfl_sync(0xffffffffu, dq, q);
          const int sIdx = idx + q * (PITCH - N2);
          if (dv > kNoPair / 2) {
            if (DO_H) ham[dv + idx + N2 * col] = stageH[sIdx];
            if (DO_S) over[dv + idx + N2 * col] = stageS[sIdx];
          }
        }
        __syncwarp();
      });
}

template <int LMAX, bool DO_H, bool DO_S>
__global__ void __launch_bounds__(256, 2) k1_pair_build_cols(const K1Params p) {
  extern __shared__ double smem[];
  constexpr int PMAX = ((LMAX + 1) * (LMAX + 1)) | 1;
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *stageH = smem + (size_t)wIn * 64 * PMAX;
  double *stageS = stageH + 32 * PMAX;
  const int nItems = __ldg(p.total) >> 5;
  const int stride = gridDim.x * nW;
  auto loadFlat = [&](int it) -> int {
    if (it >= nItems) return -1;
    int f = __ldg(p.items + 32 * it + lane);
    const int f0 = __shfl_sync(0xffffffffu, f, 0);  // lane 0 of a chunk is never padding
    return f >= 0 ? f : -2 - f0;                    // padding lanes shadow lane 0's pair: encoded < -1
  };
  auto flatOf = [](int f) { return f >= 0 ? f : -2 - f; };
  int item = blockIdx.x * nW + wIn;
  int fC = loadFlat(item), fN = loadFlat(item + stride);
  int indC = 0;
  double c1x = 0, c1y = 0, c1z = 0, c2x = 0, c2y = 0, c2z = 0;
  int s1C = 1, s2C = 1;
  if (fC != -1) {
    const int f = flatOf(fC), iAt1 = f / p.ld;
    const int iAt2 = __ldg(p.iNeighbour + f) - 1;
    indC = __ldg(p.iPair + f);
    c1x = __ldg(p.coords + 3 * iAt1); c1y = __ldg(p.coords + 3 * iAt1 + 1); c1z = __ldg(p.coords + 3 * iAt1 + 2);
    c2x = __ldg(p.coords + 3 * iAt2); c2y = __ldg(p.coords + 3 * iAt2 + 1); c2z = __ldg(p.coords + 3 * iAt2 + 2);
    s1C = __ldg(p.species + iAt1);
    s2C = __ldg(p.species + iAt2);
  }
  for (; item < nItems; item += stride) {
    const int fNN = loadFlat(item + 2 * stride);
    int iAt2N = 0, indN = 0, iAt1N = 0;
    if (fN != -1) {
      const int f = flatOf(fN);
      iAt1N = f / p.ld;
      iAt2N = __ldg(p.iNeighbour + f) - 1;
      indN = __ldg(p.iPair + f);
    }
    const int sp1 = __shfl_sync(0xffffffffu, s1C, 0) - 1, sp2 = __shfl_sync(0xffffffffu, s2C, 0) - 1;
    const SpPairProg &pg = p.progs[sp1 + p.nSpecies * sp2];
    // vect = coords(:,iAt2) - coords(:,iAt1); dist = sqrt(sum(vect**2)); vect = vect/dist
    const double vx = c2x - c1x, vy = c2y - c1y, vz = c2z - c1z;
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const double rinv = 1.0 / dist;
    const double dx = vx * rinv, dy = vy * rinv, dz = vz * rinv;
    const int desc = fC >= 0 ? indC : -1;
    switch (pg.canon1 * 3 + pg.canon2) {
      case 0: k1ColsItem<0, 0, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 1: if constexpr (LMAX >= 1) k1ColsItem<0, 1, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 2: if constexpr (LMAX >= 2) k1ColsItem<0, 2, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 3: if constexpr (LMAX >= 1) k1ColsItem<1, 0, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 4: if constexpr (LMAX >= 1) k1ColsItem<1, 1, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 5: if constexpr (LMAX >= 2) k1ColsItem<1, 2, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 6: if constexpr (LMAX >= 2) k1ColsItem<2, 0, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 7: if constexpr (LMAX >= 2) k1ColsItem<2, 1, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      case 8: if constexpr (LMAX >= 2) k1ColsItem<2, 2, DO_H, DO_S>(pg, dist, dx, dy, dz, lane, desc, stageH, stageS, p.ham, p.over); break;
      default: break;
    }
    if (fN != -1) {
      c1x = __ldg(p.coords + 3 * iAt1N); c1y = __ldg(p.coords + 3 * iAt1N + 1); c1z = __ldg(p.coords + 3 * iAt1N + 2);
      c2x = __ldg(p.coords + 3 * iAt2N); c2y = __ldg(p.coords + 3 * iAt2N + 1); c2z = __ldg(p.coords + 3 * iAt2N + 2);
      s1C = __ldg(p.species + iAt1N);
      s2C = __ldg(p.species + iAt2N);
    }
    fC = fN; fN = fNN; indC = indN;
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
int launchTyped(dftbp_b200_ctx *ctx, const K1Params &p, int grid, int warps, size_t smemBytes) {
  DFTBP_CUDA(ctx, cudaFuncSetAttribute(k1_pair_build<LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smemBytes));
  k1_pair_build<LMAX><<<grid, 32 * warps, smemBytes, ctx->stream>>>(p);
  return 0;
}

template <int LMAX, bool DO_H, bool DO_S>
int launchCols(dftbp_b200_ctx *ctx, const K1Params &p, long long items) {
  constexpr int PMAX = ((LMAX + 1) * (LMAX + 1)) | 1;
  const int warps = 8;
  const size_t smemBytes = sizeof(double) * 64 * PMAX * warps;
  int perSm = 1;
  DFTBP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k1_pair_build_cols<LMAX, DO_H, DO_S>,
                                                                32 * warps, smemBytes));
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((items + warps - 1) / warps, (long long)ctx->smCount * std::max(perSm, 1)));
  k1_pair_build_cols<LMAX, DO_H, DO_S><<<grid, 32 * warps, smemBytes, ctx->stream>>>(p);
  return 0;
}

template <int LMAX>
int launchColsAny(dftbp_b200_ctx *ctx, const K1Params &p, long long items) {
  if (p.ham && p.over) return launchCols<LMAX, true, true>(ctx, p, items);
  if (p.ham) return launchCols<LMAX, true, false>(ctx, p, items);
  return launchCols<LMAX, false, true>(ctx, p, items);
}

// the column-staged kernel needs canonical s / sp / spd species and one grid per species pair
bool colsEligible(const dftbp_b200_ctx *ctx) {
  if (ctx->lMaxAll > 2) return false;
  const char *env = getenv("DFTBP_B200_K1");
  if (env && !strcmp(env, "groups")) return false;  // A/B switch for tools/gpu_probe.py
  for (const SpPairProg &pg : ctx->progs) {
    if (pg.canon1 < 0 || pg.canon2 < 0 || pg.canon1 > 2 || pg.canon2 > 2) return false;
    if (pg.H.nGrid != pg.S.nGrid || pg.H.dist != pg.S.dist || pg.H.rMax != pg.S.rMax) return false;
  }
  return true;
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
    // staging row per lane: the largest column group, (2 lmax + 1) * mOrb elements, odd pitch
    p.pitch = ((2 * ctx->lMaxAll + 1) * ctx->mOrb) | 1;
    const size_t perWarp = sizeof(double) * 32 * (size_t)p.pitch;
    // 8 warps per CTA, 2 CTAs per SM: 16 warps x 128 registers fill the register file
    int warps = (int)std::min<size_t>(8, (ctx->smemOptin - 2048) / perWarp);
    DFTBP_CHECK(ctx, warps >= 1, "build: staging tile does not fit into shared memory");
    const size_t smemBytes = perWarp * warps;
    const int ctasPerSm = std::max<int>(1, (int)std::min<size_t>(ctx->smemPerSm / (smemBytes + 1024), 2));
    const long long items = ((g.nWorkMine + 31) / 32 + ctx->nSpecies * ctx->nSpecies) * ((ham && over) ? 2 : 1);
    const int grid = (int)std::max<long long>(1, std::min<long long>((items + warps - 1) / warps,
                                                                     (long long)ctx->smCount * ctasPerSm));
    int rc;
    if (colsEligible(ctx)) {
      const long long chunks = (g.nWorkMine + 31) / 32 + ctx->nSpecies * ctx->nSpecies;
      if (ctx->lMaxAll == 0) rc = launchColsAny<0>(ctx, p, chunks);
      else if (ctx->lMaxAll == 1) rc = launchColsAny<1>(ctx, p, chunks);
      else rc = launchColsAny<2>(ctx, p, chunks);
    } else if (ctx->lMaxAll <= 1) rc = launchTyped<1>(ctx, p, grid, warps, smemBytes);
    else if (ctx->lMaxAll == 2) rc = launchTyped<2>(ctx, p, grid, warps, smemBytes);
    else rc = launchTyped<3>(ctx, p, grid, warps, smemBytes);
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

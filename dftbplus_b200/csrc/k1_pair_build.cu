// k1_pair_build.cu -- buildH0 / buildS (dftb/nonscc.F90:66-181, 364-424) as one sm_100a kernel.
//
// Lane-per-pair: every thread evaluates the diatomic block of one neighbour pair with the
// straight-line evaluator of sk_lane_eval.cuh; the 32 pairs of a warp come from the device-built
// schedule of pair_sort.cu (one species pair, neighbouring distances), which also leaves one
// 32-byte record per pair (bond vector, sparse offset, species-pair program): a lane starts from
// ONE coalesced load instead of the chain schedule -> iNeighbour -> iPair / species -> coords.
// A thread's block is scattered over 648-byte strides in the sparse array, so the warp transposes
// through shared memory: each lane writes its elements into its own row of a [32][pitch] staging
// tile (odd pitch: conflict free), then the warp streams the 32 runs out one pair after the other
// with lanes on consecutive elements.  On-site blocks (iNeigh = 0; diagonal = on-site energies or
// 1, nonscc.F90:109-116 / :166-173) are written by a second grid-stride loop, one warp per atom.
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
  const PairRec *recs;  // sorted pair schedule with per-pair records (pair_sort.cu)
  const int *total;     // its padded length
  int ld, nSpecies, mOrb;
  int firstAtom, lastAtom;  // on-site blocks of this rank
  const int *species;
  const int *iPair;
  const double *selfegy;  // (mShell, nSpecies) on-site energies per shell (atomEigVal)
  const int *orbShell;    // (mOrb, nSpecies) shell index of each orbital
  int mShell;
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
// warp walks over the 32 rows and copies each run with lanes on consecutive elements (one or two
// store instructions per run for s/p/d).  The base offset of a run is warp-uniform, so a run costs
// one shuffle and an LDS/STG pair per 32 elements.  (The earlier flat mapping idx = lane + 32 k ->
// (row, element) kept all 32 lanes busy in every store but paid a multiply-shift division, a
// shuffle and a predicate per ELEMENT: 44 % of the kernel's instructions,
// profiles/r01e_ncu_k1_source_hotspots.txt.  A TMA bulk copy per run is not possible: the runs start
// at arbitrary 8-byte offsets, cp.async.bulk needs 16.)
template <int LMAX, int WARPS, int CTAS>
__global__ void __launch_bounds__(32 * WARPS, CTAS) k1_pair_build(const K1Params p) {
  extern __shared__ double smem[];
  // staging row per lane: the largest column group, (2 lmax + 1) * mOrb elements, odd pitch
  constexpr int PITCH = ((2 * LMAX + 1) * (LMAX + 1) * (LMAX + 1)) | 1;
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *stage = smem + (size_t)wIn * 32 * PITCH;
  double *row = stage + (size_t)lane * PITCH;
  const int nChunks = __ldg(p.total) >> 5;
  const int nMat = (p.ham && p.over) ? 2 : 1;
  const int nItems = nChunks * nMat;
  const int stride = gridDim.x * nW;
  // record of this lane's pair in item `it`: two 16-byte loads, coalesced across the warp
  auto loadRec = [&](int it, double2 &xy, double &z, int &ind, int &sp) {
    if (it >= nItems) return;
    const PairRec *r = p.recs + 32 * (size_t)(nMat == 2 ? it >> 1 : it) + lane;
    xy = __ldg(reinterpret_cast<const double2 *>(r));
    const int4 w = __ldg(reinterpret_cast<const int4 *>(r) + 1);
    z = __hiloint2double(w.y, w.x);
    ind = w.z;
    sp = w.w;
  };
  int item = blockIdx.x * nW + wIn;
  double2 xyN = make_double2(0.0, 0.0);
  double zN = 1.0;
  int indN = -1, spN = 0;
  loadRec(item, xyN, zN, indN, spN);
  for (; item < nItems; item += stride) {
    double vx = xyN.x, vy = xyN.y, vz = zN;
    const int ind = indN;
    const int sp = __shfl_sync(0xffffffffu, spN, 0);  // lane 0 of a chunk is never padding
    loadRec(item + stride, xyN, zN, indN, spN);        // the next item's record is in flight during this one
    {
      // padding lanes shadow lane 0's pair (they evaluate a valid geometry and store nothing)
      const double x0 = __shfl_sync(0xffffffffu, vx, 0), y0 = __shfl_sync(0xffffffffu, vy, 0),
                   z0 = __shfl_sync(0xffffffffu, vz, 0);
      if (ind < 0) { vx = x0; vy = y0; vz = z0; }
    }
    const bool isS = nMat == 2 ? (item & 1) : (p.ham == nullptr);
    const SpPairProg &pg = p.progs[sp];
    const SkTableDev tab = isS ? pg.S : pg.H;
    double *__restrict__ outLane = (isS ? p.over : p.ham) + lane;
    // dist = sqrt(sum(vect**2)); vect = vect/dist
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const double rinv = 1.0 / dist;
    skLaneEvalAny<LMAX>(
        pg, tab, dist, vx * rinv, vy * rinv, vz * rinv, [&](int, int e, double v) { row[e] = v; },
        [&](int first, int count) {
          __syncwarp();
          // run q: `count` staged elements of row q -> out[ind_q + first ..); lanes on consecutive
          // elements.  One shuffle (the run's base, warp-uniform), predicated LDS / STG pairs.
          const double *__restrict__ src = stage + lane;
          const bool m0 = lane < count, m1 = lane + 32 < count;
#pragma unroll 8
          for (int q = 0; q < 32; ++q) {
            const int base = __shfl_sync(0xffffffffu, ind, q);
            const bool on = base >= 0;
            double *__restrict__ d = outLane + (first + base);
            const double *__restrict__ s = src + q * PITCH;
            double v0 = 0.0, v1 = 0.0;
            if (m0) v0 = s[0];
            if (m1) v1 = s[32];
            if (on && m0) d[0] = v0;
            if (on && m1) d[32] = v1;
            if constexpr (LMAX >= 3) {  // f shells: runs of up to 7 * 16 elements
              for (int e = lane + 64; e < count; e += 32)
                if (on) d[e - lane] = s[e - lane];
            }
          }
          __syncwarp();
        });
  }
  // on-site blocks
  for (int iAt = p.firstAtom + blockIdx.x * nW + wIn; iAt <= p.lastAtom; iAt += gridDim.x * nW) {
    const int sp = __ldg(p.species + iAt) - 1;
    const int n = p.progs[sp + p.nSpecies * sp].nOrb1, nElem = n * n;
    const int ind = __ldg(p.iPair + (size_t)p.ld * iAt);
    for (int e = lane; e < nElem; e += 32) {
      const int a = e / n, b = e - a * n;
      // ham(iOrig + iOrb) = selfegy(iShellOrb(iOrb, sp), sp) on the diagonal (nonscc.F90:109-116)
      if (p.ham)
        p.ham[ind + e] = (a == b) ? __ldg(p.selfegy + __ldg(p.orbShell + a + p.mOrb * sp) + p.mShell * sp) : 0.0;
      if (p.over) p.over[ind + e] = (a == b) ? 1.0 : 0.0;
    }
  }
}

template <int LMAX, int WARPS, int CTAS>
int launchTyped(dftbp_b200_ctx *ctx, const K1Params &p, long long items) {
  constexpr int PITCH = ((2 * LMAX + 1) * (LMAX + 1) * (LMAX + 1)) | 1;
  const size_t smemBytes = sizeof(double) * 32 * (size_t)PITCH * WARPS;
  DFTBP_CHECK(ctx, smemBytes <= ctx->smemOptin, "build: staging tile does not fit into shared memory");
  DFTBP_CUDA(ctx, cudaFuncSetAttribute(k1_pair_build<LMAX, WARPS, CTAS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  int perSm = 1;
  DFTBP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k1_pair_build<LMAX, WARPS, CTAS>,
                                                                32 * WARPS, smemBytes));
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((items + WARPS - 1) / WARPS, (long long)ctx->smCount * std::max(perSm, 1)));
  k1_pair_build<LMAX, WARPS, CTAS><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p);
  return 0;
}

// resident warps per SM (fec4096): the time hardly depends on it -- 8 warps at 255 registers 75 us,
// 12 at 168 63 us, 16 at 128 65 us, 18 at 96 66 us; the staging tiles (11.5 KB per warp for spd) allow at most 19
template <int LMAX>
int launchTypedAny(dftbp_b200_ctx *ctx, const K1Params &p, long long items) {
  const char *env = getenv("DFTBP_B200_K1_CFG");  // tuning switch (tools/gpu_probe.py)
  const int cfg = env ? atoi(env) : 82;
  if (cfg == 62) return launchTyped<LMAX, 6, 2>(ctx, p, items);  // 12 warps, 168 registers: 63 us warm, 74 us with L2 flushed
  if (cfg == 63) return launchTyped<LMAX, 6, 3>(ctx, p, items);  // 18 warps, 96 registers: 66 us
  return launchTyped<LMAX, 8, 2>(ctx, p, items);                  // 16 warps, 128 registers: 65 us warm, 69 us flushed (default)
}

int launchBuild(dftbp_b200_ctx *ctx, const double *selfegy, double *ham, double *over) {
  if (!ctx) return DFTBP_B200_EINVAL;
  Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "build: no geometry set");
  DFTBP_TRY(syncPrograms(ctx));
  const SortedPairs *sched;
  DFTBP_TRY(ensureSortedPairs(ctx, false, &sched));
  K1Params p;
  p.progs = ctx->dProgs; p.recs = sched->recs; p.total = sched->total;
  p.ld = g.mNeigh + 1; p.nSpecies = ctx->nSpecies; p.mOrb = ctx->mOrb;
  p.firstAtom = g.firstAtomMine; p.lastAtom = g.lastAtomMine;
  p.species = g.species;
  p.iPair = g.iPair; p.selfegy = nullptr; p.orbShell = ctx->dOrbShell; p.mShell = ctx->mShell;
  p.ham = nullptr; p.over = nullptr;
  // zero fill only where the kernel does not write every element: other ranks' segments stay zero
  // (assembleChunks sums), or the caller's arrays are larger than the blocks
  const bool partial = ctx->nRanks > 1 || !g.coversAll;
  if (ham) {
    DFTBP_CHECK(ctx, selfegy != nullptr, "build_h0: selfegy is null");
    DFTBP_TRY(devIn(ctx, kSlotC, selfegy, (size_t)ctx->mShell * ctx->nSpecies, &p.selfegy));
    DFTBP_TRY(devOut(ctx, kSlotA, ham, (size_t)g.sparseSize, false, &p.ham));
    if (partial) DFTBP_CUDA(ctx, cudaMemsetAsync(p.ham, 0, sizeof(double) * g.sparseSize, ctx->stream));
  }
  if (over) {
    DFTBP_TRY(devOut(ctx, kSlotB, over, (size_t)g.sparseSize, false, &p.over));
    if (partial) DFTBP_CUDA(ctx, cudaMemsetAsync(p.over, 0, sizeof(double) * g.sparseSize, ctx->stream));
  }
  if (g.nWorkMine > 0) {
    ScopedTimer timer(ctx, kTimerSparseH0S);
    const long long items = ((g.nWorkMine + 31) / 32 + ctx->nSpecies * ctx->nSpecies) * ((ham && over) ? 2 : 1);
    int rc;
    if (ctx->lMaxAll <= 1) rc = launchTypedAny<1>(ctx, p, items);
    else if (ctx->lMaxAll == 2) rc = launchTypedAny<2>(ctx, p, items);
    else rc = launchTyped<3, 6, 1>(ctx, p, items);
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

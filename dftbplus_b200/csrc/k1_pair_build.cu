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
#include <cstdint>
#include <cstdlib>

#include "sk_dmat_eval.cuh"
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
  int pitch;  // staging row of the compact kernel: mOrb^2 | 1 doubles
  int dbgNoStore;  // experiment: evaluate and stage, store nothing
  // two-kernel build
  const int *items;        // per schedule slot: position of the pair in the work list
  const PairRec *recsMem;  // records in work-list order, on-site blocks included
  double *exch;            // exchange rows, one per work-list entry
  int nWork, imgCap, bulkOK;
  int dbgSkip;  // timing experiments on k1_interp: 1 = no window copies, 2 = no interpolation, 4 = no row stores
};

// ---- round 2: compact evaluator -------------------------------------------------------------------
// One work item = one chunk of 32 sorted pairs, BOTH matrices: the bond frame and the rotation
// matrices D^1, D^2 (sk_dmat_eval.cuh) are formed once per pair and shared by the H0 and the S
// block; the interpolation weights too when the two tables live on the same grid.  Every lane
// writes its whole nOrb2 x nOrb1 block into its row of the warp's [32][pitch] tile (odd pitch:
// conflict free), then the warp streams the 32 blocks out one after the other with lanes on
// consecutive elements -- one shuffle and ceil(nElem / 32) LDS / STG pairs per block -- and the tile
// is reused for the second matrix.  The evaluator is a run-time loop over the shell pairs with
// nine small compile-time bodies (a few KB of SASS in total; the Gaunt code of round 1 was 243 KB
// and streamed through the instruction cache once per item).
template <int LMAX, int WARPS, int CTAS>
__global__ void __launch_bounds__(32 * WARPS, CTAS) k1_pair_build_d(const K1Params p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  const int pitch = p.pitch;
  double *stage = smem + (size_t)wIn * 32 * pitch;
  double *row = stage + (size_t)lane * pitch;
  const unsigned stageLane = (unsigned)__cvta_generic_to_shared(stage + lane);
  const int nChunks = __ldg(p.total) >> 5;
  const int stride = gridDim.x * nW;
  auto loadRec = [&](int it, double2 &xy, double &z, int &ind, int &sp) {
    if (it >= nChunks) return;
    const PairRec *r = p.recs + 32 * (size_t)it + lane;
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
  for (; item < nChunks; item += stride) {
    double vx = xyN.x, vy = xyN.y, vz = zN;
    int ind = indN;
    const int sp = __shfl_sync(0xffffffffu, spN, 0);  // lane 0 of a chunk is never padding
    loadRec(item + stride, xyN, zN, indN, spN);        // the next item's record is in flight during this one
    {
      // padding lanes duplicate lane 0's pair: same block, same destination, same values -- the
      // write-out loop then needs no "is this row live" test
      const double x0 = __shfl_sync(0xffffffffu, vx, 0), y0 = __shfl_sync(0xffffffffu, vy, 0),
                   z0 = __shfl_sync(0xffffffffu, vz, 0);
      const int i0 = __shfl_sync(0xffffffffu, ind, 0);
      if (ind < 0) { vx = x0; vy = y0; vz = z0; ind = i0; }
    }
    const SpPairProg &pg = p.progs[sp];
    // dist = sqrt(sum(vect**2)); vect = vect/dist
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const double rinv = 1.0 / dist;
    const double ux = vx * rinv, uy = vy * rinv, uz = vz * rinv;
    skd::DMat D;
    skd::buildD<(LMAX >= 2)>(ux, uy, uz, D);
    double Y[LMAX >= 3 ? 49 : 1];
    if constexpr (LMAX >= 3) skg::harmonics<6>(ux, uy, uz, Y);
    else Y[0] = 1.0;
    const int nElem = pg.nOrb1 * pg.nOrb2;
    const bool m0 = lane < nElem && !p.dbgNoStore, m1 = lane + 32 < nElem && !p.dbgNoStore, m2 = lane + 64 < nElem && !p.dbgNoStore;
    LaneInterp li;
    for (int mat = 0; mat < 2; ++mat) {
      double *__restrict__ out = mat ? p.over : p.ham;
      if (out == nullptr) continue;
      const SkTableDev &tab = mat ? pg.S : pg.H;
      // the S table usually sits on the grid of the H table: same window, same weights
      if (mat == 0 || p.ham == nullptr || pg.sameGrid == 0) laneInterpSetup(tab, dist, li);
      skd::evalBlock<LMAX>(pg, tab, li, D, Y, row);
      __syncwarp();
      // block q: nElem staged elements of row q -> out[ind_q ..); lanes on consecutive elements.
      // Per block: one shuffle, one 64-bit address, ceil(nElem / 32) LDS / STG pairs.  The shared
      // loads are unconditional (the tile is padded by 96 elements), only the stores are predicated.
      double *__restrict__ outLane = out + lane;
      if (nElem <= 96) {
        unsigned sa = stageLane;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) {
          const int base = __shfl_sync(0xffffffffu, ind, q);
          double *__restrict__ d = outLane + base;
          double v0, v1, v2;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(sa));
          asm volatile("ld.shared.f64 %0, [%1+256];" : "=d"(v1) : "r"(sa));
          asm volatile("ld.shared.f64 %0, [%1+512];" : "=d"(v2) : "r"(sa));
          if (m0) d[0] = v0;
          if (m1) d[32] = v1;
          if (m2) d[64] = v2;
          sa += 8u * (unsigned)pitch;
        }
      } else {
        for (int q = 0; q < 32; ++q) {
          const int base = __shfl_sync(0xffffffffu, ind, q);
          double *__restrict__ d = outLane + base;
          const double *__restrict__ s = stage + lane + q * pitch;
          for (int e = 0; e + lane < nElem; e += 32) d[e] = s[e];
        }
      }
      __syncwarp();
    }
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
int launchCompact(dftbp_b200_ctx *ctx, K1Params p, long long chunks) {
  p.pitch = (ctx->mOrb * ctx->mOrb) | 1;
  p.dbgNoStore = tuneInt("DFTBP_B200_K1_NOSTORE", 0);
  const size_t smemBytes = sizeof(double) * (32 * (size_t)p.pitch * WARPS + 96);
  DFTBP_CHECK(ctx, smemBytes <= ctx->smemOptin, "build: staging tile does not fit into shared memory");
  static LaunchInfo info;
  int perSm = 1;
  DFTBP_TRY(launchInfo(ctx, info, k1_pair_build_d<LMAX, WARPS, CTAS>, 32 * WARPS, smemBytes, &perSm));
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((chunks + WARPS - 1) / WARPS, (long long)ctx->smCount * std::max(perSm, 1)));
  k1_pair_build_d<LMAX, WARPS, CTAS><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p);
  return 0;
}


// ---- round 2: two-kernel build, outputs written in the order of the sparse arrays ------------------
// Measured on a B200 (tools/ubench/write_patterns.cu): writing the 648-byte blocks of the pair
// build in the distance-sorted order of the schedule reaches 3.25 TB/s -- the two sectors at the ends
// of every block are shared with blocks that another warp writes much later, so they leave L2 as
// partial writes -- while the same bytes written in memory order as one linear image per 32 blocks
// with cp.async.bulk reach 5.3 TB/s.  The interpolation, however, wants the sorted order (32 lanes
// reading the same one or two table lines; in memory order every lane gathers from its own lines:
// 3584 L1 wavefronts per 32 pairs).  Hence two kernels:
//   k1_interp        sorted schedule, lane per pair: distance, interpolation window and weights,
//                    all integrals of H0 and S -> one 8 * VROW byte row per pair of the exchange
//                    buffer, ordered by shell-pair type (29 MB at 131 k pairs: stays in L2)
//   k1_rotate_store  work-list order (= memory order), lane per block, on-site blocks included: bond
//                    frame, D^1 / D^2, the rotation of every shell-pair type straight from registers
//                    into the warp's linear image of its 32 consecutive blocks in shared memory,
//                    then ONE 1-D bulk copy (TMA) of the image to the sparse array per matrix.
// Lanes of a warp now differ in species pair: the shell-pair types are visited in a fixed order and
// a lane takes part where its pair has that type (its sub-block offset comes from caseOff of its
// own program), so a warp pays for the union of its lanes' types (an spd-spd block for an Fe row,
// whatever the neighbours are).  On-site blocks run through the same code with the bond frame =
// the identity and "integrals" = the on-site energies (H0) or 1 (S): D is then an exact signed
// permutation and the block comes out as exactly diag(selfegy) / the unit matrix.
// Needs species without repeated l ("simple"), l <= 2, blocks that tile the array, 16-byte aligned
// outputs; anything else takes k1_pair_build_d above.

template <int LMAX>
struct ExchRow {
  static constexpr int nSlot = skd::caseSlot<LMAX>(16);
  static constexpr int vh = (nSlot + 1) & ~1;  // doubles per matrix (even: 16-byte loads)
  static constexpr int row = 2 * vh;           // H0 integrals, then S integrals
};

// The 32 pairs of a chunk sit in one or two distance bins, so their 8-row table windows overlap almost
// completely: the warp copies the union of the windows -- rows [jBase, jBase + nRows) of every
// integral column of H0 and S -- into shared memory with one batch of cp.async (ONE exposed L2 round
// trip per chunk), and every lane then interpolates from there (broadcast LDS).  Loading the windows
// lane by lane from global memory cost one exposed round trip per group of integrals and 224 LDG
// per pair; profiles/r2_k1_history.md.
constexpr int kWinRows = 20;  // rows per window the shared tile holds; wider chunks take the global path

template <int LMAX>
__global__ void __launch_bounds__(128) k1_interp(const K1Params p) {
  constexpr int VH = ExchRow<LMAX>::vh, VROW = ExchRow<LMAX>::row, PITCH = VROW | 1;
  constexpr int NI = ExchRow<LMAX>::nSlot;  // most integrals a species pair of simple species can have
  // per warp: the lanes' rows, transposed through shared memory so that every row leaves the warp as
  // one contiguous 8 * VROW byte store, and the table windows of the two matrices
  __shared__ double stageAll[4][32 * PITCH];
  __shared__ double winAll[4][2 * NI * kWinRows];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *stage = stageAll[wIn];
  double *mine = stage + lane * PITCH;
  double *win = winAll[wIn];
  const unsigned winS = (unsigned)__cvta_generic_to_shared(win);
  const int nChunks = __ldg(p.total) >> 5;
  for (int chunk = blockIdx.x * nW + wIn; chunk < nChunks; chunk += gridDim.x * nW) {
    const PairRec *r = p.recs + 32 * (size_t)chunk + lane;
    const double2 xy = __ldg(reinterpret_cast<const double2 *>(r));
    const int4 w = __ldg(reinterpret_cast<const int4 *>(r) + 1);
    const int slot = __ldg(p.items + 32 * (size_t)chunk + lane);  // row of the pair, -1 = padding
    double vx = xy.x, vy = xy.y, vz = __hiloint2double(w.y, w.x);
    const int sp = __shfl_sync(0xffffffffu, w.w, 0);  // lane 0 of a chunk is never padding
    {
      const double x0 = __shfl_sync(0xffffffffu, vx, 0), y0 = __shfl_sync(0xffffffffu, vy, 0),
                   z0 = __shfl_sync(0xffffffffu, vz, 0);
      if (w.z < 0) { vx = x0; vy = y0; vz = z0; }  // padding lanes evaluate a valid distance, store nothing
    }
    const SpPairProg &pg = p.progs[sp];
    const double dist = sqrt(vx * vx + vy * vy + vz * vz);
    const int nInteg = pg.nInteg;
    const bool same = pg.sameGrid != 0;
    LaneInterp liH, liS;
    laneInterpSetup(pg.H, dist, liH);
    if (same) liS = liH;
    else laneInterpSetup(pg.S, dist, liS);
    // union of the lanes' windows per matrix
    int jLo[2], nRows[2];
    bool inSmem = true;
#pragma unroll
    for (int mat = 0; mat < 2; ++mat) {
      const LaneInterp &li = mat ? liS : liH;
      int lo = li.mode == 0 ? li.j0 : 0x7fffffff, hi = li.mode == 0 ? li.j0 : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      jLo[mat] = lo;
      nRows[mat] = hi >= 0 ? hi - lo + 8 : 0;
      inSmem = inSmem && nRows[mat] <= kWinRows && nInteg <= NI;
    }
    if (inSmem && !(p.dbgSkip & 1)) {
#pragma unroll
      for (int mat = 0; mat < 2; ++mat) {
        if ((mat ? p.over : p.ham) == nullptr) continue;
        const SkTableDev &t = mat ? pg.S : pg.H;
        const int nR = nRows[mat], n = nR * nInteg;
        const double *__restrict__ src = t.tab + jLo[mat];
        const unsigned dst = winS + 8u * (unsigned)(mat * NI * kWinRows);
        const unsigned M = nR > 0 ? (65536u + (unsigned)nR - 1u) / (unsigned)nR : 0u;  // idx / nR = idx * M >> 16 (idx < 768)
        for (int idx = lane; idx < n; idx += 32) {
          const unsigned k = ((unsigned)idx * M) >> 16, rr = (unsigned)idx - k * (unsigned)nR;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (k * kWinRows + rr)),
                       "l"(src + (size_t)k * t.nGrid + rr)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
    }
    for (int mat = 0; mat < 2; ++mat) {
      if ((mat ? p.over : p.ham) == nullptr || (p.dbgSkip & 2)) continue;
      const SkTableDev &t = mat ? pg.S : pg.H;
      const LaneInterp &li = mat ? liS : liH;
      double *__restrict__ dst = mine + mat * VH;
      if (li.mode == 0 && inSmem) {
        const double *__restrict__ y = win + mat * NI * kWinRows + (li.j0 - jLo[mat]);
#pragma unroll 2
        for (int k = 0; k < nInteg; ++k) {
          const double *__restrict__ yc = y + k * kWinRows;
          double acc = li.w[0] * yc[0];
#pragma unroll
          for (int j = 1; j < 8; ++j) acc = fma(li.w[j], yc[j], acc);
          dst[pg.integSlot[k]] = acc;
        }
      } else {
        for (int k = 0; k < nInteg; ++k) dst[pg.integSlot[k]] = laneInterp(t, li, k);
      }
    }
    __syncwarp();
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
      const int row = __shfl_sync(0xffffffffu, slot, q);
      if (row >= 0 && lane < VROW && !(p.dbgSkip & 4)) p.exch[(size_t)row * VROW + lane] = stage[q * PITCH + lane];
    }
    __syncwarp();
  }
}

template <int LMAX, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) k1_rotate_store(const K1Params p) {
  constexpr int VH = ExchRow<LMAX>::vh, VROW = ExchRow<LMAX>::row;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *img = smem + (size_t)wIn * p.imgCap;
  const int nWork = p.nWork, nItems = (nWork + 31) >> 5;
  const int blkCap = p.mOrb * p.mOrb;
  const int stride = gridDim.x * nW;
  const bool doH = p.ham != nullptr, doS = p.over != nullptr;
  // Software pipeline over the global loads (ten warps per SM cannot hide an L2 round trip of
  // ~1000 cycles on their own): the record and the first integral row of the NEXT item are
  // requested while the second matrix of the current one is rotated, the second row of the current
  // item while its first matrix is rotated.
  struct Rec { double2 xy; int4 w; };
  auto loadRec = [&](int it, Rec &r) {
    const int i = 32 * it + lane;
    r.xy = make_double2(0.0, 0.0);
    r.w = make_int4(0, 0x3ff00000, 0, 0);  // z = 1.0
    if (it < nItems && i < nWork) {
      const PairRec *q = p.recsMem + i;
      r.xy = __ldg(reinterpret_cast<const double2 *>(q));
      r.w = __ldg(reinterpret_cast<const int4 *>(q) + 1);
    }
  };
  // integrals ordered by shell-pair type: the row of work-list entry i (rows of on-site blocks exist
  // but are never written: loaded, then replaced)
  auto loadRow = [&](int it, int mat, double (&v)[VH]) {
    const int i = 32 * it + lane;
    if (it < nItems && i < nWork) {
      const double2 *__restrict__ src = reinterpret_cast<const double2 *>(p.exch + (size_t)i * VROW + mat * VH);
#pragma unroll
      for (int k = 0; k < VH / 2; ++k) {
        const double2 t = __ldg(src + k);
        v[2 * k] = t.x; v[2 * k + 1] = t.y;
      }
    }
  };
  int item = blockIdx.x * nW + wIn;
  Rec rec;
  double vA[VH], vB[VH];
#pragma unroll
  for (int k = 0; k < VH; ++k) vA[k] = vB[k] = 0.0;
  loadRec(item, rec);
  loadRow(item, doH ? 0 : 1, vA);
  for (; item < nItems; item += stride) {
    const int i = 32 * item + lane;
    const bool valid = i < nWork;
    const double vx = rec.xy.x, vy = rec.xy.y, vz = __hiloint2double(rec.w.y, rec.w.x);
    const int ind = rec.w.z, spRaw = rec.w.w;
    if (doH && doS) loadRow(item, 1, vB);
    const bool onsite = (spRaw & kRecOnsite) != 0;
    const int sp = spRaw & (kRecOnsite - 1);
    const SpPairProg *__restrict__ pgL = p.progs + sp;
    const int n2 = pgL->nOrb2;
    const int nElem = valid ? pgL->nOrb1 * n2 : 0;
    // sub-block offsets of this lane's species pair, one 16-bit entry per shell-pair type
    unsigned co[8];
    {
      const uint4 a = __ldg(reinterpret_cast<const uint4 *>(pgL->caseOff)),
                  b = __ldg(reinterpret_cast<const uint4 *>(pgL->caseOff) + 1);
      co[0] = a.x; co[1] = a.y; co[2] = a.z; co[3] = a.w; co[4] = b.x; co[5] = b.y; co[6] = b.z; co[7] = b.w;
    }
    // do the 32 blocks tile one contiguous piece of the sparse array?
    const int nValid = min(32, nWork - 32 * item);
    const int indNext = __shfl_down_sync(0xffffffffu, ind, 1);
    const bool chained = !valid || lane + 1 >= nValid || indNext == ind + nElem;
    const bool linear = p.bulkOK && __all_sync(0xffffffffu, chained);
    const int ind0 = __shfl_sync(0xffffffffu, ind, 0);
    const int total = __shfl_sync(0xffffffffu, ind + nElem, nValid - 1) - ind0;
    const int shift = ind0 & 1;  // image element j <-> array element ind0 + j, same 16-byte phase
    const int off = linear ? ind - ind0 + shift : lane * blkCap;
    // bond frame and rotation matrices; on-site blocks: the identity frame
    skd::DMat D;
    {
      double ux = 0.0, uy = 0.0, uz = 1.0;
      if (valid && !onsite) {
        // dist = sqrt(sum(vect**2)); vect = vect/dist
        const double rinv = 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
        ux = vx * rinv; uy = vy * rinv; uz = vz * rinv;
      }
      skd::buildD<(LMAX >= 2)>(ux, uy, uz, D);
    }
    double *blk = img + off;
    // one matrix: integrals v -> the 32 blocks of the item in the image -> the sparse array
    auto emitMatrix = [&](double (&v)[VH], int mat) {
      double *__restrict__ out = mat ? p.over : p.ham;
      if (onsite) {
        // an on-site block: the on-site energy of the shell (H0: ham(iOrig + iOrb) = selfegy(iShellOrb(iOrb,
        // sp), sp), nonscc.F90:109-116) / 1 (S) on the (l, l) types, 0 elsewhere
        const int sp1 = sp % p.nSpecies;
#pragma unroll
        for (int k = 0; k < VH; ++k) v[k] = 0.0;
#pragma unroll
        for (int l = 0; l <= LMAX; ++l) {
          const int sh = pgL->shellOfL1[l];
          if (sh >= 0) {
            const double eps = mat ? 1.0 : __ldg(p.selfegy + sh + p.mShell * sp1);
#pragma unroll
            for (int k = 0; k <= l; ++k) v[skd::caseSlot<LMAX>(5 * l) + k] = eps;
          }
        }
      }
      // the previous bulk copy of this warp must have read the image before it is overwritten
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      staticFor<16>([&](auto cc) {
        constexpr int c = decltype(cc)::value, l1 = c >> 2, l2 = c & 3;
        if constexpr (l1 <= LMAX && l2 <= LMAX) {
          const unsigned o = (co[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
          const bool has = valid && o != 0xFFFFu;
          if (__any_sync(0xffffffffu, has)) {
            if (has) skd::rotBlockV<l1, l2>(D, v + skd::caseSlot<LMAX>(c), blk + o, n2);
          }
        }
      });
      // generic-proxy writes of every lane -> visible to the async proxy (bulk copy engine)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (linear) {
        if (lane == 0) {
          const int head = shift, body = (total - head) & ~1, tail = total - head - body;
          if (head) out[ind0] = img[shift];
          if (body > 0) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(img + shift + head);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + ind0 + head),
                         "r"(src), "r"(8 * body)
                         : "memory");
          }
          if (tail) out[ind0 + total - 1] = img[shift + total - 1];
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        // blocks with gaps between them (over-allocated arrays): block by block, lanes on consecutive elements
        for (int q = 0; q < nValid; ++q) {
          const int base = __shfl_sync(0xffffffffu, ind, q), n = __shfl_sync(0xffffffffu, nElem, q);
          for (int e = lane; e < n; e += 32) out[base + e] = img[q * blkCap + e];
        }
        __syncwarp();
      }
    };
    if (doH && doS) {
      emitMatrix(vA, 0);
      loadRec(item + stride, rec);
      loadRow(item + stride, 0, vA);
      emitMatrix(vB, 1);
    } else {
      double vNow[VH];
#pragma unroll
      for (int k = 0; k < VH; ++k) vNow[k] = vA[k];
      loadRec(item + stride, rec);
      loadRow(item + stride, doH ? 0 : 1, vA);
      emitMatrix(vNow, doH ? 0 : 1);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class T>
int growExchange(dftbp_b200_ctx *ctx, long long need) {
  if (ctx->pairExchangeCap >= need) return 0;
  if (ctx->pairExchange) cudaFree(ctx->pairExchange);
  ctx->pairExchange = nullptr;
  ctx->pairExchangeCap = 0;
  cudaError_t e = cudaMalloc((void **)&ctx->pairExchange, sizeof(T) * (size_t)need);
  if (e != cudaSuccess)
    return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc(%lld) for the pair exchange buffer failed: %s",
                (long long)(sizeof(T) * need), cudaGetErrorString(e));
  ctx->pairExchangeCap = need;
  return 0;
}

template <int LMAX, int WARPS>
int launchTwoKernel(dftbp_b200_ctx *ctx, K1Params p, long long chunks) {
  const Geometry &g = ctx->geo;
  DFTBP_TRY(growExchange<double>(ctx, (long long)ExchRow<LMAX>::row * std::max<long long>(g.nWorkMine, 1)));
  p.exch = ctx->pairExchange;
  p.nWork = (int)g.nWorkMine;
  p.imgCap = 32 * ctx->mOrb * ctx->mOrb + 2;
  p.dbgSkip = tuneInt("DFTBP_B200_K1I_SKIP", 0);
  const int only = tuneInt("DFTBP_B200_K1_ONLY", 0);  // timing experiment: 1 = interpolation kernel only, 2 = rotation only
  if (only != 2) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((chunks + 3) / 4, (long long)ctx->smCount * 16));
    k1_interp<LMAX><<<grid, 128, 0, ctx->stream>>>(p);
  }
  if (only == 1) return 0;
  const size_t smemBytes = sizeof(double) * (size_t)p.imgCap * WARPS;
  DFTBP_CHECK(ctx, smemBytes <= ctx->smemOptin, "build: block images do not fit into shared memory");
  static LaunchInfo info;
  DFTBP_TRY(launchInfo(ctx, info, k1_rotate_store<LMAX, WARPS>, 32 * WARPS, smemBytes, nullptr));
  const long long items = (g.nWorkMine + 31) / 32;
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + WARPS - 1) / WARPS, (long long)ctx->smCount));
  k1_rotate_store<LMAX, WARPS><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p);
  ctx->launches++;
  return 0;
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
    int rc;
    const long long chunks = (g.nWorkMine + 31) / 32 + ctx->nSpecies * ctx->nSpecies;
    p.items = sched->items; p.recsMem = sched->recsMem;
    p.bulkOK = (((uintptr_t)p.ham | (uintptr_t)p.over) & 15) == 0;
    const int kcfg = tuneInt("DFTBP_B200_K1_KERNEL", 1);  // 0: one-kernel form in schedule order (A/B)
    const bool twoKernel = kcfg != 0 && ctx->allSimple && ctx->lMaxAll <= 2 && g.coversAll;
    if (twoKernel) {
      if (ctx->lMaxAll <= 1) rc = launchTwoKernel<1, 16>(ctx, p, chunks);
      else rc = launchTwoKernel<2, 8>(ctx, p, chunks);  // 8 warps x 254 registers: 32 us; 9 / 10 warps (168 registers, spills): 47 us
    } else {
      // one-kernel form in schedule order; warps per CTA by what the staging rows (mOrb^2 doubles per lane)
      // leave of the shared memory: species with repeated shells have blocks of up to 16 x 16
      const size_t perWarp = sizeof(double) * 32 * (size_t)((ctx->mOrb * ctx->mOrb) | 1);
      const int fit = (int)((ctx->smemOptin - 1024) / perWarp);
      if (ctx->lMaxAll <= 1) rc = fit >= 16 ? launchCompact<1, 8, 2>(ctx, p, chunks) : fit >= 4 ? launchCompact<1, 4, 1>(ctx, p, chunks)
                                                                                                 : launchCompact<1, 1, 1>(ctx, p, chunks);
      else if (ctx->lMaxAll == 2) rc = fit >= 10 ? launchCompact<2, 5, 2>(ctx, p, chunks) : fit >= 3 ? launchCompact<2, 3, 1>(ctx, p, chunks)
                                                                                                    : launchCompact<2, 1, 1>(ctx, p, chunks);
      else rc = fit >= 3 ? launchCompact<3, 3, 1>(ctx, p, chunks) : launchCompact<3, 1, 1>(ctx, p, chunks);
    }
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

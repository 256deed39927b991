// k5_forces.cu -- derivative_shift (dftb/forces.F90:27-38, 119-217, 548-669) with the reference's
// finite-difference derivator (getFirstDerivFiniteDiff, dftb/nonscc.F90:428-484).
//
// The reference materialises H' and S' (mOrb x mOrb x 3 each, 12 full SK evaluations per pair)
// and then contracts them; here the six displaced blocks are contracted on the fly in registers --
// no dH/dR ever exists in memory:
//
//   g(j) = 2 [ sum rho.H'_j - sum Erho.S'_j + sum_spin sum rho_s . 1/2 (S'_j V1_s + V2_s S'_j) ]
//        = 2 sum_e [ wH_e H'_j,e + wS_e S'_j,e ],
//   wH = rho_1,  wS = -Erho + sum_spin 1/2 (rho_s V1_s^T + V2_s^T rho_s)        (independent of j)
//   X'_j = (X(v - delta e_j) - X(v + delta e_j)) / (2 delta)                      (nonscc.F90:471-481)
//
// so g(j) = (1/delta) sum_e [ wH_e (H_e(-) - H_e(+)) + wS_e (S_e(-) - S_e(+)) ];
// deriv(:,iAtom1) += g, deriv(:,iAtom2f) -= g.
//
// Lane-per-pair, like K1: every thread owns one neighbour pair of the sorted schedule
// (pair_sort.cu) and runs the straight-line evaluator of sk_lane_eval.cuh six times per matrix.
// One work item = (chunk of 32 pairs, matrix): the H part (weights rho_1) and the S part (weights
// wS) are independent sums, so they are separate items and a warp stages ONE weight array:
//   1. the warp loads the 32 weight blocks cooperatively (lanes on consecutive elements: coalesced)
//      into its [32][pitch] shared tile, forming wS on the way;
//   2. every lane contracts its six displaced blocks against its own row of the tile;
//   3. six atomic adds per lane.
// FP64 CUDA-core bound (6 block evaluations per pair and matrix), not HBM bound.
#include <cstdlib>

#include "block_desc.cuh"
#include "flat_stream.cuh"
#include "sk_lane_eval.cuh"

using namespace dftbp_b200;

namespace {

struct K5Params {
  const SpPairProg *progs;
  const PairRec *recs;  // sorted pair schedule with per-pair records (pair_sort.cu)
  const int2 *atoms;    // (iAtom1, iAtom2f) per slot
  const int *total;
  int ld, nSpecies, mOrb, nAtom, nSpin, sparseSize, pitch;
  int part, nParts;  // this call takes the chunks part, part + nParts, ... of the schedule (dftbp_b200_set_pair_partition)
  const double *DM;     // rho, first spin channel = wH
  const double *negWS;  // -wS: Erho (no SCC) or the array prepared by k5_force_weights
  double delta;
  double *deriv;   // (3, nAtom) or nullptr (stress only)
  double *stress;  // 9 raw accumulators sum fac * intermed(jj) * vect(ii) at [jj + 3 ii], or nullptr
};

// this lane's pair of chunk `chunk`: bond vector, sparse offset, the two atoms, species-pair
// program (warp-uniform).  Padding lanes shadow lane 0's pair and are not live; pairs whose
// neighbour folds onto the atom itself do not contribute to the forces (forces.F90:186, :631) but
// do contribute to the stress (stress.F90:164-176).
struct K5Pair {
  double v0, v1, v2;
  int ind, iAt1, iAt2f, sp;
  bool live;
};

__device__ __forceinline__ K5Pair k5LoadPair(const K5Params &p, int chunk, int lane) {
  const PairRec *r = p.recs + 32 * (size_t)chunk + lane;
  const double2 xy = __ldg(reinterpret_cast<const double2 *>(r));
  const int4 w = __ldg(reinterpret_cast<const int4 *>(r) + 1);
  const int2 at = __ldg(p.atoms + 32 * (size_t)chunk + lane);
  K5Pair k;
  k.v0 = xy.x; k.v1 = xy.y; k.v2 = __hiloint2double(w.y, w.x);
  k.ind = w.z;
  k.sp = __shfl_sync(0xffffffffu, w.w, 0);  // lane 0 of a chunk is never padding
  const double x0 = __shfl_sync(0xffffffffu, k.v0, 0), y0 = __shfl_sync(0xffffffffu, k.v1, 0),
               z0 = __shfl_sync(0xffffffffu, k.v2, 0);
  const bool valid = k.ind >= 0;
  if (!valid) { k.v0 = x0; k.v1 = y0; k.v2 = z0; }
  k.iAt1 = at.x; k.iAt2f = at.y;
  k.live = valid && (at.x != at.y || p.stress != nullptr);
  return k;
}

// flag = 1 if any off-diagonal element of any shift block is non-zero (one warp per block); the
// diagonals go to diag(mOrb, nBlocks), compact: the element-wise kernels read V(a,a) of the two atoms
// of every sparse element, and in the blocks every value sits in a sector of its own
__global__ void __launch_bounds__(256) k5_shift_structure(const double *__restrict__ shift, int mOrb, int nBlocks,
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

// -wS = Erho - sum_spin 1/2 (rho_s V1_s^T + V2_s^T rho_s) for every off-site block of the call's chunks
// (the SCC part of derivative_blockEuclidean, forces.F90:640-656, independent of the Cartesian
// direction): a streaming kernel.  Plain SCC and collinear-spin shifts are diagonal blocks (the
// device flag says so); then (rho V1^T + V2^T rho)(b,a) = rho(b,a) (V1(a,a) + V2(b,b)) and no matrix
// product is needed: spin loop unrolled, all loads of a block -- (1 + 3 NSPIN) per element, three
// elements per lane -- issued before the first use.
// It walks along the pair schedule: one warp per chunk of 32 sorted pairs (one species pair: the
// block shape is warp-uniform), the 32 blocks one after the other with lanes on consecutive
// elements -- the form the partition of the force call needs (dftbp_b200_set_pair_partition: a call
// only touches the chunks part, part + nParts, ... and only their blocks get weights).
template <int NSPIN, bool FULL>
__device__ __forceinline__ void forceWeightsSched(const K5Params &p, const double *__restrict__ EDM,
                                                  const double *__restrict__ shift, double *__restrict__ negWS,
                                                  int selfImages) {
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  const int nChunks = __ldg(p.total) >> 5;
  const int mo2 = p.mOrb * p.mOrb, dstep = p.mOrb + 1;
  const int nSpin = FULL ? p.nSpin : NSPIN;
  for (int chunk = p.part + p.nParts * (blockIdx.x * nW + wIn); chunk < nChunks; chunk += p.nParts * gridDim.x * nW) {
    const PairRec *r = p.recs + 32 * (size_t)chunk + lane;
    const int4 w = __ldg(reinterpret_cast<const int4 *>(r) + 1);
    const int2 at = __ldg(p.atoms + 32 * (size_t)chunk + lane);
    const int sp = __shfl_sync(0xffffffffu, w.w, 0);  // lane 0 of a chunk is never padding
    const SpPairProg &pg = p.progs[sp];
    const int n1 = pg.nOrb1, n2 = pg.nOrb2, nElem = n1 * n2;
    const bool liveMine = w.z >= 0 && (at.x != at.y || selfImages);
    const unsigned M2 = (1048576u + (unsigned)n2 - 1u) / (unsigned)n2;  // e / n2 = e * M2 >> 20 (e < 256)
    // this lane's elements of every block: e = lane, lane + 32, lane + 64 (+ a loop for f shells)
    int a[3], b[3];
    bool on[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int e = lane + 32 * k;
      on[k] = e < nElem;
      const int ee = on[k] ? e : 0;
      a[k] = (int)(((unsigned)ee * M2) >> 20);  // element (row b, col a) of the n2 x n1 block
      b[k] = ee - a[k] * n2;
    }
    const unsigned liveMask = __ballot_sync(0xffffffffu, liveMine);
    for (int q = 0; q < 32; ++q) {
      if (!((liveMask >> q) & 1u)) continue;  // warp-uniform
      const int ind = __shfl_sync(0xffffffffu, w.z, q);
      const int iAt1 = __shfl_sync(0xffffffffu, at.x, q), iAt2f = __shfl_sync(0xffffffffu, at.y, q);
      const double *__restrict__ edm = EDM + ind;
      double *__restrict__ out = negWS + ind;
      if (!FULL) {
        double s[3], rr[3][NSPIN], x1[3][NSPIN], x2[3][NSPIN];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int e = lane + 32 * k;
          s[k] = on[k] ? __ldg(edm + e) : 0.0;
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
            rr[k][iSpin] = on[k] ? __ldg(p.DM + (size_t)p.sparseSize * iSpin + ind + e) : 0.0;
            x1[k][iSpin] = on[k] ? __ldg(shift + (size_t)mo2 * (iAt1 + (size_t)p.nAtom * iSpin) + a[k] * dstep) : 0.0;
            x2[k][iSpin] = on[k] ? __ldg(shift + (size_t)mo2 * (iAt2f + (size_t)p.nAtom * iSpin) + b[k] * dstep) : 0.0;
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double v = s[k];
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) v = fma(-0.5 * rr[k][iSpin], x1[k][iSpin] + x2[k][iSpin], v);
          if (on[k]) out[lane + 32 * k] = v;
        }
        for (int e = lane + 96; e < nElem; e += 32) {  // blocks with f shells
          const int aa = (int)(((unsigned)e * M2) >> 20), bb = e - aa * n2;
          double v = __ldg(edm + e);
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin)
            v = fma(-0.5 * __ldg(p.DM + (size_t)p.sparseSize * iSpin + ind + e),
                    __ldg(shift + (size_t)mo2 * (iAt1 + (size_t)p.nAtom * iSpin) + aa * dstep) +
                        __ldg(shift + (size_t)mo2 * (iAt2f + (size_t)p.nAtom * iSpin) + bb * dstep),
                    v);
          out[e] = v;
        }
      } else {
        for (int e = lane; e < nElem; e += 32) {
          const int aa = (int)(((unsigned)e * M2) >> 20), bb = e - aa * n2;
          double v = __ldg(edm + e);
          for (int iSpin = 0; iSpin < nSpin; ++iSpin) {
            const double *__restrict__ rho = p.DM + (size_t)p.sparseSize * iSpin + ind;
            const double *__restrict__ V1 = shift + (size_t)mo2 * (iAt1 + (size_t)p.nAtom * iSpin);
            const double *__restrict__ V2 = shift + (size_t)mo2 * (iAt2f + (size_t)p.nAtom * iSpin);
            // (rho V1^T)(b,a) = sum_c rho(b,c) V1(a,c);  (V2^T rho)(b,a) = sum_c V2(c,b) rho(c,a)
            double x = 0.0;
            for (int c = 0; c < n1; ++c) x = fma(__ldg(rho + bb + n2 * c), __ldg(V1 + aa + p.mOrb * c), x);
            for (int c = 0; c < n2; ++c) x = fma(__ldg(V2 + c + p.mOrb * bb), __ldg(rho + c + n2 * aa), x);
            v = fma(-0.5, x, v);
          }
          out[e] = v;
        }
      }
    }
  }
}

template <int NSPIN, bool FULL>
__global__ void __launch_bounds__(256) k5_force_weights_sched(const K5Params p, const double *__restrict__ EDM,
                                                              const double *__restrict__ shift,
                                                              const int *__restrict__ shiftFull,
                                                              double *__restrict__ negWS, int selfImages) {
  if ((__ldg(shiftFull) != 0) != FULL) return;  // the other instantiation handles this shift structure
  forceWeightsSched<NSPIN, FULL>(p, EDM, shift, negWS, selfImages);
}

// Diagonal shift blocks, the blocks visited in the order of the sparse arrays instead of along the
// schedule: reads and writes of consecutive blocks are then neighbours in memory (measured on
// fec4096: 75 us against 88 us for k5_force_weights_sched<2, false>, whose 648-byte pieces are
// scattered).  slotOf[i] = slot of work-list entry i in the schedule; a call with a partition only
// forms the weights of its own chunks ((slot >> 5) % nParts == part).
template <int NSPIN>
__global__ void __launch_bounds__(256) k5_force_weights_diag(const K5Params p, const BlockDesc *__restrict__ desc,
                                                             const int *__restrict__ slotOf, long long nWork,
                                                             const double *__restrict__ EDM,
                                                             const double *__restrict__ shift,
                                                             const int *__restrict__ shiftFull,
                                                             double *__restrict__ negWS, int selfImages) {
  if (__ldg(shiftFull) != 0) return;  // full blocks: k5_force_weights_sched<1, true>
  const int lane = threadIdx.x & 31;
  const int mo2 = p.mOrb * p.mOrb, dstep = p.mOrb + 1;
  forEachBlockBatchAux<1>(desc, slotOf, nWork, [&](const BlockDesc(&b)[1], const int(&slot)[1], int) {
    const bool mine = slot[0] >= 0 && (slot[0] >> 5) % p.nParts == p.part;
    if (!mine || (b[0].iAt1 == b[0].iAt2f && !selfImages)) return;
    const int n2 = b[0].n12 >> 8, nElem = (b[0].n12 & 0xFF) * n2;
    const unsigned M2 = (1048576u + (unsigned)n2 - 1u) / (unsigned)n2;  // e / n2 = e * M2 >> 20 (e < 256)
    for (int e0 = lane; e0 < nElem; e0 += 96) {
      double s[3], r[3][NSPIN], x1[3][NSPIN], x2[3][NSPIN];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = e0 + 32 * k;
        const bool on = e < nElem;
        const int ee = on ? e : 0;
        const int a = (int)(((unsigned)ee * M2) >> 20), bb = ee - a * n2;  // element (row bb, col a)
        s[k] = on ? __ldg(EDM + b[0].ind + ee) : 0.0;
#pragma unroll
        for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
          r[k][iSpin] = on ? __ldg(p.DM + (size_t)p.sparseSize * iSpin + b[0].ind + ee) : 0.0;
          x1[k][iSpin] = on ? __ldg(shift + (size_t)mo2 * (b[0].iAt1 + (size_t)p.nAtom * iSpin) + a * dstep) : 0.0;
          x2[k][iSpin] = on ? __ldg(shift + (size_t)mo2 * (b[0].iAt2f + (size_t)p.nAtom * iSpin) + bb * dstep) : 0.0;
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = e0 + 32 * k;
        double v = s[k];
#pragma unroll
        for (int iSpin = 0; iSpin < NSPIN; ++iSpin) v = fma(-0.5 * r[k][iSpin], x1[k][iSpin] + x2[k][iSpin], v);
        if (e < nElem) negWS[b[0].ind + e] = v;
      }
    }
  });
}

// Diagonal shift blocks through the span pipeline of flat_stream.cuh: Erho and the NSPIN channels of
// rho arrive by TMA bulk copies, lanes on consecutive elements of the sparse arrays whatever block
// they belong to; the shift diagonals come from the compact copy k5_shift_structure leaves.  Weights
// of every block of the list are formed (a force call with a partition reads only its own): used on
// one rank / one partition.
template <int NSPIN>
__global__ void __launch_bounds__(32 * flat::warpsFor(1 + NSPIN), 1)
    k5_force_weights_flat(const K5Params p, const BlockDesc *__restrict__ desc, long long nWork,
                          const int *__restrict__ groupFirst, const double *__restrict__ EDM,
                          const double *__restrict__ diag, const int *__restrict__ shiftFull,
                          double *__restrict__ negWS, const double *__restrict__ shift, int selfImages, int spanLo,
                          int spanHi, int *hint) {
  // what this call found, for the launches of the next one (pinned host word, mapped: one posted store)
  if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile int *>(hint) = __ldg(shiftFull) != 0 ? 1 : 0;
  if (__ldg(shiftFull) != 0) {  // full shift blocks (DFTB+U-like): per-pair matrix products along the schedule
    // selfImages bit 1: the dedicated full-block kernel has been queued as well and does this
    if (!(selfImages & 2)) forceWeightsSched<1, true>(p, EDM, shift, negWS, selfImages & 1);
    return;
  }
  constexpr int NIN = 1 + NSPIN, WARPS = flat::warpsFor(NIN);
  extern __shared__ __align__(128) unsigned char flatRaw[];
  flat::Smem<NIN, WARPS> &sm = *reinterpret_cast<flat::Smem<NIN, WARPS> *>(flatRaw);
  __shared__ unsigned m2[kMaxOrb + 1];
  flat::initDivTable(m2);
  const double *src[NIN];
  src[0] = EDM;
#pragma unroll
  for (int iSpin = 0; iSpin < NSPIN; ++iSpin) src[1 + iSpin] = p.DM + (size_t)p.sparseSize * iSpin;
  const int mOrb = p.mOrb;
  const int spinStride = mOrb * p.nAtom;  // diag(:, :, iSpin + 1) - diag(:, :, iSpin)
  flat::streamSpans<NIN, WARPS>(
      sm, src, p.sparseSize, desc, (int)nWork, groupFirst, spanLo, spanHi, [&](const flat::Elem<NIN>(&el)[flat::kBatch]) {
        double x1[flat::kBatch][NSPIN], x2[flat::kBatch][NSPIN];
#pragma unroll
        for (int u = 0; u < flat::kBatch; ++u) {
          const int n2 = el[u].lo.w >> 8, e = el[u].e;
          const int a = (int)(((unsigned)e * m2[n2]) >> 20), bb = e - a * n2;  // element (row bb, col a)
          const double *__restrict__ V1 = diag + (mOrb * el[u].hi.x + a);
          const double *__restrict__ V2 = diag + (mOrb * el[u].hi.y + bb);
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) {
            x1[u][iSpin] = __ldg(V1 + (size_t)spinStride * iSpin);
            x2[u][iSpin] = __ldg(V2 + (size_t)spinStride * iSpin);
          }
        }
#pragma unroll
        for (int u = 0; u < flat::kBatch; ++u) {
          double w = el[u].v[0];
#pragma unroll
          for (int iSpin = 0; iSpin < NSPIN; ++iSpin) w = fma(-0.5 * el[u].v[1 + iSpin], x1[u][iSpin] + x2[u][iSpin], w);
          if (el[u].ok) negWS[el[u].x] = w;
        }
      });
}

template <int LMAX, bool STRESS, int WARPS, int UNR>
__global__ void __launch_bounds__(32 * WARPS, 1) k5_derivative_shift(const K5Params p) {
  extern __shared__ double smem[];
  double st[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // STRESS: this lane's share of st(jj,ii)
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  double *stage = smem + (size_t)wIn * 32 * p.pitch;
  double *row = stage + (size_t)lane * p.pitch;
  const int nChunks = __ldg(p.total) >> 5;
  const int nMine = (nChunks - p.part + p.nParts - 1) / p.nParts;  // chunks part, part + nParts, ...
  const int nItems = 2 * nMine;
  for (int item = blockIdx.x * nW + wIn; item < nItems; item += gridDim.x * nW) {
    const int chunk = p.part + p.nParts * (item >> 1);
    const bool isS = item & 1;
    const K5Pair pr = k5LoadPair(p, chunk, lane);
    const bool live = pr.live;
    const int ind = pr.ind, iAt1 = pr.iAt1, iAt2f = pr.iAt2f;
    const SpPairProg &pg = p.progs[pr.sp];
    const SkTableDev tab = isS ? pg.S : pg.H;
    const int nElem = pg.nOrb1 * pg.nOrb2;

    // ---- 1. stage the contraction weights of the 32 pairs: rho_1 blocks for the H part, -wS blocks
    // (Erho without SCC, the array prepared by k5_force_weights with SCC) for the S part.  Flat
    // index idx = lane + 32 k over the 32 * nElem elements, (pair, element) = (idx / nElem,
    // idx % nElem) by multiply-shift; cp.async (LDGSTS) moves the 8-byte elements global -> shared
    // without passing through registers, so the whole tile is in flight at once: one HBM latency
    // per item.
    {
      const int desc = live ? ind : -1;
      const unsigned M = (1048576u + (unsigned)nElem - 1u) / (unsigned)nElem;
      const double *__restrict__ src = isS ? p.negWS : p.DM;
      const unsigned stageS = (unsigned)__cvta_generic_to_shared(stage);
#pragma unroll 4
      for (int k = 0; k < nElem; ++k) {
        const unsigned idx = (unsigned)(lane + 32 * k);
        const unsigned q = (idx * M) >> 20;
        const unsigned e = idx - q * (unsigned)nElem;
        const int bq = __shfl_sync(0xffffffffu, desc, (int)q);
        if (bq >= 0)  // dead pairs never read their row
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stageS + 8u * (q * (unsigned)p.pitch + e)),
                       "l"(src + bq + e)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    // ---- 2. fold the weights of this lane's pair into W (independent of the bond vector), then six
    // displaced evaluations, each an interpolation + harmonics + short dot product (sk_lane_eval.cuh)
    const double v0 = pr.v0, v1 = pr.v1, v2 = pr.v2;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (__any_sync(0xffffffffu, live)) {
      if (live) skLaneWBuildAny<LMAX>(pg, row);
#pragma unroll UNR
      for (int k = 0; k < 6; ++k) {
        const int jj = k >> 1;
        // vect(jj) = vect(jj) - real(2*ii-3) * deltaXDiff : ii=1 -> +delta, ii=2 -> -delta
        const double sgn = (k & 1) ? -1.0 : 1.0;
        const double d0 = jj == 0 ? v0 + sgn * p.delta : v0;
        const double d1 = jj == 1 ? v1 + sgn * p.delta : v1;
        const double d2 = jj == 2 ? v2 + sgn * p.delta : v2;
        const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double rinv = 1.0 / dist;
        double acc = live ? skLaneWDotAny<LMAX>(pg, tab, dist, d0 * rinv, d1 * rinv, d2 * rinv, row) : 0.0;
        acc *= isS ? sgn : -sgn;  // (tmp(ii=2) - tmp(ii=1)); the S tile holds -wS
        if (jj == 0) g0 += acc;
        if (jj == 1) g1 += acc;
        if (jj == 2) g2 += acc;
      }
    }
    // ---- 3. deriv(:,iAtom1) += g, deriv(:,iAtom2f) -= g
    if (live) {
      const double inv = 1.0 / p.delta;
      g0 *= inv; g1 *= inv; g2 *= inv;
      if (p.deriv && iAt1 != iAt2f) {
        atomicAdd(p.deriv + 3 * iAt1, g0);
        atomicAdd(p.deriv + 3 * iAt1 + 1, g1);
        atomicAdd(p.deriv + 3 * iAt1 + 2, g2);
        atomicAdd(p.deriv + 3 * iAt2f, -g0);
        atomicAdd(p.deriv + 3 * iAt2f + 1, -g1);
        atomicAdd(p.deriv + 3 * iAt2f + 2, -g2);
      }
      if (STRESS) {
        // st(jj,ii) += fac * intermed(jj) * vect(ii), vect = coords(:,iAtom1) - coords(:,iAtom2) = -v,
        // fac = 2 unless the neighbour is a periodic image of the atom itself (stress.F90:158-176)
        const double fac = iAt1 != iAt2f ? -2.0 : -1.0;
        const double g[3] = {fac * g0, fac * g1, fac * g2}, v[3] = {v0, v1, v2};
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) st[jj + 3 * ii] = fma(g[jj], v[ii], st[jj + 3 * ii]);
      }
    }
    __syncwarp();  // the tile is restaged by the next item
  }
  if (STRESS) {  // one reduction per warp and kernel: 9 atomics per warp
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      double x = st[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
      if (lane == 0) atomicAdd(p.stress + k, x);
    }
  }
}

// st = -0.5 * st / cellVol (stress.F90:182)
__global__ void k5_stress_scale(const double *__restrict__ acc, double cellVol, double *__restrict__ st) {
  if (threadIdx.x < 9) st[threadIdx.x] = -0.5 * acc[threadIdx.x] / cellVol;
}

// ---- Differentiation = Richardson (getFirstDerivRichardson, dftb/nonscc.F90:488-600) -----------------
// The reference extrapolates every matrix element of H' and S' separately: central differences
// with h = 1/8, 1/16, ... (powers of two, exactly representable), a Richardson table of at most
// three extrapolation columns, element-wise convergence test |dd(mm-1) - dd(mm)| <= 1e-11 * dd(mm)
// (no absolute value on the right-hand side: elements with a negative derivative never pass it and
// run until h reaches epsilon^(1/4), row 9), converged elements are frozen.  The table of a block
// stops growing when all its elements have converged, but a frozen element does not change any
// more and an unconverged one only exists when the table ran to the last row, so the result of an
// element does not depend on the other elements of the block.  That is data dependent per element
// and cannot be folded into the weights like the finite-difference form: every lane walks through
// its pair's block in slices of kSlice elements, keeps the last two table rows of the slice in
// local memory, re-evaluates the block for every slice and row (the evaluator is cheap compared
// with the bookkeeping) and contracts the converged derivatives with the weights.
// Rarely used (the default derivator is FiniteDiff); written for parity, not for speed.
template <int LMAX>
__global__ void __launch_bounds__(64) k5_derivative_shift_richardson(const K5Params p) {
  constexpr int kSlice = 32, kMaxRows = 20, kMaxCols = 3;
  constexpr double kTol = 1.0e-11;
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nW = blockDim.x >> 5;
  const int nChunks = __ldg(p.total) >> 5;
  double tmp[kSlice];                    // tmp(:,:,1), then the central difference of the new row
  double rowA[(kMaxCols + 1) * kSlice];  // dd(:,:,0:3,ii-1) and dd(:,:,0:3,ii), swapped every row
  double rowB[(kMaxCols + 1) * kSlice];
  double done[kSlice];                   // converged value per element
  bool conv[kSlice];
  for (int chunk = p.part + p.nParts * (blockIdx.x * nW + wIn); chunk < nChunks; chunk += p.nParts * gridDim.x * nW) {
    const K5Pair pr = k5LoadPair(p, chunk, lane);
    if (!pr.live) continue;  // the evaluator has no warp-level synchronisation: lanes may diverge
    const int ind = pr.ind, iAt1 = pr.iAt1, iAt2f = pr.iAt2f;
    const SpPairProg &pg = p.progs[pr.sp];
    const int nElem = pg.nOrb1 * pg.nOrb2;
    const double v0 = pr.v0, v1 = pr.v1, v2 = pr.v2;
    double g[3] = {0.0, 0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
      const SkTableDev tab = which ? pg.S : pg.H;
      // g = 2 (sum rho_1 . H' - sum (-wS) . S'); the S weights arrive negated (negWS)
      const double *__restrict__ wgt = (which ? p.negWS : p.DM) + ind;
      const double wSign = which ? -2.0 : 2.0;
      for (int iCart = 0; iCart < 3; ++iCart) {
        for (int e0 = 0; e0 < nElem; e0 += kSlice) {
          const int nS = min(kSlice, nElem - e0);
          // central difference (tmp(:,:,2) - tmp(:,:,1)) / (2 hh) of elements [e0, e0 + nS) -> tmp
          auto centralDiff = [&](double hh) {
            for (int kk = 1; kk <= 2; ++kk) {
              const double sh = -(double)(2 * kk - 3) * hh;  // kk = 1: +hh, kk = 2: -hh (nonscc.F90:539)
              const double d0 = iCart == 0 ? v0 + sh : v0, d1 = iCart == 1 ? v1 + sh : v1,
                           d2 = iCart == 2 ? v2 + sh : v2;
              const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
              const double ux = d0 / dist, uy = d1 / dist, uz = d2 / dist;
              skLaneEvalAny<LMAX>(
                  pg, tab, dist, ux, uy, uz,
                  [&](int first, int e, double v) {
                    const unsigned k = (unsigned)(first + e - e0);
                    if (k < (unsigned)kSlice) tmp[k] = kk == 1 ? v : (v - tmp[k]) / (2.0 * hh);
                  },
                  [](int, int) {});
            }
          };
          double hh = 0.125;
          double *prev = rowA, *cur = rowB;
          centralDiff(hh);
          for (int k = 0; k < nS; ++k) {
            cur[k] = tmp[k];
            for (int c = 1; c <= kMaxCols; ++c) cur[c * kSlice + k] = 0.0;
            conv[k] = false;
            done[k] = 0.0;
          }
          int ii = 0, mm = 0;
          bool all = false;
          while (ii < kMaxRows && !all) {
            hh = hh / 2.0;
            if (hh <= 1.220703125e-4) break;  // epsilon(1.0_dp)**0.25_dp = 2^-13
            ii += 1;
            double *t = prev; prev = cur; cur = t;
            centralDiff(hh);
            mm = min(ii, kMaxCols);
            all = true;
            for (int k = 0; k < nS; ++k) {
              if (conv[k]) continue;
              cur[k] = tmp[k];
              for (int kk = 1; kk <= mm; ++kk) {
                const double pw = (double)(1 << (2 * kk));
                cur[kk * kSlice + k] = (pw * cur[(kk - 1) * kSlice + k] - prev[(kk - 1) * kSlice + k]) / (pw - 1.0);
              }
              const double hi = cur[mm * kSlice + k];
              const double diff = fabs(cur[(mm - 1) * kSlice + k] - hi);
              if (diff <= kTol * hi) {
                conv[k] = true;
                done[k] = hi;
              } else {
                all = false;
              }
            }
          }
          // deriv(:,:,iCart) = dd(:,:,mm,ii)
          double acc = 0.0;
          for (int k = 0; k < nS; ++k)
            acc = fma(__ldg(wgt + e0 + k), conv[k] ? done[k] : cur[mm * kSlice + k], acc);
          g[iCart] = fma(wSign, acc, g[iCart]);
        }
      }
    }
    if (p.deriv && iAt1 != iAt2f) {
      atomicAdd(p.deriv + 3 * iAt1, g[0]);
      atomicAdd(p.deriv + 3 * iAt1 + 1, g[1]);
      atomicAdd(p.deriv + 3 * iAt1 + 2, g[2]);
      atomicAdd(p.deriv + 3 * iAt2f, -g[0]);
      atomicAdd(p.deriv + 3 * iAt2f + 1, -g[1]);
      atomicAdd(p.deriv + 3 * iAt2f + 2, -g[2]);
    }
    if (p.stress) {  // stress.F90:158-176; vect = -v
      const double fac = iAt1 != iAt2f ? -2.0 : -1.0;
      const double v[3] = {v0, v1, v2};
      for (int ii = 0; ii < 3; ++ii)
        for (int jj = 0; jj < 3; ++jj) atomicAdd(p.stress + jj + 3 * ii, fac * g[jj] * v[ii]);
    }
  }
}

template <int LMAX>
int launchRichardson(dftbp_b200_ctx *ctx, const K5Params &p, long long chunks) {
  const int grid = (int)std::max<long long>(1, std::min<long long>((chunks + 1) / 2, (long long)ctx->smCount * 8));
  k5_derivative_shift_richardson<LMAX><<<grid, 64, 0, ctx->stream>>>(p);
  return 0;
}

template <int LMAX, bool STRESS, int WARPS, int UNR>
int launchTypedS(dftbp_b200_ctx *ctx, const K5Params &p, long long items, int warpsMax, size_t perWarp) {
  const int warps = std::min(WARPS, warpsMax);
  if (warps < WARPS) return fail(ctx, DFTBP_B200_EINVAL, "derivative_shift: weight tile too large for %d warps", WARPS);
  const size_t smemBytes = perWarp * WARPS;
  static LaunchInfo info;
  int perSm = 1;  // resident CTAs per SM (registers and the weight tile decide)
  DFTBP_TRY(launchInfo(ctx, info, k5_derivative_shift<LMAX, STRESS, WARPS, UNR>, 32 * WARPS, smemBytes, &perSm));
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((items + WARPS - 1) / WARPS, (long long)ctx->smCount * std::max(perSm, 1)));
  k5_derivative_shift<LMAX, STRESS, WARPS, UNR><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p);
  return 0;
}

// Warps per CTA (one CTA per SM): the weight tile is 32 x mOrb^2 doubles per warp -- 21 KB for spd,
// 66 KB with f shells.  s/p/d: 8 warps with the full 255 registers beat 10 warps at 168 registers
// (119 vs 135 us on fec4096: the evaluator is register starved, not occupancy starved); unrolling
// the displacement loop by two does not help (136 us).
template <int LMAX, bool STRESS>
int launchByWarps(dftbp_b200_ctx *ctx, const K5Params &p, long long items, int warpsMax, size_t perWarp) {
  // (species with repeated shells have weight blocks of up to 16 x 16: fewer warps fit)
  if constexpr (LMAX <= 2) {
    if (warpsMax >= 8) return launchTypedS<LMAX, STRESS, 8, 1>(ctx, p, items, warpsMax, perWarp);
    if (warpsMax >= 4) return launchTypedS<LMAX, STRESS, 4, 1>(ctx, p, items, warpsMax, perWarp);
  }
  if (warpsMax >= 2) return launchTypedS<LMAX, STRESS, 2, 1>(ctx, p, items, warpsMax, perWarp);
  return launchTypedS<LMAX, STRESS, 1, 1>(ctx, p, items, warpsMax, perWarp);
}

template <int LMAX>
int launchTyped(dftbp_b200_ctx *ctx, const K5Params &p, long long items, int warpsMax, size_t perWarp) {
  if (p.stress) return launchByWarps<LMAX, true>(ctx, p, items, warpsMax, perWarp);
  return launchByWarps<LMAX, false>(ctx, p, items, warpsMax, perWarp);
}

}  // namespace

namespace {

template <int NSPIN>
int launchWeightsFlatT(dftbp_b200_ctx *ctx, const K5Params &p, const BlockDesc *desc, long long nWork,
                       const int *groupFirst, const double *EDM, const double *diag, const int *flag, double *negWS,
                       const double *shift, int selfImages, int spanLo, int spanHi) {
  constexpr int NIN = 1 + NSPIN, WARPS = flat::warpsFor(NIN);
  const size_t smemBytes = sizeof(flat::Smem<NIN, WARPS>);
  static LaunchInfo info;
  DFTBP_TRY(launchInfo(ctx, info, k5_force_weights_flat<NSPIN>, 32 * WARPS, smemBytes, nullptr));
  const long long nSpans = std::max(0, spanHi - spanLo);
  const int grid = (int)std::max<long long>(1, std::min<long long>((nSpans + WARPS - 1) / WARPS, ctx->smCount));
  k5_force_weights_flat<NSPIN><<<grid, 32 * WARPS, smemBytes, ctx->stream>>>(p, desc, nWork, groupFirst, EDM, diag, flag,
                                                                            negWS, shift, selfImages, spanLo, spanHi, ctx->hShiftHint);
  return 0;
}

int launchWeightsFlat(dftbp_b200_ctx *ctx, int nSpin, const K5Params &p, const BlockDesc *desc, long long nWork,
                      const int *groupFirst, const double *EDM, const double *diag, const int *flag, double *negWS,
                      const double *shift, int selfImages, int spanLo, int spanHi) {
  if (nSpin == 1) return launchWeightsFlatT<1>(ctx, p, desc, nWork, groupFirst, EDM, diag, flag, negWS, shift, selfImages, spanLo, spanHi);
  if (nSpin == 2) return launchWeightsFlatT<2>(ctx, p, desc, nWork, groupFirst, EDM, diag, flag, negWS, shift, selfImages, spanLo, spanHi);
  return launchWeightsFlatT<4>(ctx, p, desc, nWork, groupFirst, EDM, diag, flag, negWS, shift, selfImages, spanLo, spanHi);
}

// derivative_shift (deriv != nullptr) and / or the stress tensor (st != nullptr) from one pass over the pairs
int forcesImpl(dftbp_b200_ctx *ctx, double *deriv, double *st, double cellVol, const double *DM, int nSpin,
               const double *EDM, const double *shift, int diffType, double delta) {
  if (!ctx) return DFTBP_B200_EINVAL;
  const Geometry &g = ctx->geo;
  DFTBP_CHECK(ctx, g.nAtom > 0, "derivative_shift: no geometry set");
  DFTBP_CHECK(ctx, (deriv || st) && DM && EDM, "derivative_shift: null argument");
  DFTBP_CHECK(ctx, nSpin == 1 || nSpin == 2 || nSpin == 4, "derivative_shift: nSpin must be 1, 2 or 4");
  DFTBP_CHECK(ctx, diffType == DFTBP_B200_DIFF_FINITE || diffType == DFTBP_B200_DIFF_RICHARDSON,
              "derivative_shift: invalid differentiator type");
  DFTBP_CHECK(ctx, !st || cellVol > 0.0, "stress: cell volume must be positive");
  DFTBP_TRY(syncPrograms(ctx));
  const SortedPairs *sched;
  DFTBP_TRY(ensureSortedPairs(ctx, true, &sched));
  K5Params p;
  p.progs = ctx->dProgs; p.recs = sched->recs; p.atoms = sched->atoms; p.total = sched->total;
  p.ld = g.mNeigh + 1; p.nSpecies = ctx->nSpecies;
  p.mOrb = ctx->mOrb; p.nAtom = g.nAtom; p.nSpin = nSpin; p.sparseSize = g.sparseSize;
  // nonscc.F90:43: deltaXDiff = epsilon(1.0_dp)**0.25_dp
  p.delta = delta > 0.0 ? delta : 1.220703125e-4;
  p.deriv = nullptr; p.stress = nullptr;
  // a share of the pairs (set_pair_partition) is taken when the schedule is built (pair_sort.cu): the
  // kernels then walk through all chunks of the schedule they are given
  p.part = 0; p.nParts = 1;
  // the non-SCC routine takes DM(:) only; the block routine DM(:, nSpin)
  const int nDm = shift ? nSpin : 1;
  DFTBP_TRY(devIn(ctx, kSlotA, DM, (size_t)g.sparseSize * nDm, &p.DM));
  DFTBP_TRY(devIn(ctx, kSlotB, EDM, (size_t)g.sparseSize, &p.negWS));
  const double *dEDM = p.negWS, *dShift = nullptr;
  if (shift)
    DFTBP_TRY(devIn(ctx, kSlotC, shift, (size_t)ctx->mOrb * ctx->mOrb * g.nAtom * nSpin, &dShift));
  if (deriv) {
    DFTBP_TRY(devOut(ctx, kSlotD, deriv, (size_t)3 * g.nAtom, false, &p.deriv));
    DFTBP_CUDA(ctx, cudaMemsetAsync(p.deriv, 0, sizeof(double) * 3 * g.nAtom, ctx->stream));
  }
  double *dSt = nullptr;
  if (st) {
    void *acc;
    DFTBP_TRY(stagingGet(ctx, kSlotF, sizeof(double) * 9, &acc));
    p.stress = (double *)acc;
    DFTBP_CUDA(ctx, cudaMemsetAsync(p.stress, 0, sizeof(double) * 9, ctx->stream));
    DFTBP_TRY(devOut(ctx, kSlotG, st, 9, false, &dSt));
  }
  if (g.nWorkLoad > 0) {
    ScopedTimer timer(ctx, kTimerForceCalc);
    if (shift) {
      // plain SCC / collinear spin shifts are diagonal blocks: detect it, then form -wS once per block
      int *flag = ctx->dInfo + 1;
      DFTBP_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
      double *dDiag;
      DFTBP_TRY(shiftDiagBuffer(ctx, (long long)ctx->mOrb * g.nAtom * nSpin, &dDiag));
      k5_shift_structure<<<std::min(ctx->smCount * 8, (g.nAtom * nSpin + 7) / 8), 256, 0, ctx->stream>>>(
          dShift, ctx->mOrb, g.nAtom * nSpin, flag, dDiag);
      void *scratch;
      DFTBP_TRY(stagingGet(ctx, kSlotE, sizeof(double) * (size_t)g.sparseSize, &scratch));
      // both instantiations are launched; the one that does not match the device flag returns at once
      const long long myChunks = (((g.nWorkLoad + 31) / 32 + ctx->nSpecies * ctx->nSpecies) + ctx->pairParts - 1) / ctx->pairParts;
      const int wGrid = gridFor(ctx, myChunks, 8, 6);
      const int self = st ? 1 : 0;
      // diagonal shifts: the blocks in memory order (descriptors of the work list + their schedule slots)
      const int dGrid = gridFor(ctx, (g.nWorkLoad + kBlocksPerChunk - 1) / kBlocksPerChunk, 8, 4);
      // one partition and a sorted list of not too small blocks: the TMA span pipeline (flat_stream.cuh)
      const bool useFlat = g.groupFirstLoad != nullptr && (long long)ctx->mOrb * g.nAtom * nSpin < (1LL << 31) &&
                           ((long long)ctx->loadInd.size() == g.nWorkLoad + 1 || (ctx->loadInd.empty() && g.coversAll)) &&
                           tuneInt("DFTBP_B200_K5_FLAT", 1) != 0;
      if (useFlat) {
        // one launch: the kernel takes the full-block path itself when the device flag says so.  A share
        // of the pairs (set_pair_partition) is a contiguous range of the work list = of the sparse arrays:
        // only the spans it touches are streamed.
        const long long wLo = shareBegin(g.nWorkLoad, ctx->pairPart, ctx->pairParts),
                        wHi = shareBegin(g.nWorkLoad, ctx->pairPart + 1, ctx->pairParts);
        int eLo, eHi;
        DFTBP_TRY(loadElementRange(ctx, wLo, wHi, &eLo, &eHi));
        const int spanLo = eLo / flat::kSpan;
        const int spanHi = wHi > wLo ? (eHi + flat::kSpan - 1) / flat::kSpan : spanLo;
        const bool fullLastTime = *ctx->hShiftHint == 1;
        DFTBP_TRY(launchWeightsFlat(ctx, nSpin, p, g.descLoad, g.nWorkLoad, g.groupFirstLoad, dEDM, dDiag, flag, (double *)scratch,
                                    dShift, self | (fullLastTime ? 2 : 0), spanLo, spanHi));
        if (fullLastTime) {
          k5_force_weights_sched<1, true><<<wGrid, 256, 0, ctx->stream>>>(p, dEDM, dShift, flag, (double *)scratch, self);
          ctx->launches++;
        }
        ctx->launches += 2;
      } else {
        if (nSpin == 1) k5_force_weights_diag<1><<<dGrid, 256, 0, ctx->stream>>>(p, g.descLoad, sched->slotOf, g.nWorkLoad, dEDM, dShift, flag, (double *)scratch, self);
        else if (nSpin == 2) k5_force_weights_diag<2><<<dGrid, 256, 0, ctx->stream>>>(p, g.descLoad, sched->slotOf, g.nWorkLoad, dEDM, dShift, flag, (double *)scratch, self);
        else k5_force_weights_diag<4><<<dGrid, 256, 0, ctx->stream>>>(p, g.descLoad, sched->slotOf, g.nWorkLoad, dEDM, dShift, flag, (double *)scratch, self);
        k5_force_weights_sched<1, true><<<wGrid, 256, 0, ctx->stream>>>(p, dEDM, dShift, flag, (double *)scratch, self);
        ctx->launches += 3;
      }
      p.negWS = (const double *)scratch;
    }
    p.pitch = (ctx->mOrb * ctx->mOrb) | 1;
    const size_t perWarp = sizeof(double) * 32 * (size_t)p.pitch;
    // 10 warps x <= 204 registers fill the register file; the weight tile usually limits first
    const int warps = (int)std::min<size_t>(10, (ctx->smemOptin - 2048) / perWarp);
    DFTBP_CHECK(ctx, warps >= 1, "derivative_shift: weight tile does not fit into shared memory");
    const long long items = 2 * ((((g.nWorkLoad + 31) / 32 + ctx->nSpecies * ctx->nSpecies) + ctx->pairParts - 1) / ctx->pairParts);
    int rc;
    if (diffType == DFTBP_B200_DIFF_RICHARDSON) {
      const long long chunks = items / 2;
      if (ctx->lMaxAll <= 1) rc = launchRichardson<1>(ctx, p, chunks);
      else if (ctx->lMaxAll == 2) rc = launchRichardson<2>(ctx, p, chunks);
      else rc = launchRichardson<3>(ctx, p, chunks);
    } else if (ctx->lMaxAll <= 1) rc = launchTyped<1>(ctx, p, items, warps, perWarp);
    else if (ctx->lMaxAll == 2) rc = launchTyped<2>(ctx, p, items, warps, perWarp);
    else rc = launchTyped<3>(ctx, p, items, warps, perWarp);
    if (rc) return rc;
    ctx->launches++;
  }
  if (st) {
    k5_stress_scale<<<1, 32, 0, ctx->stream>>>(p.stress, cellVol, dSt);
    ctx->launches++;
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_derivative_shift(dftbp_b200_ctx *ctx, double *deriv, const double *DM, int nSpin, const double *EDM,
                                const double *shift, int diffType, double delta) {
  if (ctx && !deriv) return fail(ctx, DFTBP_B200_EINVAL, "derivative_shift: deriv is null");
  return forcesImpl(ctx, deriv, nullptr, 0.0, DM, nSpin, EDM, shift, diffType, delta);
}

int dftbp_b200_stress(dftbp_b200_ctx *ctx, double *st, const double *DM, int nSpin, const double *EDM,
                      const double *shift, int diffType, double delta, double cellVol) {
  if (ctx && !st) return fail(ctx, DFTBP_B200_EINVAL, "stress: st is null");
  return forcesImpl(ctx, nullptr, st, cellVol, DM, nSpin, EDM, shift, diffType, delta);
}

}  // extern "C"

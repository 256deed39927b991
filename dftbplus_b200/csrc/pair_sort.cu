// pair_sort.cu -- device-side schedule of the lane-per-pair kernels (K1 pair build, K5 forces).
//
// Those kernels give every THREAD one neighbour pair (like the reference's OpenMP loop gives every
// CPU thread one atom, dftb/nonscc.F90:399-422), so the 32 lanes of a warp must (a) belong to one
// ordered species pair -- the shell loops and the generated rotation code are then warp-uniform --
// and (b) sit at neighbouring distances, so that the 8-row windows of the SK tables they
// interpolate from fall into the same one or two cache lines (broadcast loads instead of gathers).
// A counting sort by key = (species pair, distance bin) gives both; it runs on the device in two
// small launches (histogram, whose last CTA scans it with per-bucket padding to a multiple of 32; scatter;
// histogram and scatter aggregate in shared memory, one global atomic per CTA and occupied bin)
// and is redone whenever the coordinates change.  The order inside a bin is arbitrary (atomic
// cursors): every pair is evaluated independently and written to its own iPair offset, so the
// results do not depend on it.
#include "common.cuh"

using namespace dftbp_b200;

namespace {

constexpr int kSortThreads = 1024;  // one big CTA aggregates many items per histogram bin in shared memory

// distance bins per species pair over [0, SK cutoff]: as fine as the shared-memory histogram allows
// (1024 bins = 0.011 bohr at an 11 bohr cutoff, half a table row; 64 bins still keep a warp within
// a few rows)
int binsFor(int nBuckets) { return nBuckets <= 8 ? 1024 : nBuckets <= 32 ? 256 : nBuckets <= 128 ? 64 : 16; }

// key of work item `flat` and its record (bond vector, sparse offset, species-pair program); on-site
// blocks (iNeigh = 0) get key -1 and a record flagged kRecOnsite with a zero bond vector
__device__ __forceinline__ int pairKey(int flat, int ld, int nSpecies, int nBins, double binScale,
                                       const double *__restrict__ coords, const int *__restrict__ species,
                                       const int *__restrict__ iNeighbour, const int *__restrict__ iPair, PairRec &r) {
  const int iAt1 = flat / ld, iNeigh = flat - iAt1 * ld;
  const int sp1 = __ldg(species + iAt1) - 1;
  r.ind = __ldg(iPair + flat);
  if (iNeigh == 0) {  // on-site blocks are not pairs
    r.vx = r.vy = r.vz = 0.0;
    r.sp = (sp1 + nSpecies * sp1) | kRecOnsite;
    return -1;
  }
  const int iAt2 = __ldg(iNeighbour + flat) - 1;
  const int sp2 = __ldg(species + iAt2) - 1;
  // vect(:) = coords(:,iAt2) - coords(:,iAt1), nonscc.F90:413
  r.vx = __ldg(coords + 3 * iAt2) - __ldg(coords + 3 * iAt1);
  r.vy = __ldg(coords + 3 * iAt2 + 1) - __ldg(coords + 3 * iAt1 + 1);
  r.vz = __ldg(coords + 3 * iAt2 + 2) - __ldg(coords + 3 * iAt1 + 2);
  r.sp = sp1 + nSpecies * sp2;
  const double d = sqrt(r.vx * r.vx + r.vy * r.vy + r.vz * r.vz);
  const int bin = min(nBins - 1, max(0, (int)(d * binScale)));
  return (sp1 + nSpecies * sp2) * nBins + bin;
}

// one CTA: exclusive scan bucket by bucket (nBins entries each), every bucket starts on a multiple of 32
// The padding slots behind every bucket get ind = -1 / item = -1 here (at most 31 per bucket), so the
// schedule arrays need no fill pass.  Runs in the last CTA of ps_keys to finish (ticket counter): the
// histogram is complete then, and a launch of its own would cost more than the scan.
__device__ __forceinline__ void scanBuckets(int *__restrict__ hist, int nBuckets, int nBins, int *__restrict__ total,
                                            int *__restrict__ items, PairRec *__restrict__ recs, int *warpSum,
                                            int *baseSh) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) *baseSh = 0;
  __syncthreads();
  for (int b = 0; b < nBuckets; ++b) {
    const int v = t < nBins ? __ldcg(hist + b * nBins + t) : 0;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warpSum[w] = x;
    __syncthreads();
    if (w == 0) {
      int s = warpSum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, off);
        if (lane >= off) s += y;
      }
      warpSum[lane] = s;  // inclusive sums of the warps
    }
    __syncthreads();
    const int base = *baseSh;
    const int incl = x + (w > 0 ? warpSum[w - 1] : 0);
    if (t < nBins) hist[b * nBins + t] = base + incl - v;
    __syncthreads();
    const int endB = base + warpSum[31];  // first slot behind the bucket
    if (t < 32 && endB + t < ((endB + 31) & ~31)) {
      items[endB + t] = -1;
      recs[endB + t].ind = -1;
    }
    if (t == 1023) *baseSh = (base + incl + 31) & ~31;
    __syncthreads();
  }
  if (t == 0) *total = *baseSh;
}

// pass 1: key of every work item + histogram.  Each CTA owns a contiguous slab of the work list and
// counts in shared memory first, so only one global atomic per (CTA, occupied bin) remains --
// crystalline systems put thousands of pairs into the same few bins.
__global__ void __launch_bounds__(kSortThreads) ps_keys(const int *__restrict__ work, long long nWork, int ld,
                                                        int nSpecies, int nBins, int nKeys,
                                                        const double *__restrict__ coords,
                                                        const int *__restrict__ species,
                                                        const int *__restrict__ iNeighbour,
                                                        const int *__restrict__ iPair, double binScale,
                                                        int *__restrict__ keys, int *__restrict__ hist,
                                                        PairRec *__restrict__ recsMem, int *__restrict__ ticket,
                                                        int nBuckets, int *__restrict__ total,
                                                        int *__restrict__ items, PairRec *__restrict__ recs,
                                                        long long shareLo, long long shareHi) {
  extern __shared__ int sh[];
  __shared__ int warpSum[32];
  __shared__ int baseSh, lastSh;
  for (int k = threadIdx.x; k < nKeys; k += kSortThreads) sh[k] = 0;
  __syncthreads();
  const long long per = (nWork + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = min(nWork, lo + per);
  for (long long i = lo + threadIdx.x; i < hi; i += kSortThreads) {
    PairRec r;
    // a schedule of one share of the pairs (dftbp_b200_set_pair_partition): a contiguous range of the
    // work list -- a property of the pair, not of the (non-deterministic) order inside a bin, so that
    // the shares of different GPUs are disjoint and complete
    if (i < shareLo || i >= shareHi) {
      keys[i] = -1;
      continue;
    }
    const int key = pairKey(__ldg(work + i), ld, nSpecies, nBins, binScale, coords, species, iNeighbour, iPair, r);
    keys[i] = key;
    recsMem[i] = r;  // the same record in the order of the work list (= the order of the sparse arrays)
    if (key >= 0) atomicAdd(sh + key, 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nKeys; k += kSortThreads) {
    const int c = sh[k];
    if (c) atomicAdd(hist + k, c);
  }
  // the last CTA to get here scans the complete histogram
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) lastSh = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!lastSh) return;
  __threadfence();
  scanBuckets(hist, nBuckets, nBins, total, items, recs, warpSum, &baseSh);
  if (threadIdx.x == 0) *ticket = 0;  // ready for the next sort
}

// pass 2: every CTA counts its slab again (from the stored keys), reserves one contiguous range
// per occupied bin with a single global atomic, and places its items with shared-memory atomics.
// Together with the flat id it writes the pair's record (bond vector, sparse offset, species-pair
// program, the two atoms): the consumers read one coalesced 32-byte record per lane instead of
// walking the index chain.
__global__ void __launch_bounds__(kSortThreads) ps_scatter(const int *__restrict__ work, long long nWork,
                                                           int nKeys, const int *__restrict__ keys,
                                                           int *__restrict__ cursors, int *__restrict__ items,
                                                           PairRec *__restrict__ recs, int2 *__restrict__ atoms,
                                                           int ld, const PairRec *__restrict__ recsMem,
                                                           const int *__restrict__ iNeighbour,
                                                           const int *__restrict__ img2CentCell,
                                                           int *__restrict__ slotOf) {
  extern __shared__ int sh[];
  for (int k = threadIdx.x; k < nKeys; k += kSortThreads) sh[k] = 0;
  __syncthreads();
  const long long per = (nWork + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = min(nWork, lo + per);
  for (long long i = lo + threadIdx.x; i < hi; i += kSortThreads) {
    const int key = keys[i];
    if (key >= 0) atomicAdd(sh + key, 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nKeys; k += kSortThreads) {
    const int c = sh[k];
    if (c) sh[k] = atomicAdd(cursors + k, c);  // count -> first position of this CTA's range
  }
  __syncthreads();
  for (long long i = lo + threadIdx.x; i < hi; i += kSortThreads) {
    const int key = keys[i];
    if (key < 0) {
      slotOf[i] = -1;  // on-site block: in no chunk of the schedule
      continue;
    }
    const int flat = __ldg(work + i);
    const int iAt1 = flat / ld;
    const int iAt2 = __ldg(iNeighbour + flat) - 1;
    const int4 *src = reinterpret_cast<const int4 *>(recsMem + i);
    const int4 lo = __ldg(src), hi = __ldg(src + 1);
    const int pos = atomicAdd(sh + key, 1);
    items[pos] = (int)i;  // position in the work list: row of the pair in per-pair exchange buffers
    slotOf[i] = pos;      // and the inverse: where the schedule holds work-list entry i
    int4 *dst = reinterpret_cast<int4 *>(recs + pos);
    dst[0] = lo;
    dst[1] = hi;
    atoms[pos] = make_int2(iAt1, __ldg(img2CentCell + iAt2) - 1);
  }
}

// New coordinates on an unchanged neighbour list (dftbp_b200_update_coords): the pairs, their species
// pairs and their sparse offsets are the same, only the bond vectors change.  The order of the schedule
// (species pair: still exact; distance bin: a locality heuristic of the table windows) is kept and the
// records are rewritten in place, both in work-list order and at the slot of every pair.
__global__ void __launch_bounds__(256) ps_refresh(const int *__restrict__ work, long long nWork, int ld, int nSpecies,
                                                  const double *__restrict__ coords, const int *__restrict__ species,
                                                  const int *__restrict__ iNeighbour, const int *__restrict__ iPair,
                                                  const int *__restrict__ slotOf, PairRec *__restrict__ recsMem,
                                                  PairRec *__restrict__ recs) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nWork; i += (long long)gridDim.x * blockDim.x) {
    PairRec r;
    pairKey(__ldg(work + i), ld, nSpecies, 1, 0.0, coords, species, iNeighbour, iPair, r);
    int4 *m = reinterpret_cast<int4 *>(recsMem + i);
    const int4 lo = *reinterpret_cast<const int4 *>(&r), hi = *(reinterpret_cast<const int4 *>(&r) + 1);
    m[0] = lo;
    m[1] = hi;
    const int slot = __ldg(slotOf + i);
    if (slot >= 0) {
      int4 *d = reinterpret_cast<int4 *>(recs + slot);
      d[0] = lo;
      d[1] = hi;
    }
  }
}

template <class T>
int growTo(dftbp_b200_ctx *ctx, T **p, long long *cap, long long need) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (size_t)need);
  if (e != cudaSuccess)
    return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc(%lld) for the pair schedule failed: %s",
                (long long)(sizeof(T) * need), cudaGetErrorString(e));
  *cap = need;
  return 0;
}

}  // namespace

namespace dftbp_b200 {

void freeSortedPairs(dftbp_b200_ctx *g) {
  for (SortedPairs *sp : {&g->sortLoad, &g->sortMine}) {
    cudaFree(sp->items);
    cudaFree(sp->recs);
    cudaFree(sp->recsMem);
    cudaFree(sp->slotOf);
    cudaFree(sp->atoms);
    cudaFree(sp->total);
  }
  cudaFree(g->sortKeys);
  cudaFree(g->sortBins);
  g->sortMine = g->sortLoad = SortedPairs();
  g->sortKeys = g->sortBins = nullptr;
  g->sortKeysCap = g->sortBinsCap = 0;
}

int ensureSortedPairs(dftbp_b200_ctx *ctx, bool forLoad, const SortedPairs **out) {
  Geometry &g = ctx->geo;
  // one rank and one share: both work lists are the same, and so are the schedules
  const bool shared = g.workLoad == g.workMine && ctx->pairParts == 1;
  SortedPairs &sp = (forLoad && !shared) ? ctx->sortLoad : ctx->sortMine;
  const int part = (forLoad && !shared) ? ctx->pairPart : 0, nParts = (forLoad && !shared) ? ctx->pairParts : 1;
  const long long nWorkAllOfList = (forLoad && !shared) ? g.nWorkLoad : g.nWorkMine;
  const long long shareLo = shareBegin(nWorkAllOfList, part, nParts), shareHi = shareBegin(nWorkAllOfList, part + 1, nParts);
  *out = &sp;
  if (sp.valid && !sp.stale) return 0;
  // kRefreshes updates of the coordinates reuse the order: at MD-sized displacements (~0.01 bohr per step)
  // the relative distances move by a few table rows in that time, well inside the shared-memory window of
  // k1_interp; chunks that outgrow it take the global-memory path (slower, same result) until the next sort
  constexpr int kRefreshes = 16;
  if (sp.valid && sp.stale && sp.refreshes < kRefreshes && tuneInt("DFTBP_B200_SORT_REFRESH", 1) != 0) {
    const int *workR = (forLoad && !shared) ? g.workLoad : g.workMine;
    const long long nWorkR = (forLoad && !shared) ? g.nWorkLoad : g.nWorkMine;
    if (nWorkR > 0) {
      const int grid = (int)std::max<long long>(1, std::min<long long>((nWorkR + 255) / 256, (long long)ctx->smCount * 8));
      ps_refresh<<<grid, 256, 0, ctx->stream>>>(workR, nWorkR, g.mNeigh + 1, ctx->nSpecies, g.coords, g.species,
                                                g.iNeighbour, g.iPair, sp.slotOf, sp.recsMem, sp.recs);
      ctx->launches++;
    }
    sp.stale = false;
    sp.refreshes++;
    return 0;
  }
  const int *work = (forLoad && !shared) ? g.workLoad : g.workMine;
  const long long nWork = (forLoad && !shared) ? g.nWorkLoad : g.nWorkMine;
  const int nBuckets = ctx->nSpecies * ctx->nSpecies;
  const int nBins = binsFor(nBuckets);
  const long long nBins64 = (long long)nBuckets * nBins;
  DFTBP_CHECK(ctx, nBins64 * sizeof(int) <= ctx->smemOptin - 1024, "too many species pairs for the pair schedule");
  const int nKeys = (int)nBins64;
  const long long cap = ((nWork + 31) & ~31LL) + 32LL * nBuckets + 32;
  if (sp.capacity < cap) {
    cudaFree(sp.items); cudaFree(sp.recs); cudaFree(sp.atoms);
    sp.items = nullptr; sp.recs = nullptr; sp.atoms = nullptr;
    sp.capacity = 0;
    long long c = 0, c2 = 0, c3 = 0;
    DFTBP_TRY(growTo(ctx, &sp.items, &c, cap));
    DFTBP_TRY(growTo(ctx, &sp.recs, &c2, cap));
    DFTBP_TRY(growTo(ctx, &sp.atoms, &c3, cap));
    sp.capacity = c;
  }
  if (!sp.total) DFTBP_CUDA(ctx, cudaMalloc((void **)&sp.total, sizeof(int)));
  DFTBP_TRY(growTo(ctx, &sp.recsMem, &sp.recsMemCap, std::max(nWork, 1LL)));
  DFTBP_TRY(growTo(ctx, &sp.slotOf, &sp.slotOfCap, std::max(nWork, 1LL)));
  DFTBP_TRY(growTo(ctx, &ctx->sortKeys, &ctx->sortKeysCap, std::max(nWork, 1LL)));
  DFTBP_TRY(growTo(ctx, &ctx->sortBins, &ctx->sortBinsCap, nBins64 + 1));  // + the ticket of ps_keys
  DFTBP_CUDA(ctx, cudaMemsetAsync(ctx->sortBins, 0, sizeof(int) * (nBins64 + 1), ctx->stream));
  // padding slots (item = -1, ind = -1: the consumers then shadow lane 0 of the chunk) are written by ps_scan
  const double binScale = (double)nBins / (ctx->cutoff > 0.0 ? ctx->cutoff : 1.0);
  if (nWork > 0) {
    // slabs of at least 512 items per CTA (aggregation in shared memory), at most one CTA per SM: the
    // kernels are chains of dependent loads, so all SMs should share them
    const int grid = (int)std::max<long long>(1, std::min<long long>(ctx->smCount, (nWork + 511) / 512));
    const size_t smemBytes = sizeof(int) * (size_t)nKeys;
    static LaunchInfo infoKeys, infoScatter;
    DFTBP_TRY(launchInfo(ctx, infoKeys, ps_keys, kSortThreads, smemBytes, nullptr));
    DFTBP_TRY(launchInfo(ctx, infoScatter, ps_scatter, kSortThreads, smemBytes, nullptr));
    ps_keys<<<grid, kSortThreads, smemBytes, ctx->stream>>>(work, nWork, g.mNeigh + 1, ctx->nSpecies, nBins, nKeys,
                                                           g.coords, g.species, g.iNeighbour, g.iPair, binScale,
                                                           ctx->sortKeys, ctx->sortBins, sp.recsMem,
                                                           ctx->sortBins + nBins64, nBuckets, sp.total, sp.items, sp.recs,
                                                           shareLo, shareHi);
    ps_scatter<<<grid, kSortThreads, smemBytes, ctx->stream>>>(work, nWork, nKeys, ctx->sortKeys, ctx->sortBins,
                                                              sp.items, sp.recs, sp.atoms, g.mNeigh + 1, sp.recsMem,
                                                              g.iNeighbour, g.img2CentCell, sp.slotOf);
    ctx->launches += 2;
  } else {
    DFTBP_CUDA(ctx, cudaMemsetAsync(sp.total, 0, sizeof(int), ctx->stream));
  }
  sp.valid = true;
  sp.stale = false;
  sp.refreshes = 0;
  return 0;
}

}  // namespace dftbp_b200

// pair_sort.cu -- device-side schedule of the lane-per-pair kernels (K1 pair build, K5 forces).
//
// Those kernels give every THREAD one neighbour pair (like the reference's OpenMP loop gives every
// CPU thread one atom, dftb/nonscc.F90:399-422), so the 32 lanes of a warp must (a) belong to one
// ordered species pair -- the shell loops and the generated rotation code are then warp-uniform --
// and (b) sit at neighbouring distances, so that the 8-row windows of the SK tables they
// interpolate from fall into the same one or two cache lines (broadcast loads instead of gathers).
// A counting sort by key = (species pair, distance bin) gives both; it runs on the device in three
// small launches (histogram, single-CTA scan with per-bucket padding to a multiple of 32, scatter)
// and is redone whenever the coordinates change.  The order inside a bin is arbitrary (atomic
// cursors): every pair is evaluated independently and written to its own iPair offset, so the
// results do not depend on it.
#include "common.cuh"

using namespace dftbp_b200;

namespace {

constexpr int kSortBins = 1024;  // distance bins per species pair over [0, SK cutoff]

__global__ void __launch_bounds__(256) ps_keys(const int *__restrict__ work, long long nWork, int ld,
                                               int nSpecies, const double *__restrict__ coords,
                                               const int *__restrict__ species,
                                               const int *__restrict__ iNeighbour, double binScale,
                                               int *__restrict__ keys, int *__restrict__ hist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nWork) return;
  const int flat = __ldg(work + i);
  const int iAt1 = flat / ld, iNeigh = flat - iAt1 * ld;
  if (iNeigh == 0) {  // on-site blocks are not pairs
    keys[i] = -1;
    return;
  }
  const int iAt2 = __ldg(iNeighbour + flat) - 1;
  const int sp1 = __ldg(species + iAt1) - 1, sp2 = __ldg(species + iAt2) - 1;
  const double vx = __ldg(coords + 3 * iAt2) - __ldg(coords + 3 * iAt1);
  const double vy = __ldg(coords + 3 * iAt2 + 1) - __ldg(coords + 3 * iAt1 + 1);
  const double vz = __ldg(coords + 3 * iAt2 + 2) - __ldg(coords + 3 * iAt1 + 2);
  const double r = sqrt(vx * vx + vy * vy + vz * vz);
  const int bin = min(kSortBins - 1, max(0, (int)(r * binScale)));
  const int key = (sp1 + nSpecies * sp2) * kSortBins + bin;
  keys[i] = key;
  atomicAdd(hist + key, 1);
}

// one CTA of kSortBins threads: exclusive scan bucket by bucket, every bucket starts on a multiple of 32
__global__ void __launch_bounds__(kSortBins) ps_scan(int *__restrict__ hist, int nBuckets, int *__restrict__ total) {
  __shared__ int warpSum[32];
  __shared__ int baseSh;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) baseSh = 0;
  __syncthreads();
  for (int b = 0; b < nBuckets; ++b) {
    const int v = hist[b * kSortBins + t];
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
    const int base = baseSh;
    const int incl = x + (w > 0 ? warpSum[w - 1] : 0);
    hist[b * kSortBins + t] = base + incl - v;
    __syncthreads();
    if (t == kSortBins - 1) baseSh = (base + incl + 31) & ~31;
    __syncthreads();
  }
  if (t == 0) *total = baseSh;
}

__global__ void __launch_bounds__(256) ps_scatter(const int *__restrict__ work, long long nWork,
                                                  const int *__restrict__ keys, int *__restrict__ cursors,
                                                  int *__restrict__ items) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nWork) return;
  const int key = keys[i];
  if (key < 0) return;
  const int pos = atomicAdd(cursors + key, 1);
  items[pos] = __ldg(work + i);
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
  cudaFree(g->sortLoad.items);
  cudaFree(g->sortLoad.total);
  cudaFree(g->sortMine.items);
  cudaFree(g->sortMine.total);
  cudaFree(g->sortKeys);
  cudaFree(g->sortBins);
  g->sortMine = g->sortLoad = SortedPairs();
  g->sortKeys = g->sortBins = nullptr;
  g->sortKeysCap = g->sortBinsCap = 0;
}

int ensureSortedPairs(dftbp_b200_ctx *ctx, bool forLoad, const SortedPairs **out) {
  Geometry &g = ctx->geo;
  const bool shared = g.workLoad == g.workMine;  // one rank: both work lists are the same
  SortedPairs &sp = (forLoad && !shared) ? ctx->sortLoad : ctx->sortMine;
  *out = &sp;
  if (sp.valid) return 0;
  const int *work = (forLoad && !shared) ? g.workLoad : g.workMine;
  const long long nWork = (forLoad && !shared) ? g.nWorkLoad : g.nWorkMine;
  const int nBuckets = ctx->nSpecies * ctx->nSpecies;
  const long long nBins = (long long)nBuckets * kSortBins;
  const long long cap = ((nWork + 31) & ~31LL) + 32LL * nBuckets + 32;
  if (sp.capacity < cap) {
    if (sp.items) cudaFree(sp.items);
    sp.items = nullptr;
    long long c = 0;
    DFTBP_TRY(growTo(ctx, &sp.items, &c, cap));
    sp.capacity = c;
  }
  if (!sp.total) DFTBP_CUDA(ctx, cudaMalloc((void **)&sp.total, sizeof(int)));
  DFTBP_TRY(growTo(ctx, &ctx->sortKeys, &ctx->sortKeysCap, std::max(nWork, 1LL)));
  DFTBP_TRY(growTo(ctx, &ctx->sortBins, &ctx->sortBinsCap, nBins));
  DFTBP_CUDA(ctx, cudaMemsetAsync(ctx->sortBins, 0, sizeof(int) * nBins, ctx->stream));
  DFTBP_CUDA(ctx, cudaMemsetAsync(sp.items, 0xFF, sizeof(int) * sp.capacity, ctx->stream));
  const double binScale = (double)kSortBins / (ctx->cutoff > 0.0 ? ctx->cutoff : 1.0);
  if (nWork > 0) {
    const int grid = (int)((nWork + 255) / 256);
    ps_keys<<<grid, 256, 0, ctx->stream>>>(work, nWork, g.mNeigh + 1, ctx->nSpecies, g.coords, g.species,
                                           g.iNeighbour, binScale, ctx->sortKeys, ctx->sortBins);
    ps_scan<<<1, kSortBins, 0, ctx->stream>>>(ctx->sortBins, nBuckets, sp.total);
    ps_scatter<<<grid, 256, 0, ctx->stream>>>(work, nWork, ctx->sortKeys, ctx->sortBins, sp.items);
    ctx->launches += 3;
  } else {
    DFTBP_CUDA(ctx, cudaMemsetAsync(sp.total, 0, sizeof(int), ctx->stream));
  }
  sp.valid = true;
  return 0;
}

}  // namespace dftbp_b200

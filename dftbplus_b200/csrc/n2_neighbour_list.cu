// n2_neighbour_list.cu -- C ABI of the device neighbour list / sparse descriptor:
// CUDA backend of n2_neighbour_driver.h (launches on the context's stream, cub::DeviceScan for the
// prefix sums) and the host-side getCellTranslations.
#include <cmath>

#include <cub/device/device_scan.cuh>

#include "block_desc.cuh"
#include "flat_stream.cuh"
#include "n2_neighbour_driver.h"

using namespace dftbp_b200;

namespace {

struct CudaBackend {
  dftbp_b200_ctx *ctx;
  nl::Buf<char> scanTemp;

  int alloc(void **p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess)
      return fail(ctx, DFTBP_B200_ENOMEM, "neighbour list: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return 0;
  }
  void release(void *p) { cudaFree(p); }
  int upload(void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    if (isDevicePtr(src)) {
      DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      // pageable host memory: the copy has left the caller's buffer when the call returns
      DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    return 0;
  }
  int download(void *dst, const void *src, size_t bytes) {
    DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  int fill(void *p, int byteValue, size_t bytes) {
    if (bytes == 0) return 0;
    DFTBP_CUDA(ctx, cudaMemsetAsync(p, byteValue, bytes, ctx->stream));
    return 0;
  }
  int scan(const int *in, int *out, long long n) {
    size_t bytes = 0;
    DFTBP_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ctx->stream));
    DFTBP_TRY(scanTemp.ensure(*this, bytes));
    DFTBP_CUDA(ctx, cub::DeviceScan::ExclusiveSum(scanTemp.p, bytes, in, out, n, ctx->stream));
    ctx->launches++;
    return 0;
  }
  template <class... P, class... A>
  int launch(void (*kernel)(P...), long long nItems, A... args) {
    if (nItems <= 0) return 0;
    // per-atom kernels carry long serial loops per thread: small CTAs spread them over the SMs
    const int block = nItems >= 256LL * ctx->smCount ? 128 : 32;
    long long grid = (nItems + block - 1) / block;
    const long long cap = (long long)ctx->smCount * 32;
    if (grid > cap) grid = cap;
    kernel<<<(unsigned)grid, block, 0, ctx->stream>>>(args...);
    ctx->launches++;
    DFTBP_CUDA(ctx, cudaGetLastError());
    return 0;
  }
};

struct NeighbourState {
  nl::NeighbourList list;
  nl::Buf<char> scanTemp;
};

NeighbourState *stateOf(dftbp_b200_ctx *ctx) {
  if (!ctx->neighbourState) ctx->neighbourState = new NeighbourState();
  return static_cast<NeighbourState *>(ctx->neighbourState);
}

template <class T>
int copyOut(dftbp_b200_ctx *ctx, T *dst, const T *src, size_t count) {
  if (!dst || count == 0) return 0;
  cudaMemcpyKind kind = isDevicePtr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  DFTBP_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(T), kind, ctx->stream));
  return 0;
}

}  // namespace

namespace dftbp_b200 {
NeighbourListView neighbourListView(dftbp_b200_ctx *ctx) {
  NeighbourListView v;
  if (!ctx->neighbourState) return v;
  const nl::NeighbourList &L = static_cast<NeighbourState *>(ctx->neighbourState)->list;
  if (!L.valid || L.symmetric) return v;
  v.valid = true;
  v.nAtom = L.nAtom; v.nAllAtom = L.nAllAtom; v.mNeigh = L.mNeigh;
  v.coords = L.coords.p; v.neighDist2 = L.neighDist2.p; v.species = L.species.p; v.img2CentCell = L.img2CentCell.p;
  v.nNeighbour = L.nNeighbour.p; v.iNeighbour = L.iNeighbour.p;
  return v;
}

void freeNeighbourList(dftbp_b200_ctx *ctx) {
  if (!ctx->neighbourState) return;
  NeighbourState *st = static_cast<NeighbourState *>(ctx->neighbourState);
  CudaBackend be{ctx};
  st->list.free(be);
  st->scanTemp.free(be);
  delete st;
  ctx->neighbourState = nullptr;
}
}  // namespace dftbp_b200

namespace {

static_assert(sizeof(nl::GeoDesc) == sizeof(BlockDesc) && offsetof(nl::GeoDesc, ind) == offsetof(BlockDesc, ind) &&
                  offsetof(nl::GeoDesc, n12) == offsetof(BlockDesc, n12) &&
                  offsetof(nl::GeoDesc, iAt2f) == offsetof(BlockDesc, iAt2f) &&
                  offsetof(nl::GeoDesc, iNeigh) == offsetof(BlockDesc, iNeigh),
              "nl::GeoDesc must mirror BlockDesc");

template <class T>
int poolCopy(dftbp_b200_ctx *ctx, T **dst, const T *src, size_t count) {
  void *blk = nullptr;
  DFTBP_TRY(poolAlloc(ctx, &blk, (count ? count : 1) * sizeof(T)));
  *dst = static_cast<T *>(blk);
  if (count && src)
    DFTBP_CUDA(ctx, cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

// The device list becomes the context's geometry without a round trip of the index arrays: the arrays
// are copied device-to-device into pool blocks, the work list, the block descriptors, the element ->
// descriptor index of the span pipeline and the hasDuplicates flag are built by kernels
// (n2_neighbour_kernels.cuh: geo_*); the host reads back the first block of every atom (nAtom + 1
// ints: sizes the arrays, serves the atom-range calls of packHS) and one flag word.  The list is valid
// by construction (species, images and offsets come from the kernels of update_neighbour_list), and its
// blocks tile [0, sparseSize) exactly.  One rank only (see the caller).
int adoptOnDevice(dftbp_b200_ctx *ctx, NeighbourState *st) {
  const nl::NeighbourList &L = st->list;
  CudaBackend be{ctx};
  be.scanTemp = st->scanTemp;
  struct KeepTemp {  // the scan scratch may have grown
    NeighbourState *st; CudaBackend *be;
    ~KeepTemp() { st->scanTemp = be->scanTemp; }
  } keep{st, &be};
  const int nAtom = L.nAtom, ld = L.mNeigh + 1;
  releaseGeometry(ctx);
  Geometry g;
  g.nAtom = nAtom; g.nAllAtom = L.nAllAtom; g.mNeigh = L.mNeigh; g.nCellVec = L.nCellVec;
  g.sparseSize = L.sparseSize; g.nOrb = L.nOrb;
  g.coversAll = true;
  DFTBP_TRY(poolCopy(ctx, &g.coords, (const double *)L.coords.p, (size_t)3 * L.nAllAtom));
  DFTBP_TRY(poolCopy(ctx, &g.species, (const int *)L.species.p, L.nAllAtom));
  DFTBP_TRY(poolCopy(ctx, &g.img2CentCell, (const int *)L.img2CentCell.p, L.nAllAtom));
  DFTBP_TRY(poolCopy(ctx, &g.iCellVec, (const int *)L.iCellVec.p, L.nAllAtom));
  DFTBP_TRY(poolCopy(ctx, &g.cellVec, (const double *)L.cellVec.p, (size_t)3 * L.nCellVec));
  DFTBP_TRY(poolCopy(ctx, &g.nNeighbourSK, (const int *)L.nNeighbourSK.p, nAtom));
  DFTBP_TRY(poolCopy(ctx, &g.iNeighbour, (const int *)L.iNeighbour.p, (size_t)ld * nAtom));
  DFTBP_TRY(poolCopy(ctx, &g.iPair, (const int *)L.iPair.p, (size_t)ld * nAtom));
  DFTBP_TRY(poolCopy(ctx, &g.iAtomStart, (const int *)L.iAtomStart.p, (size_t)nAtom + 1));
  ctx->geo = g;  // from here on the blocks belong to the context (released with the geometry on any failure)

  // first block of every atom: scratch in the list's own buffers (atomSize / atomOffset are free after the update)
  std::vector<int> hStart((size_t)nAtom + 1);
  nl::NeighbourList &Lw = st->list;
  DFTBP_TRY(nl::countBlocks(be, L, Lw.atomSize.p, Lw.atomOffset.p, hStart.data()));
  const int nBlocks = hStart[nAtom];
  ctx->blockStartOfAtom.assign(hStart.begin(), hStart.end());
  ctx->loadInd.clear();
  ctx->rangeLo = ctx->rangeHi = -1;
  const size_t nGroups = ((size_t)L.sparseSize + flat::kGroup - 1) / flat::kGroup;
  Geometry &G = ctx->geo;
  DFTBP_TRY(poolCopy(ctx, &G.workAll, (const int *)nullptr, (size_t)nBlocks));
  DFTBP_TRY(poolCopy(ctx, &G.descAll, (const BlockDesc *)nullptr, (size_t)nBlocks));
  DFTBP_TRY(poolCopy(ctx, &G.groupFirstAll, (const int *)nullptr, nGroups));
  int hFlag = 0;
  DFTBP_TRY(nl::describeBlocks(be, L, (const int *)Lw.atomOffset.p, nBlocks, flat::kGroup, G.workAll,
                               reinterpret_cast<nl::GeoDesc *>(G.descAll), G.groupFirstAll, Lw.sizes.p, &hFlag));
  G.hasDuplicates = hFlag != 0;
  G.workMine = G.workLoad = G.workAll;
  G.descLoad = G.descAll;
  G.groupFirstLoad = G.groupFirstAll;
  G.nWorkAll = G.nWorkMine = G.nWorkLoad = nBlocks;
  G.nPairsMine = (long long)nBlocks - nAtom;
  G.firstAtomMine = 0;
  G.lastAtomMine = nAtom - 1;
  ctx->sortMine.valid = ctx->sortLoad.valid = false;
  return 0;
}

}  // namespace

extern "C" {

int dftbp_b200_cell_translations(const double *latVec, const double *recVec2p, double cutoff, int maxCellVec,
                                 int *nCellVec, double *cellVec, double *rCellVec) {
  if (!latVec || !recVec2p || !nCellVec || !(cutoff >= 0.0)) return DFTBP_B200_EINVAL;
  // getRanges_ (type/latpointiter.F90:250-274) with posExtension = negExtension = 1
  int rng[3];
  for (int i = 0; i < 3; ++i) {
    const double *v = recVec2p + 3 * i;
    rng[i] = (int)std::floor(cutoff * std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) + 1;
  }
  const long long total = (2LL * rng[0] + 1) * (2LL * rng[1] + 1) * (2LL * rng[2] + 1);
  *nCellVec = (int)total;
  if (!cellVec || !rCellVec) return 0;  // size query
  if (maxCellVec < total) return DFTBP_B200_EINVAL;
  // origin first (getLatticePoints, periodic.F90:318-323), then every other point of the
  // parallelepiped with the third index running fastest (getNextPoint, latpointiter.F90:214-228)
  long long n = 0;
  auto put = [&](int a, int b, int c) {
    const double p[3] = {(double)a, (double)b, (double)c};
    for (int k = 0; k < 3; ++k) {
      cellVec[3 * n + k] = p[k];
      // rCellVec(:,ii) = matmul(latVec, cellVec(:,ii)), periodic.F90:252
      rCellVec[3 * n + k] = latVec[k] * p[0] + latVec[k + 3] * p[1] + latVec[k + 6] * p[2];
    }
    ++n;
  };
  put(0, 0, 0);
  for (int a = -rng[0]; a <= rng[0]; ++a)
    for (int b = -rng[1]; b <= rng[1]; ++b)
      for (int c = -rng[2]; c <= rng[2]; ++c)
        if (a != 0 || b != 0 || c != 0) put(a, b, c);
  return 0;
}

int dftbp_b200_fold_coords_to_cell(dftbp_b200_ctx *ctx, int nAtom, double *coord, const double *latVec) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nAtom > 0 && coord && latVec, "fold_coords_to_cell: bad arguments");
  double lat[9];
  if (isDevicePtr(latVec)) {
    DFTBP_CUDA(ctx, cudaMemcpy(lat, latVec, sizeof(lat), cudaMemcpyDeviceToHost));
  } else {
    std::memcpy(lat, latVec, sizeof(lat));
  }
  nl::FoldParams fp;
  DFTBP_CHECK(ctx, nl::foldSetup(lat, fp), "fold_coords_to_cell: singular lattice");
  double *d = nullptr;
  DFTBP_TRY(devOut(ctx, kSlotA, coord, (size_t)3 * nAtom, true, &d));
  CudaBackend be{ctx};
  DFTBP_TRY(be.launch(nl::nl_fold, nAtom, nAtom, fp, d));
  return finish(ctx);
}

int dftbp_b200_update_neighbour_list(dftbp_b200_ctx *ctx, int nAtom, const double *coord0, const int *species0,
                                     int nCellVec, const double *cellVec, const double *rCellVec, double cutoff,
                                     double skCutoff, int symmetric, int *nAllAtom, int *mNeigh, int *sparseSize) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->nSpecies > 0, "update_neighbour_list: call set_orbitals first");
  DFTBP_CHECK(ctx, nAtom > 0 && nCellVec >= 1 && coord0 && species0 && cellVec && rCellVec,
              "update_neighbour_list: bad arguments");
  DFTBP_CHECK(ctx, cutoff >= 0.0 && skCutoff >= 0.0, "update_neighbour_list: negative cutoff");
  DFTBP_CHECK(ctx, (long long)nCellVec * nAtom < (1LL << 31) - 1, "update_neighbour_list: too many candidate images");
  {
    std::vector<int> sp(nAtom);
    if (isDevicePtr(species0)) {
      DFTBP_CUDA(ctx, cudaMemcpy(sp.data(), species0, sizeof(int) * nAtom, cudaMemcpyDeviceToHost));
    } else {
      std::memcpy(sp.data(), species0, sizeof(int) * nAtom);
    }
    for (int i = 0; i < nAtom; ++i)
      DFTBP_CHECK(ctx, sp[i] >= 1 && sp[i] <= ctx->nSpecies, "update_neighbour_list: species out of range");
  }
  NeighbourState *st = stateOf(ctx);
  ctx->listGeneration++;
  CudaBackend be{ctx};
  be.scanTemp = st->scanTemp;
  ScopedTimer timer(ctx, kTimerNeighbourList);
  const int rc = nl::updateNeighbourList(be, st->list, nAtom, coord0, species0, ctx->dNOrbSpecies, nCellVec, cellVec,
                                         rCellVec, cutoff, skCutoff, symmetric != 0);
  st->scanTemp = be.scanTemp;
  // host inputs were staged with asynchronous copies: they have been consumed once the stream is idle
  // (updateNeighbourList ends with a blocking read-back unless it failed early)
  if (rc != 0) {
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  if (st->list.errAtom1 != 0)
    return fail(ctx, DFTBP_B200_EGEOMETRY, "Atoms %d and %d too close together", st->list.errAtom1,
                st->list.errAtom2);
  if (nAllAtom) *nAllAtom = st->list.nAllAtom;
  if (mNeigh) *mNeigh = st->list.mNeigh;
  if (sparseSize) *sparseSize = st->list.sparseSize;
  return 0;
}

int dftbp_b200_get_neighbour_list(dftbp_b200_ctx *ctx, double *coords, int *species, int *img2CentCell,
                                  int *iCellVec, int *nNeighbour, int *iNeighbour, double *neighDist2,
                                  int *nNeighbourSK, int *iPair, int *iAtomStart) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->neighbourState && stateOf(ctx)->list.valid,
              "get_neighbour_list: no neighbour list (call update_neighbour_list first)");
  const nl::NeighbourList &L = stateOf(ctx)->list;
  const size_t rows = (size_t)(L.mNeigh + 1) * L.nAtom;
  DFTBP_TRY(copyOut(ctx, coords, (const double *)L.coords.p, (size_t)3 * L.nAllAtom));
  DFTBP_TRY(copyOut(ctx, species, (const int *)L.species.p, L.nAllAtom));
  DFTBP_TRY(copyOut(ctx, img2CentCell, (const int *)L.img2CentCell.p, L.nAllAtom));
  DFTBP_TRY(copyOut(ctx, iCellVec, (const int *)L.iCellVec.p, L.nAllAtom));
  DFTBP_TRY(copyOut(ctx, nNeighbour, (const int *)L.nNeighbour.p, L.nAtom));
  DFTBP_TRY(copyOut(ctx, iNeighbour, (const int *)L.iNeighbour.p, rows));
  DFTBP_TRY(copyOut(ctx, neighDist2, (const double *)L.neighDist2.p, rows));
  DFTBP_TRY(copyOut(ctx, nNeighbourSK, (const int *)L.nNeighbourSK.p, L.nAtom));
  DFTBP_TRY(copyOut(ctx, iPair, (const int *)L.iPair.p, rows));
  DFTBP_TRY(copyOut(ctx, iAtomStart, (const int *)L.iAtomStart.p, (size_t)L.nAtom + 1));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dftbp_b200_nr_of_neighbours(dftbp_b200_ctx *ctx, double cutoff, int *nNeighbour) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->neighbourState && stateOf(ctx)->list.valid,
              "nr_of_neighbours: no neighbour list (call update_neighbour_list first)");
  DFTBP_CHECK(ctx, cutoff >= 0.0 && nNeighbour, "nr_of_neighbours: bad arguments");
  const nl::NeighbourList &L = stateOf(ctx)->list;
  int *dOut;
  DFTBP_TRY(devOut(ctx, kSlotA, nNeighbour, (size_t)L.nAtom, false, &dOut));
  CudaBackend be{ctx};
  DFTBP_TRY(be.launch(nl::nl_nr_of_neighbours, L.nAtom, L.nAtom, L.mNeigh + 1, cutoff * cutoff,
                      (const int *)L.nNeighbour.p, (const double *)L.neighDist2.p, dOut));
  return finish(ctx);
}

int dftbp_b200_adopt_neighbour_list(dftbp_b200_ctx *ctx) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->neighbourState && stateOf(ctx)->list.valid,
              "adopt_neighbour_list: no neighbour list (call update_neighbour_list first)");
  const nl::NeighbourList &L = stateOf(ctx)->list;
  DFTBP_CHECK(ctx, !L.symmetric, "adopt_neighbour_list: a symmetric list is not a geometry for the pair kernels");
  if (ctx->nRanks > 1 || tuneInt("DFTBP_B200_ADOPT_HOST", 0) != 0) {
    // several ranks: the atom chunks / workload-balanced atom sets of set_geometry are host-side logic
    DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // device pointers: set_geometry copies them device-to-device and reads what it validates
    return dftbp_b200_set_geometry(ctx, L.nAtom, L.nAllAtom, L.mNeigh, L.coords.p, L.species.p, L.img2CentCell.p,
                                   L.iCellVec.p, L.nCellVec, L.cellVec.p, L.nNeighbourSK.p, L.iNeighbour.p,
                                   L.iPair.p, L.iAtomStart.p, L.sparseSize);
  }
  return adoptOnDevice(ctx, stateOf(ctx));
}

}  // extern "C"

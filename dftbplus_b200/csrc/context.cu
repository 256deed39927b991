// context.cu -- lifecycle, TOrbitals / TSlakoCont upload, geometry upload.
#include <algorithm>
#include <cmath>

#include "block_desc.cuh"
#include "flat_stream.cuh"
#include "sk_table_prep.h"

using namespace dftbp_b200;

namespace {

void freeGeometry(dftbp_b200_ctx *ctx, Geometry &g) {
  auto rel = [&](void *p) { poolRelease(ctx, p); };
  rel(g.coords); rel(g.species); rel(g.img2CentCell); rel(g.iCellVec);
  rel(g.cellVec); rel(g.nNeighbourSK); rel(g.iNeighbour); rel(g.iPair);
  rel(g.iAtomStart);
  if (g.workMine != g.workAll) rel(g.workMine);
  if (g.workLoad != g.workAll) rel(g.workLoad);
  rel(g.workAll);
  if (g.descLoad != g.descAll) rel(g.descLoad);
  rel(g.descAll);
  if (g.groupFirstLoad != g.groupFirstAll) rel(g.groupFirstLoad);
  rel(g.groupFirstAll);
  g = Geometry();
}

template <class T>
std::vector<T> toHost(const T *p, size_t n) {
  std::vector<T> v(n);
  if (n == 0) return v;
  if (isDevicePtr(p)) cudaMemcpy(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost);
  else std::memcpy(v.data(), p, n * sizeof(T));
  return v;
}

}  // namespace

// getChunkRanges, common/schedule.F90:230-263 (0-based rank, 1-based inclusive range)
static void chunkRange(int groupSize, int myRank, int first, int last, int *lf, int *ll) {
  int len = last - first + 1, nLocal = len / groupSize, rem = len % groupSize;
  if (myRank < rem) {
    nLocal += 1;
    *lf = first + myRank * nLocal;
  } else {
    *lf = first + rem * (nLocal + 1) + (myRank - rem) * nLocal;
  }
  *ll = std::min(*lf + nLocal - 1, last);
}

extern "C" {

const char *dftbp_b200_version(void) { return "dftbp_b200 0.1.0 (sm_100a)"; }

int dftbp_b200_create(dftbp_b200_ctx **out, int device, int rank, int nRanks) {
  if (!out) return DFTBP_B200_EINVAL;
  *out = nullptr;
  int nDev = 0;
  cudaError_t e = cudaGetDeviceCount(&nDev);
  dftbp_b200_ctx *ctx = new dftbp_b200_ctx();
  *out = ctx;  // returned even on failure so that last_error is readable
  if (e != cudaSuccess || nDev == 0)
    return fail(ctx, DFTBP_B200_ENODEVICE,
                "no CUDA device available (%s); libdftbp_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= nDev)
    return fail(ctx, DFTBP_B200_EINVAL, "device %d out of range (0..%d)", device, nDev - 1);
  if (nRanks < 1 || rank < 0 || rank >= nRanks)
    return fail(ctx, DFTBP_B200_EINVAL, "bad rank %d of %d", rank, nRanks);
  ctx->device = device; ctx->rank = rank; ctx->nRanks = nRanks;
  DFTBP_CUDA(ctx, cudaSetDevice(device));
  DFTBP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  DFTBP_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  ctx->smCount = prop.multiProcessorCount;
  ctx->smemOptin = prop.sharedMemPerBlockOptin;
  ctx->smemPerSm = prop.sharedMemPerMultiprocessor;
  DFTBP_CUDA(ctx, cudaMalloc((void **)&ctx->dInfo, 16 * sizeof(int)));
  // [0] info / nLevels, [1] shift flag, [2..3] K3 tile counter, [4..5] K1 / [6..7] K5 / [8..9] K4 item schedulers
  DFTBP_CUDA(ctx, cudaMemset(ctx->dInfo, 0, 16 * sizeof(int)));
  DFTBP_CUDA(ctx, cudaMallocHost((void **)&ctx->hShiftHint, sizeof(int)));
  *ctx->hShiftHint = -1;
  return 0;
}

int dftbp_b200_destroy(dftbp_b200_ctx *ctx) {
  if (!ctx) return 0;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  freeGeometry(ctx, ctx->geo);
  poolDestroy(ctx);
  freeSortedPairs(ctx);
  freeNeighbourList(ctx);
  freeShortGamma(ctx);
  freeMixer(ctx);
  for (void *p : ctx->tableAllocs) cudaFree(p);
  for (Staging &s : ctx->staging) cudaFree(s.ptr);
  cudaFree(ctx->dNOrbSpecies); cudaFree(ctx->dOrbShell); cudaFree(ctx->dProgs);
  cudaFree(ctx->pairExchange);
  cudaFree(ctx->shiftDiag);
  if (ctx->hShiftHint) cudaFreeHost(ctx->hShiftHint);
  cudaFree(ctx->solverWork); cudaFree(ctx->dInfo);
  solverDestroy(ctx);
  for (TimedSpan &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

const char *dftbp_b200_last_error(const dftbp_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dftbp_b200_set_stream(dftbp_b200_ctx *ctx, void *s) {
  if (!ctx) return DFTBP_B200_EINVAL;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (s == nullptr) {
    if (!ctx->ownStream) {
      DFTBP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->ownStream = true;
    }
    return 0;
  }
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (s == (void *)1) ? cudaStreamLegacy : (cudaStream_t)s;
  ctx->ownStream = false;
  return 0;
}

int dftbp_b200_timer_enable(dftbp_b200_ctx *ctx, int on) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (TimedSpan &sp : ctx->spans) {
    ctx->eventPool.push_back(sp.a);
    ctx->eventPool.push_back(sp.b);
  }
  ctx->spans.clear();
  for (int i = 0; i < kTimerCount; ++i) {
    ctx->timerMs[i] = 0.0;
    ctx->timerCalls[i] = 0;
  }
  ctx->timing = on != 0;
  return 0;
}

int dftbp_b200_timer_read(dftbp_b200_ctx *ctx, int slot, double *ms, long long *calls) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, slot >= 0 && slot < kTimerCount && ms && calls, "timer_read: bad arguments");
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (TimedSpan &sp : ctx->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) {
      ctx->timerMs[sp.slot] += t;
      ctx->timerCalls[sp.slot] += 1;
    }
    ctx->eventPool.push_back(sp.a);
    ctx->eventPool.push_back(sp.b);
  }
  ctx->spans.clear();
  *ms = ctx->timerMs[slot];
  *calls = ctx->timerCalls[slot];
  return 0;
}

void *dftbp_b200_get_stream(dftbp_b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int dftbp_b200_synchronize(dftbp_b200_ctx *ctx) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

long long dftbp_b200_kernel_launches(const dftbp_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }
long long dftbp_b200_num_pairs(const dftbp_b200_ctx *ctx) { return ctx ? ctx->geo.nPairsMine : 0; }

int dftbp_b200_set_orbitals(dftbp_b200_ctx *ctx, int nSpecies, int mShell, const int *nShell,
                            const int *angShell) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nSpecies > 0 && mShell > 0 && nShell && angShell, "set_orbitals: bad arguments");
  ctx->nSpecies = nSpecies;
  ctx->mShell = mShell;
  ctx->nShell.assign(nShell, nShell + nSpecies);
  ctx->angShell.assign(angShell, angShell + (size_t)mShell * nSpecies);
  ctx->nOrbSpecies.assign(nSpecies, 0);
  ctx->mOrb = 0;
  ctx->lMaxAll = 0;
  for (int sp = 0; sp < nSpecies; ++sp) {
    DFTBP_CHECK(ctx, nShell[sp] >= 1 && nShell[sp] <= mShell, "set_orbitals: nShell out of range");
    for (int sh = 0; sh < nShell[sp]; ++sh) {
      int l = angShell[sh + mShell * sp];
      DFTBP_CHECK(ctx, l >= 0 && l <= 3, "set_orbitals: angular momentum must be 0..3");  // mAngRot_ = 3, sk.F90:28
      ctx->nOrbSpecies[sp] += 2 * l + 1;
      ctx->lMaxAll = std::max(ctx->lMaxAll, l);
    }
    ctx->mOrb = std::max(ctx->mOrb, ctx->nOrbSpecies[sp]);
  }
  DFTBP_CHECK(ctx, ctx->mOrb <= kMaxOrb, "set_orbitals: more than 16 orbitals per atom");
  // shell index of each orbital (iShellOrb, type/orbitals.F90:33-35), 0-based
  std::vector<int> orbShell((size_t)ctx->mOrb * nSpecies, 0);
  for (int sp = 0; sp < nSpecies; ++sp) {
    int o = 0;
    for (int sh = 0; sh < nShell[sp]; ++sh)
      for (int k = 0; k < 2 * angShell[sh + mShell * sp] + 1; ++k) orbShell[o++ + ctx->mOrb * sp] = sh;
  }
  DFTBP_TRY(uploadNew(ctx, &ctx->dNOrbSpecies, ctx->nOrbSpecies.data(), nSpecies));
  DFTBP_TRY(uploadNew(ctx, &ctx->dOrbShell, orbShell.data(), orbShell.size()));
  // Per-species-pair programs: shell structure and integral count (the rotation itself is the
  // generated code of sk_gaunt_code.cuh; the SK tables are attached by set_sk_table).
  ctx->progs.assign((size_t)nSpecies * nSpecies, SpPairProg());
  for (int sp2 = 0; sp2 < nSpecies; ++sp2)
    for (int sp1 = 0; sp1 < nSpecies; ++sp1) {
      SpPairProg &pg = ctx->progs[sp1 + nSpecies * sp2];
      std::memset(&pg, 0, sizeof(pg));
      pg.nOrb1 = ctx->nOrbSpecies[sp1];
      pg.nOrb2 = ctx->nOrbSpecies[sp2];
      int ind = 0, lmax1 = 0, lmax2 = 0;
      // integrals in getFullTable order (dftbplus/parser.F90:4096-4127): atom-1 shells outer, min(l1,l2)+1 each
      for (int sh1 = 0; sh1 < nShell[sp1]; ++sh1) {
        const int l1 = angShell[sh1 + mShell * sp1];
        lmax1 = std::max(lmax1, l1);
        for (int sh2 = 0; sh2 < nShell[sp2]; ++sh2) {
          const int l2 = angShell[sh2 + mShell * sp2];
          lmax2 = std::max(lmax2, l2);
          ind += std::min(l1, l2) + 1;
        }
      }
      pg.nInteg = ind;
      DFTBP_CHECK(ctx, ind <= kMaxInteg, "set_orbitals: too many integrals per species pair");
      pg.nSh1 = nShell[sp1];
      pg.nSh2 = nShell[sp2];
      auto canon = [&](int sp) {
        for (int sh = 0; sh < nShell[sp]; ++sh)
          if (angShell[sh + mShell * sp] != sh) return -1;
        return nShell[sp] - 1;
      };
      pg.canon1 = canon(sp1);
      pg.canon2 = canon(sp2);
      // tuning switch: every species pair through the run-time shell loops (22 KB of code for spd
      // instead of 16-38 KB per canonical case) -- A/B of the instruction-fetch bound
      if (tuneInt("DFTBP_B200_FORCE_GENERIC", 0)) pg.canon1 = pg.canon2 = -1;
      for (int sh = 0; sh < nShell[sp1]; ++sh)
        pg.shells1 |= (unsigned long long)angShell[sh + mShell * sp1] << (4 * sh);
      for (int sh = 0; sh < nShell[sp2]; ++sh)
        pg.shells2 |= (unsigned long long)angShell[sh + mShell * sp2] << (4 * sh);
      pg.nY = (lmax1 + lmax2 + 1) * (lmax1 + lmax2 + 1);
      DFTBP_CHECK(ctx, fillShellProgram(pg, ctx->lMaxAll), "set_orbitals: too many shell pairs per species pair");
    }
  ctx->allSimple = true;
  for (const SpPairProg &pg : ctx->progs) ctx->allSimple = ctx->allSimple && pg.simple;
  ctx->progsDirty = true;
  return 0;
}

int dftbp_b200_set_sk_table(dftbp_b200_ctx *ctx, int which, int sp1, int sp2, int nGrid,
                            int nInteg, double dist, const double *skTab, int method) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->nSpecies > 0, "set_sk_table: call set_orbitals first");
  DFTBP_CHECK(ctx, which == 0 || which == 1, "set_sk_table: which must be 0 (H) or 1 (S)");
  DFTBP_CHECK(ctx, sp1 >= 1 && sp1 <= ctx->nSpecies && sp2 >= 1 && sp2 <= ctx->nSpecies,
              "set_sk_table: species out of range");
  DFTBP_CHECK(ctx, skTab && dist > 0.0, "set_sk_table: bad table");
  SpPairProg &pg = ctx->progs[(sp1 - 1) + ctx->nSpecies * (sp2 - 1)];
  if (nInteg != pg.nInteg)
    return fail(ctx, DFTBP_B200_EINVAL, "set_sk_table: species pair (%d,%d) needs %d integrals, got %d",
                sp1, sp2, pg.nInteg, nInteg);
  DFTBP_CHECK(ctx, method == DFTBP_B200_SK_EQGRID_NEW || method == DFTBP_B200_SK_EQGRID_OLD,
              "set_sk_table: unknown interpolation method");
  std::vector<double> host = toHost(skTab, (size_t)nGrid * nInteg), tail;
  SkTableDev &t = which == 0 ? pg.H : pg.S;
  // SlakoEqGrid_interNew_ / _interOld_ abort for tables shorter than nInter + 1 points
  // (slakoeqgrid.F90:207-210, :271-273)
  if (!prepareSkTable(t, nGrid, nInteg, dist, method, host.data(), tail))
    return fail(ctx, DFTBP_B200_EINVAL, "SlakoEqGrid: Not enough points in the SK-table for interpolation!");
  host.resize(host.size() + 8, 0.0);  // the 8-row window of a short skEqGridOld table may pass the last column
  double *dTab = nullptr, *dTail = nullptr;
  DFTBP_TRY(uploadNew(ctx, &dTab, host.data(), host.size()));
  DFTBP_TRY(uploadNew(ctx, &dTail, tail.data(), tail.size()));
  ctx->tableAllocs.push_back(dTab);
  ctx->tableAllocs.push_back(dTail);
  t.tab = dTab; t.tail = dTail;
  ctx->cutoff = std::max(ctx->cutoff, skTableCutoff(nGrid, dist, method));
  ctx->progsDirty = true;
  return 0;
}

int dftbp_b200_set_pair_partition(dftbp_b200_ctx *ctx, int part, int nParts) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, nParts >= 1 && part >= 0 && part < nParts, "set_pair_partition: need 0 <= part < nParts");
  if (ctx->pairPart != part || ctx->pairParts != nParts) ctx->sortLoad.valid = false;  // the share is fixed when the schedule is built
  ctx->pairPart = part;
  ctx->pairParts = nParts;
  return 0;
}

int dftbp_b200_get_sk_cutoff(dftbp_b200_ctx *ctx, double *cutoff) {
  if (!ctx || !cutoff) return DFTBP_B200_EINVAL;
  *cutoff = ctx->cutoff;
  return 0;
}

int dftbp_b200_set_geometry(dftbp_b200_ctx *ctx, int nAtom, int nAllAtom, int mNeigh,
                            const double *coords, const int *species, const int *img2CentCell,
                            const int *iCellVec, int nCellVec, const double *cellVec,
                            const int *nNeighbourSK, const int *iNeighbour, const int *iPair,
                            const int *iAtomStart, int sparseSize) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->nSpecies > 0, "set_geometry: call set_orbitals first");
  DFTBP_CHECK(ctx, nAtom > 0 && nAllAtom >= nAtom && mNeigh >= 0 && nCellVec >= 1 && sparseSize > 0,
              "set_geometry: bad sizes");
  DFTBP_CHECK(ctx, coords && species && img2CentCell && iCellVec && cellVec && nNeighbourSK &&
                       iNeighbour && iPair && iAtomStart, "set_geometry: null array");
  const int ld = mNeigh + 1;
  std::vector<int> hN = toHost(nNeighbourSK, nAtom), hNb = toHost(iNeighbour, (size_t)ld * nAtom);
  std::vector<int> hImg = toHost(img2CentCell, nAllAtom), hSp = toHost(species, nAllAtom);
  std::vector<int> hStart = toHost(iAtomStart, nAtom + 1), hPair = toHost(iPair, (size_t)ld * nAtom);
  std::vector<int> hVec = toHost(iCellVec, nAllAtom);
  for (int i = 0; i < nAtom; ++i) {
    DFTBP_CHECK(ctx, hN[i] >= 0 && hN[i] <= mNeigh, "set_geometry: nNeighbourSK exceeds mNeigh");
    DFTBP_CHECK(ctx, hSp[i] >= 1 && hSp[i] <= ctx->nSpecies, "set_geometry: species out of range");
    DFTBP_CHECK(ctx, hStart[i + 1] - hStart[i] == ctx->nOrbSpecies[hSp[i] - 1],
                "set_geometry: iAtomStart inconsistent with the orbitals of the species");
    for (int k = 0; k <= hN[i]; ++k) {
      int a2 = hNb[k + (size_t)ld * i];
      DFTBP_CHECK(ctx, a2 >= 1 && a2 <= nAllAtom, "set_geometry: iNeighbour entry out of range");
      DFTBP_CHECK(ctx, hImg[a2 - 1] >= 1 && hImg[a2 - 1] <= nAtom, "set_geometry: img2CentCell entry out of range");
      DFTBP_CHECK(ctx, hVec[a2 - 1] >= 1 && hVec[a2 - 1] <= nCellVec, "set_geometry: iCellVec entry out of range");
    }
  }
  // every block must lie inside the sparse arrays; the build kernels write each element exactly
  // once and skip the zero fill only if the blocks cover the arrays completely (they do for
  // getSparseDescriptor, sparse2dense.F90:3617-3662; an over-allocated caller gets a memset)
  long long covered = 0;
  for (int i = 0; i < nAtom; ++i)
    for (int k = 0; k <= hN[i]; ++k) {
      const long long f = k + (long long)ld * i;
      const int n1 = hStart[i + 1] - hStart[i];
      const int j = hImg[hNb[f] - 1] - 1, n2 = hStart[j + 1] - hStart[j];
      DFTBP_CHECK(ctx, hPair[f] >= 0 && (long long)hPair[f] + (long long)n1 * n2 <= sparseSize,
                  "set_geometry: iPair block exceeds sparseSize");
      covered += (long long)n1 * n2;
    }
  Geometry g;
  freeGeometry(ctx, ctx->geo);
  g.coversAll = covered == sparseSize;
  g.nAtom = nAtom; g.nAllAtom = nAllAtom; g.mNeigh = mNeigh; g.nCellVec = nCellVec;
  g.sparseSize = sparseSize; g.nOrb = hStart[nAtom] - 1;
  DFTBP_TRY(uploadNew(ctx, &g.coords, coords, (size_t)3 * nAllAtom));
  DFTBP_TRY(uploadNew(ctx, &g.species, species, nAllAtom));
  DFTBP_TRY(uploadNew(ctx, &g.img2CentCell, img2CentCell, nAllAtom));
  DFTBP_TRY(uploadNew(ctx, &g.iCellVec, iCellVec, nAllAtom));
  DFTBP_TRY(uploadNew(ctx, &g.cellVec, cellVec, (size_t)3 * nCellVec));
  DFTBP_TRY(uploadNew(ctx, &g.nNeighbourSK, nNeighbourSK, nAtom));
  DFTBP_TRY(uploadNew(ctx, &g.iNeighbour, iNeighbour, (size_t)ld * nAtom));
  DFTBP_TRY(uploadNew(ctx, &g.iPair, iPair, (size_t)ld * nAtom));
  DFTBP_TRY(uploadNew(ctx, &g.iAtomStart, iAtomStart, nAtom + 1));

  // Work lists: one entry per sparse block.
  std::vector<int> all, mine, load;
  std::vector<int> stamp(nAtom + 1, -1);
  for (int i = 0; i < nAtom; ++i) {
    for (int k = 0; k <= hN[i]; ++k) all.push_back(k + ld * i);
    stamp[i + 1] = i;  // the on-site block
    for (int k = 1; k <= hN[i]; ++k) {
      int f = hImg[hNb[k + (size_t)ld * i] - 1];
      if (stamp[f] == i) g.hasDuplicates = true;
      stamp[f] = i;
    }
  }
  DFTBP_TRY(uploadNew(ctx, &g.workAll, all.data(), all.size()));
  g.nWorkAll = (long long)all.size();
  ctx->blockStartOfAtom.assign(nAtom + 1, 0);
  for (int i = 0; i < nAtom; ++i) ctx->blockStartOfAtom[i + 1] = ctx->blockStartOfAtom[i] + hN[i] + 1;
  // resolve iNeighbour -> img2CentCell -> iAtomStart / iCellVec / iPair once per geometry (block_desc.cuh)
  auto describe = [&](const std::vector<int> &work) {
    std::vector<BlockDesc> d(work.size());
    for (size_t w = 0; w < work.size(); ++w) {
      const int flat = work[w], i = flat / ld, k = flat - i * ld;
      const int a2 = hNb[flat] - 1, f = hImg[a2] - 1;
      BlockDesc &b = d[w];
      b.ind = hPair[flat];
      b.ii = hStart[i] - 1;
      b.jj = hStart[f] - 1;
      b.n12 = (hStart[i + 1] - hStart[i]) | ((hStart[f + 1] - hStart[f]) << 8);
      b.iAt1 = i;
      b.iAt2f = f;
      b.iVec = hVec[a2] - 1;
      b.iNeigh = k;
    }
    return d;
  };
  // element -> descriptor index of a descriptor list for the TMA-streamed kernels (flat_stream.cuh)
  auto groupIndex = [&](const std::vector<BlockDesc> &d, int **out) -> int {
    std::vector<int> first;
    if (!flat::buildGroupIndex(d.data(), (long long)d.size(), sparseSize, first)) return 0;
    return uploadNew(ctx, out, first.data(), first.size());
  };
  {
    std::vector<BlockDesc> d = describe(all);
    DFTBP_TRY(uploadNew(ctx, &g.descAll, d.data(), d.size()));
    DFTBP_TRY(groupIndex(d, &g.groupFirstAll));
  }
  if (ctx->nRanks == 1) {
    g.workMine = g.workLoad = g.workAll;
    g.descLoad = g.descAll;
    g.groupFirstLoad = g.groupFirstAll;
    g.nWorkMine = g.nWorkLoad = g.nWorkAll;
    g.nPairsMine = g.nWorkAll - nAtom;
    g.firstAtomMine = 0;
    g.lastAtomMine = nAtom - 1;
  } else {
    int first, last;
    chunkRange(ctx->nRanks, ctx->rank, 1, nAtom, &first, &last);  // distributeRangeInChunks
    for (int i = first - 1; i <= last - 1; ++i)
      for (int k = 0; k <= hN[i]; ++k) mine.push_back(k + ld * i);
    // getIndicesWithWorkload, common/schedule.F90:267-307
    std::vector<long long> rankLoad(ctx->nRanks, 0);
    for (int i = 0; i < nAtom; ++i) {
      int r = (int)(std::min_element(rankLoad.begin(), rankLoad.end()) - rankLoad.begin());
      rankLoad[r] += std::max(1, hN[i]);
      if (r == ctx->rank)
        for (int k = 0; k <= hN[i]; ++k) load.push_back(k + ld * i);
    }
    DFTBP_TRY(uploadNew(ctx, &g.workMine, mine.data(), mine.size()));
    DFTBP_TRY(uploadNew(ctx, &g.workLoad, load.data(), load.size()));
    {
      std::vector<BlockDesc> d = describe(load);
      DFTBP_TRY(uploadNew(ctx, &g.descLoad, d.data(), d.size()));
      DFTBP_TRY(groupIndex(d, &g.groupFirstLoad));
    }
    g.nWorkMine = (long long)mine.size();
    g.nWorkLoad = (long long)load.size();
    g.nPairsMine = g.nWorkMine - std::max(0, last - first + 1);
    g.firstAtomMine = first - 1;
    g.lastAtomMine = last - 1;
  }
  {
    const std::vector<int> &wl = ctx->nRanks == 1 ? all : load;
    ctx->loadInd.resize(wl.size() + 1);
    int end = 0;
    for (size_t w = 0; w < wl.size(); ++w) {
      const int flat = wl[w], i = flat / ld;
      const int f = hImg[hNb[flat] - 1] - 1;
      ctx->loadInd[w] = hPair[flat];
      end = hPair[flat] + (hStart[i + 1] - hStart[i]) * (hStart[f + 1] - hStart[f]);
    }
    ctx->loadInd[wl.size()] = end;
  }
  ctx->geo = g;
  ctx->sortMine.valid = ctx->sortLoad.valid = false;
  ctx->rangeLo = ctx->rangeHi = -1;
  return 0;
}

int dftbp_b200_update_coords(dftbp_b200_ctx *ctx, const double *coords) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, ctx->geo.nAllAtom > 0 && coords, "update_coords: no geometry set");
  cudaMemcpyKind kind = isDevicePtr(coords) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  DFTBP_CUDA(ctx, cudaMemcpyAsync(ctx->geo.coords, coords, sizeof(double) * 3 * ctx->geo.nAllAtom, kind,
                                  ctx->stream));
  if (kind == cudaMemcpyHostToDevice) DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // same pairs, new bond vectors: the records of the schedules are refreshed at their next use; the order
  // (species pair, distance bin) stays until it has drifted too far (ensureSortedPairs)
  ctx->sortMine.stale = ctx->sortLoad.stale = true;
  return 0;
}

}  // extern "C"

namespace dftbp_b200 {
void releaseGeometry(dftbp_b200_ctx *ctx) { freeGeometry(ctx, ctx->geo); }

// (re-)upload the species-pair programs after tables changed
int syncPrograms(dftbp_b200_ctx *ctx) {
  if (!ctx->progsDirty) return 0;
  for (SpPairProg &pg : ctx->progs) {
    if (pg.H.tab == nullptr || pg.S.tab == nullptr)
      return fail(ctx, DFTBP_B200_EINVAL,
                  "SK tables missing: set_sk_table must be called for H and S of every species pair");
    // H and S of one .skf file share the grid: the kernels then form the interpolation window and
    // weights once per pair
    pg.sameGrid = pg.H.nGrid == pg.S.nGrid && pg.H.dist == pg.S.dist && pg.H.method == pg.S.method;
  }
  DFTBP_TRY(uploadNew(ctx, &ctx->dProgs, ctx->progs.data(), ctx->progs.size()));
  ctx->progsDirty = false;
  return 0;
}
}  // namespace dftbp_b200

// common.cuh -- context, error handling and host/device pointer plumbing of libdftbp_b200.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dftbp_b200.h"

namespace dftbp_b200 {

constexpr int kMaxOrb = 16;   // s+p+d+f
constexpr int kMaxInteg = 32; // >= integrals per species pair (spdf-spdf = 30)
constexpr int kWarpsPerCta = 8;

// One equidistant SK table on the device (TSlakoEqGrid, dftb/slakoeqgrid.F90:25-33).
struct SkTableDev {
  const double *tab;   // (nGrid, nInteg) column-major
  const double *tail;  // (3, nInteg): coefficients dd, ee, ff of the 5th-order tail (poly5ToZero,
                       // math/interpolation.F90:31-67; slakoeqgrid.F90:228-240 / :295-307)
  double dist, invDist, rMax;
  double tailX0, tailInvDx;  // tail abscissa xr = (r - tailX0) * tailInvDx
  int nGrid, nInteg, method;
  int mIndM1;  // skEqGridOld: mInd - 1, the first grid interval beyond the tail (slakoeqgrid.F90:268,295)
};

// Everything the pair kernels need for one ordered species pair (sp1 = atom, sp2 = neighbour).
struct SpPairProg {
  SkTableDev H, S;
  int nOrb1, nOrb2, nInteg, nY;
  int nSh1, nSh2;                     // shells of the atom / of the neighbour
  unsigned long long shells1, shells2;  // angular momentum per shell, 4 bits each (shell 0 lowest)
  int canon1, canon2;  // lmax if the species' shells are exactly l = 0..lmax in order (s, sp, spd, spdf), else -1
  int sameGrid;  // the H and the S table have the same grid (length, spacing, method)
  int nShPair;   // nSh1 * nSh2
  // Shell-pair program of the compact evaluator (sk_dmat_eval.cuh; filled by fillShellProgram,
  // sk_table_prep.h).  Shell pairs in the order of rotateH0's loops (atom-1 shells outer):
  //   shPair[i]   = (first element of the sub-block in the column-major block) << 4 | (l1 * 4 + l2)
  //   integDst[k] = element of the block where integral k is parked until its shell pair is rotated
  //                 (the first lmin + 1 elements of that sub-block)
  unsigned short shPair[kMaxInteg];
  unsigned short integDst[kMaxInteg];
  // Two-kernel build (k1_pair_build.cu), species whose shells all have different l ("simple"):
  //   caseOff[c]   = first element of the sub-block of shell-pair type c = l1 * 4 + l2, 0xFFFF if absent
  //   integSlot[k] = position of integral k in the exchange row of a pair, which is ordered by
  //                  shell-pair type (caseSlot<LMAX>, sk_dmat_eval.cuh) instead of by shell loops
  //   shellOfL1[l] = shell of the atom (species 1) with angular momentum l, -1 if none
  alignas(16) unsigned short caseOff[16];
  unsigned char integSlot[kMaxInteg];
  signed char shellOfL1[4];
  int simple;
};

// Device-built schedule of the lane-per-pair kernels (pair_sort.cu): the off-site pairs of one work
// list, counting-sorted by (species pair, distance bin) and padded with -1 so that every aligned
// group of 32 entries (= one warp) holds one species pair and neighbouring table rows.
// One record per schedule slot, written by the sort: everything a lane needs to start on its pair
// without the dependent chain schedule -> iNeighbour -> img2CentCell / iPair / species -> coords.
struct __align__(16) PairRec {
  double vx, vy, vz;  // coords(:,iAtom2) - coords(:,iAtom1)   (nonscc.F90:413)
  int ind;            // iPair(iNeigh, iAtom1); -1 = padding slot
  int sp;             // sp1 + nSpecies * sp2 (0-based): index of the species-pair program
};
static_assert(sizeof(PairRec) == 32, "PairRec must be two 16-byte words");
constexpr int kRecOnsite = 1 << 30;  // PairRec::sp of an on-site block (records in work-list order only)

struct SortedPairs {
  int *items = nullptr;    // per slot: position of the pair in the work list, -1 = padding
  PairRec *recs = nullptr; // per slot, same order as items
  PairRec *recsMem = nullptr;  // per work-list entry (the order of the sparse arrays), on-site blocks included
  long long recsMemCap = 0;
  int *slotOf = nullptr;       // per work-list entry: its slot in the schedule (-1: on-site block)
  long long slotOfCap = 0;
  int2 *atoms = nullptr;   // per slot: (iAtom1, iAtom2f), 0-based (K5: where the gradient goes)
  int *total = nullptr;    // device scalar: padded length (multiple of 32)
  long long capacity = 0;  // allocated entries of items / recs / atoms
  bool valid = false;      // false after set_geometry (new neighbour list) / a new pair partition
  bool stale = false;      // update_coords: same pairs, new bond vectors -- the records must be refreshed
  int refreshes = 0;       // refreshes since the last full sort
};

struct BlockDesc;  // block_desc.cuh

struct Geometry {
  int nAtom = 0, nAllAtom = 0, mNeigh = 0, nCellVec = 0, sparseSize = 0, nOrb = 0;
  double *coords = nullptr;      // (3, nAllAtom)
  int *species = nullptr;        // (nAllAtom)
  int *img2CentCell = nullptr;   // (nAllAtom)
  int *iCellVec = nullptr;       // (nAllAtom)
  double *cellVec = nullptr;     // (3, nCellVec)
  int *nNeighbourSK = nullptr;   // (nAtom)
  int *iNeighbour = nullptr;     // (mNeigh+1, nAtom)
  int *iPair = nullptr;          // (mNeigh+1, nAtom)
  int *iAtomStart = nullptr;     // (nAtom+1)
  int *workAll = nullptr;        // flat (iNeigh + ld*iAtom0) of every block (iNeigh >= 0), all atoms
  int *workMine = nullptr;       // same, this rank's contiguous atom chunk (build_*)
  int *workLoad = nullptr;       // same, this rank's workload-balanced atoms (mulliken, forces)
  BlockDesc *descAll = nullptr;  // resolved index chain of every entry of workAll (block_desc.cuh)
  BlockDesc *descLoad = nullptr; // same for workLoad (== descAll on one rank)
  long long nWorkAll = 0, nWorkMine = 0, nWorkLoad = 0, nPairsMine = 0;
  bool hasDuplicates = false;    // two blocks of one atom fold onto the same dense block
  bool coversAll = true;         // the blocks tile [0, sparseSize) completely (no zero fill needed)
  int firstAtomMine = 0, lastAtomMine = -1;  // 0-based inclusive atom chunk of workMine
  // element -> descriptor index of the TMA-streamed element-wise kernels (flat_stream.cuh): per group of
  // 32 elements of the sparse arrays the descriptor of the block its first element lies in; null when a
  // list is not sorted by offset (those lists take the warp-per-block kernels)
  int *groupFirstAll = nullptr, *groupFirstLoad = nullptr;
};

struct Staging {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct PendingCopy {
  void *host;
  const void *dev;
  size_t bytes;
};

// Device-time accounting per routine, named after the reference's global timers
// (common/environment.F90:96-134) so CPU-vs-GPU tables line up.
enum TimerSlot {
  kTimerSparseH0S = 0,    // "Sparse H0 and S build"          K1
  kTimerSparseToDense,    // "Sparse to dense"                K2a (zero-fill + scatter)
  kTimerDenseToSparse,    // "Dense to sparse"                K2b
  kTimerDensityMatrix,    // "Density matrix creation"        K3
  kTimerMulliken,         // Mulliken (part of "SCC")         K4
  kTimerForceCalc,        // "Force calculation"              K5
  kTimerDiagonalization,  // "Diagonalisation"                cuSOLVER
  kTimerAddShift,         // addShift (part of "SCC")         K6
  kTimerNeighbourList,    // updateNeighbourList + sparse descriptor          N2
  kTimerCount
};

struct TimedSpan {
  cudaEvent_t a, b;
  int slot;
};

}  // namespace dftbp_b200

struct dftbp_b200_ctx {
  int device = 0, rank = 0, nRanks = 1;
  cudaStream_t stream = nullptr;
  bool ownStream = true;
  std::string err;
  long long launches = 0;
  int dmColFirst = 1, dmColLast = 0;            // column panel of the current density-matrix call
  int pairPart = 0, pairParts = 1;              // share of the pair schedule the force calls take (set_pair_partition)
  std::vector<long long> blockStartOfAtom;      // first entry of every atom in workAll / descAll (nAtom + 1)
  std::vector<int> loadInd;                     // sparse offset of every entry of descLoad, then the end of the last one
                                                // (empty for a geometry whose descriptors were built on the device:
                                                // loadElementRange reads the two offsets it needs from descLoad)
  long long rangeLo = -1, rangeHi = -1;         // cache of loadElementRange: work-list range ...
  int rangeELo = 0, rangeEHi = 0;               // ... and the sparse elements it covers
  int smCount = 148;
  size_t smemOptin = 232448;  // max dynamic shared memory per CTA
  size_t smemPerSm = 233472;
  int lMaxAll = 0;            // highest angular momentum of any species
  bool allSimple = false;     // no species has two shells of the same l (two-kernel pair build possible)
  double *pairExchange = nullptr;  // per work-list entry: interpolated integrals of H0 and S (k1_pair_build.cu)
  long long pairExchangeCap = 0;   // allocated doubles
  // What the shift-structure kernel found in the PREVIOUS call with block shifts (pinned, mapped host int,
  // written by one store of that call's span-pipeline kernel -- a copy on the stream cost ~8 us per call:
  // -1 unknown, 0 diagonal blocks, 1 full blocks).  Only a
  // hint: the device flag of the current call decides; the hint picks which launches are queued (full blocks
  // last time: also the dedicated full-block kernel, which is faster than the fallback inside the
  // span-pipeline launch).
  int *hShiftHint = nullptr;
  double *shiftDiag = nullptr;     // (mOrb, nAtom, nSpin): diagonals of the block shifts of the current call, compact
  long long shiftDiagCap = 0;

  // orbitals
  int nSpecies = 0, mShell = 0, mOrb = 0;
  std::vector<int> nShell, angShell, nOrbSpecies;
  int *dNOrbSpecies = nullptr;
  int *dOrbShell = nullptr;  // (mOrb, nSpecies): shell index of each orbital
  // SK programs
  std::vector<dftbp_b200::SpPairProg> progs;  // [sp1 + nSpecies*sp2]
  dftbp_b200::SpPairProg *dProgs = nullptr;
  bool progsDirty = true;
  std::vector<void *> tableAllocs;
  double cutoff = 0.0;

  dftbp_b200::Geometry geo;
  // lane-per-pair schedules (pair_sort.cu); buffers are grow-only and survive set_geometry
  dftbp_b200::SortedPairs sortMine, sortLoad;  // of geo.workMine / geo.workLoad
  int *sortKeys = nullptr;                     // scratch: bin of every work item
  int *sortBins = nullptr;                     // scratch: histogram / cursors
  long long sortKeysCap = 0, sortBinsCap = 0;
  std::vector<dftbp_b200::Staging> staging;
  std::vector<dftbp_b200::PendingCopy> pending;
  bool timing = false;
  std::vector<dftbp_b200::TimedSpan> spans;
  std::vector<cudaEvent_t> eventPool;
  double timerMs[dftbp_b200::kTimerCount] = {0};
  long long timerCalls[dftbp_b200::kTimerCount] = {0};
  void *solverHandle = nullptr;
  void *blasHandle = nullptr;  // cuBLAS, for the triangular solves of the large-n generalized eigensolve
  void *solverWork = nullptr;
  size_t solverWorkBytes = 0;
  int *dInfo = nullptr;
  void *neighbourState = nullptr;  // device-built neighbour list (n2_neighbour_list.cu)
  void *shortGammaState = nullptr; // short-range gamma values of the current list (n4_coulomb.cu)
  void *mixerState = nullptr;      // device Broyden mixer (n4_mixer.cu)
  unsigned long long listGeneration = 0;  // counts update_neighbour_list calls (values derived from a list name theirs)
  // device blocks of the per-geometry arrays, kept across set_geometry calls (poolAlloc / poolRelease):
  // a geometry step re-uploads ~20 arrays of nearly the same sizes, and cudaFree + cudaMalloc of each
  // cost more than everything else in set_geometry (each cudaFree synchronises the device)
  std::multimap<size_t, void *> poolIdle;        // capacity -> block
  std::unordered_map<void *, size_t> poolCap;    // every block of the pool, handed out or idle
};

namespace dftbp_b200 {

inline int fail(dftbp_b200_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) {
    ctx->err = buf;
    // a failed call must not leave copy-back requests behind: the next successful call's finish()
    // would write staging data into the failed caller's (possibly dead) host buffers
    ctx->pending.clear();
  }
  return code;
}

#define DFTBP_CUDA(ctx, call)                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return dftbp_b200::fail(ctx, DFTBP_B200_ECUDA, "%s failed: %s (%s:%d)", #call,            \
                              cudaGetErrorString(e_), __FILE__, __LINE__);                      \
  } while (0)

#define DFTBP_CHECK(ctx, cond, msg)                                                             \
  do {                                                                                          \
    if (!(cond)) return dftbp_b200::fail(ctx, DFTBP_B200_EINVAL, "%s (%s:%d)", msg, __FILE__, __LINE__); \
  } while (0)

#define DFTBP_TRY(call)                 \
  do {                                  \
    int rc_ = (call);                   \
    if (rc_ != 0) return rc_;           \
  } while (0)

inline bool isDevicePtr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Staging slots (grow-only device buffers for host-pointer callers).
enum Slot { kSlotA = 0, kSlotB, kSlotC, kSlotD, kSlotE, kSlotF, kSlotG, kSlotCount };

inline int stagingGet(dftbp_b200_ctx *ctx, int slot, size_t bytes, void **out) {
  if ((int)ctx->staging.size() < kSlotCount) ctx->staging.resize(kSlotCount);
  Staging &s = ctx->staging[slot];
  if (s.bytes < bytes) {
    if (s.ptr) DFTBP_CUDA(ctx, cudaFree(s.ptr));
    s.ptr = nullptr;
    s.bytes = 0;
    cudaError_t e = cudaMalloc(&s.ptr, bytes);
    if (e != cudaSuccess)
      return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc of %zu staging bytes failed: %s", bytes,
                  cudaGetErrorString(e));
    s.bytes = bytes;
  }
  *out = s.ptr;
  return 0;
}

// Device view of an input array (copies H2D when the caller passed a host pointer).
template <class T>
inline int devIn(dftbp_b200_ctx *ctx, int slot, const T *p, size_t count, const T **out) {
  if (p == nullptr) return fail(ctx, DFTBP_B200_EINVAL, "null input pointer");
  if (isDevicePtr(p)) {
    *out = p;
    return 0;
  }
  void *d;
  DFTBP_TRY(stagingGet(ctx, slot, count * sizeof(T), &d));
  DFTBP_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *out = static_cast<const T *>(d);
  return 0;
}

// Device view of an output (or in/out) array; host pointers get a staging buffer whose content
// is copied back by finish().
template <class T>
inline int devOut(dftbp_b200_ctx *ctx, int slot, T *p, size_t count, bool copyIn, T **out) {
  if (p == nullptr) return fail(ctx, DFTBP_B200_EINVAL, "null output pointer");
  if (isDevicePtr(p)) {
    *out = p;
    return 0;
  }
  void *d;
  DFTBP_TRY(stagingGet(ctx, slot, count * sizeof(T), &d));
  if (copyIn)
    DFTBP_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  ctx->pending.push_back({p, d, count * sizeof(T)});
  *out = static_cast<T *>(d);
  return 0;
}

// Copy staged outputs back and make host results visible; no-op (and no sync) for device callers.
inline int finish(dftbp_b200_ctx *ctx) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->pending.clear();
    return fail(ctx, DFTBP_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  }
  if (ctx->pending.empty()) return 0;
  for (const PendingCopy &c : ctx->pending)
    DFTBP_CUDA(ctx, cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->pending.clear();
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// capacities of pool blocks: multiples of a quarter of the largest power of two below the request
inline size_t poolRound(size_t bytes) {
  if (bytes < 512) return 512;
  size_t p = 512;
  while (p * 2 <= bytes) p *= 2;
  const size_t q = p / 4;
  return (bytes + q - 1) / q * q;
}

inline int poolAlloc(dftbp_b200_ctx *ctx, void **out, size_t bytes) {
  const size_t r = poolRound(bytes);
  auto it = ctx->poolIdle.lower_bound(r);
  if (it != ctx->poolIdle.end() && it->first <= 2 * r) {
    *out = it->second;
    ctx->poolIdle.erase(it);
    return 0;
  }
  cudaError_t e = cudaMalloc(out, r);
  if (e != cudaSuccess) {  // give the idle blocks back and try once more
    for (auto &kv : ctx->poolIdle) {
      ctx->poolCap.erase(kv.second);
      cudaFree(kv.second);
    }
    ctx->poolIdle.clear();
    cudaGetLastError();
    e = cudaMalloc(out, r);
  }
  if (e != cudaSuccess)
    return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc(%zu) failed: %s", r, cudaGetErrorString(e));
  ctx->poolCap[*out] = r;
  return 0;
}

inline void poolRelease(dftbp_b200_ctx *ctx, void *p) {
  if (!p) return;
  auto it = ctx->poolCap.find(p);
  if (it == ctx->poolCap.end()) cudaFree(p);  // not from the pool
  else ctx->poolIdle.emplace(it->second, p);
}

inline void poolDestroy(dftbp_b200_ctx *ctx) {
  for (auto &kv : ctx->poolIdle) cudaFree(kv.second);
  ctx->poolIdle.clear();
  ctx->poolCap.clear();
}

template <class T>
inline int uploadNew(dftbp_b200_ctx *ctx, T **dst, const T *src, size_t count) {
  if (*dst) poolRelease(ctx, *dst);
  *dst = nullptr;
  if (count == 0) return 0;
  void *blk = nullptr;
  DFTBP_TRY(poolAlloc(ctx, &blk, count * sizeof(T)));
  *dst = static_cast<T *>(blk);
  cudaMemcpyKind kind = isDevicePtr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  DFTBP_CUDA(ctx, cudaMemcpyAsync(*dst, src, count * sizeof(T), kind, ctx->stream));
  // the source may be a temporary host vector: make the copy complete before returning
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// grow-only buffer for the compact diagonals of block shifts (k5_forces.cu, k6_add_shift.cu)
inline int shiftDiagBuffer(dftbp_b200_ctx *ctx, long long need, double **out) {
  if (ctx->shiftDiagCap < need) {
    if (ctx->shiftDiag) cudaFree(ctx->shiftDiag);
    ctx->shiftDiag = nullptr;
    ctx->shiftDiagCap = 0;
    cudaError_t e = cudaMalloc((void **)&ctx->shiftDiag, sizeof(double) * (size_t)need);
    if (e != cudaSuccess)
      return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc(%lld) for the shift diagonals failed: %s",
                  (long long)(sizeof(double) * need), cudaGetErrorString(e));
    ctx->shiftDiagCap = need;
  }
  *out = ctx->shiftDiag;
  return 0;
}

// first work-list entry of share `part` of nParts (dftbp_b200_set_pair_partition): contiguous ranges
inline long long shareBegin(long long nWork, int part, int nParts) { return nWork * part / nParts; }

// tuning switch read from the environment (tools/gpu_probe.py A/B runs); def when unset.  The
// environment is read once per switch and process, not on every launch.
inline int tuneInt(const char *name, int def) {
  struct Entry { const char *name; int value; };
  static Entry cache[32];
  static int nCache = 0;
  for (int i = 0; i < nCache; ++i)
    if (cache[i].name == name || std::strcmp(cache[i].name, name) == 0) return cache[i].value;
  const char *env = getenv(name);
  const int v = env ? atoi(env) : def;
  if (nCache < 32) cache[nCache++] = Entry{name, v};
  return v;
}

// the same without the cache: switches a test flips inside one process (solver path selection; once per call)
inline int tuneIntNow(const char *name, int def) {
  const char *env = getenv(name);
  return env ? atoi(env) : def;
}

inline int gridFor(const dftbp_b200_ctx *ctx, long long nWarpItems, int warpsPerCta, int ctasPerSm) {
  long long need = (nWarpItems + warpsPerCta - 1) / warpsPerCta;
  long long cap = (long long)ctx->smCount * ctasPerSm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Opt-in shared memory size and resident CTAs per SM of one kernel instantiation, asked from the
// runtime once per (device, size) instead of on every launch: the two calls cost a few microseconds
// of host time each, which showed up as idle gaps between the short kernels of a pair step.
struct LaunchInfo {
  int device = -1;
  size_t smemBytes = ~(size_t)0;
  int perSm = 1;
};

template <class Kernel>
inline int launchInfo(dftbp_b200_ctx *ctx, LaunchInfo &cache, Kernel kernel, int threads, size_t smemBytes, int *perSm) {
  if (cache.device != ctx->device || cache.smemBytes != smemBytes) {
    if (smemBytes > 48 * 1024)
      DFTBP_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    int n = 1;
    DFTBP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smemBytes));
    cache.device = ctx->device;
    cache.smemBytes = smemBytes;
    cache.perSm = n > 0 ? n : 1;
  }
  if (perSm) *perSm = cache.perSm;
  return 0;
}

// RAII span: records an event pair around the launches of one routine when timing is enabled.
struct ScopedTimer {
  dftbp_b200_ctx *ctx;
  TimedSpan span;
  bool on;
  ScopedTimer(dftbp_b200_ctx *c, int slot) : ctx(c), on(c->timing) {
    if (!on) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!ctx->eventPool.empty()) {
        e = ctx->eventPool.back();
        ctx->eventPool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    span.a = get();
    span.b = get();
    span.slot = slot;
    cudaEventRecord(span.a, ctx->stream);
  }
  ~ScopedTimer() {
    if (!on) return;
    cudaEventRecord(span.b, ctx->stream);
    ctx->spans.push_back(span);
  }
};

// implemented in context.cu / solver.cu / pair_sort.cu
int syncPrograms(dftbp_b200_ctx *ctx);
// (re)build the sorted pair schedule of workMine (forLoad = false) or workLoad (true) if stale
int ensureSortedPairs(dftbp_b200_ctx *ctx, bool forLoad, const SortedPairs **out);
void freeSortedPairs(dftbp_b200_ctx *ctx);
void solverDestroy(dftbp_b200_ctx *ctx);
void freeNeighbourList(dftbp_b200_ctx *ctx);
// the arrays of the device neighbour list of the last update_neighbour_list (n2_neighbour_list.cu), for the
// kernels of other translation units that walk it (short-range gamma, n4_coulomb.cu)
struct NeighbourListView {
  bool valid = false;
  int nAtom = 0, nAllAtom = 0, mNeigh = 0;
  const double *coords = nullptr, *neighDist2 = nullptr;
  const int *species = nullptr, *img2CentCell = nullptr, *nNeighbour = nullptr, *iNeighbour = nullptr;
};
NeighbourListView neighbourListView(dftbp_b200_ctx *ctx);
void freeShortGamma(dftbp_b200_ctx *ctx);
void freeMixer(dftbp_b200_ctx *ctx);
// give the arrays of the current geometry back to the block pool (context.cu)
void releaseGeometry(dftbp_b200_ctx *ctx);

}  // namespace dftbp_b200

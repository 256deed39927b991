// flat_stream.cuh -- span pipeline of the element-wise sparse kernels (SCC force weights of K5,
// addShift K6): the sparse arrays streamed through shared memory by the TMA engine.
//
// The warp-per-block kernels of round 1 (block_desc.cuh) kept every byte that was in flight in a
// register: a warp had the 2-3 loads per lane of ONE block outstanding, waited a full HBM round trip
// and went on to the next block -- 45-50 % of the copy bandwidth, bound by latency x registers
// (profiles/r02f_ncu_all_kernels.txt: long_scoreboard 8.3 stalls per issue at 41 % occupancy).
// Here the element space [0, n) of the sparse arrays is cut into spans of kSpan elements, dealt out
// round-robin to the WARPS of a persistent grid.  Every warp runs its own pipeline: lane 0 asks the
// TMA engine for the next span of every input array (cp.async.bulk global -> shared, completion
// counted in bytes on the warp's mbarrier) while the warp works on the current one, so the bytes in
// flight sit in shared memory, not in registers, and no CTA-wide barrier ties the warps together.
// The lanes take the elements of a span from shared memory (stride 1, conflict free), find the
// sparse block of each one from a per-32-element index of the descriptor list (built with the
// geometry) plus a short walk, and write their results with plain coalesced stores.  Blocks need not
// tile the array: an element outside every block is skipped.
//
// Alignment: a bulk copy needs 16-byte aligned addresses and sizes.  Spans start at even elements;
// an array whose base is only 8-byte aligned (the second spin channel of an odd-sized sparse array)
// is copied from its next even address into slot 2 of its row, the element in front of it and an
// odd tail element arrive through 8-byte cp.async: slot(o) = o + phase for element x0 + o.
#pragma once
#include <cstdint>
#include <vector>

#include "block_desc.cuh"

namespace dftbp_b200 {
namespace flat {

constexpr int kSpan = 128;       // elements per span and warp: 1 KB per array
constexpr int kGroup = 32;       // granularity of the element -> descriptor index
constexpr int kRow = kSpan + 4;  // row of one array in a stage (spare slots for the phase shift)
constexpr int kStages = 2;
constexpr int kPerLane = kSpan / 32;
constexpr int kBatch = 2;        // elements of a lane whose loads are issued together

__host__ __device__ constexpr int warpsFor(int nIn) { return nIn <= 3 ? 32 : nIn <= 6 ? 16 : 8; }

template <int NIN, int WARPS>
struct alignas(128) Smem {
  double in[WARPS][kStages][NIN][kRow];
  unsigned long long full[WARPS][kStages];
};

// Host side: groupFirst[g] = the last block that starts at or before element 32 g (0 if none), for
// the descriptor list d sorted by offset; false if the list is not sorted.
inline bool buildGroupIndex(const BlockDesc *d, long long nBlocks, long long n, std::vector<int> &groupFirst) {
  if (nBlocks == 0) return false;
  for (long long b = 1; b < nBlocks; ++b)
    if (d[b].ind <= d[b - 1].ind) return false;
  const long long nGroups = (n + kGroup - 1) / kGroup;
  groupFirst.assign((size_t)nGroups, 0);
  long long j = 0;
  for (long long g = 0; g < nGroups; ++g) {
    const long long x0 = g * kGroup;
    while (j + 1 < nBlocks && d[j + 1].ind <= x0) ++j;
    groupFirst[(size_t)g] = (int)j;
  }
  return true;
}

#ifdef __CUDACC__

__device__ __forceinline__ unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarWait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

// m2[n] for n = 1 .. kMaxOrb: e / n = e * m2[n] >> 20 for e < 256 (block elements -> column); call with
// all threads of the CTA, m2 in shared memory
__device__ __forceinline__ void initDivTable(unsigned *m2) {
  if (threadIdx.x <= kMaxOrb) m2[threadIdx.x] = threadIdx.x > 0 ? (1048576u + threadIdx.x - 1u) / threadIdx.x : 0u;
  __syncthreads();
}

// One element of a span as the kernels see it: element x of the sparse arrays = element e of the block
// described by (lo, hi) = the two 16-byte words of its BlockDesc; v[a] = src[a][x].  Elements outside
// every block (ok = false) carry e = 0 of a neighbouring block: loads through them stay inside the
// arrays, their results are not stored.
template <int NIN>
struct Elem {
  int x, e;
  int4 lo, hi;
  double v[NIN];
  bool ok;
};

// body(const Elem<NIN> (&el)[kBatch]): kBatch elements of a lane at once, so that a kernel can issue
// the dependent loads of all of them (shift values: L1 / L2 hits) before the first use.
// All index arithmetic is 32 bit: the sparse arrays have fewer than 2^31 elements (the C ABI passes
// sparseSize as int).
template <int NIN, int WARPS, class Body>
__device__ __forceinline__ void streamSpans(Smem<NIN, WARPS> &sm, const double *const (&src)[NIN], int n,
                                            const BlockDesc *__restrict__ desc, int nBlocks,
                                            const int *__restrict__ groupFirst, int spanLo, int spanHi, Body &&body) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nSpans = min(spanHi, (n + kSpan - 1) / kSpan);  // spans [spanLo, spanHi) of the arrays
  const int stride = (int)gridDim.x * WARPS;
  int phase[NIN];
#pragma unroll
  for (int a = 0; a < NIN; ++a) phase[a] = (int)(((uintptr_t)src[a] >> 3) & 1);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sm.full[w][s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const unsigned inBase = smemAddr(&sm.in[w][0][0][0]), barBase = smemAddr(&sm.full[w][0]);
  constexpr unsigned kStageBytes = 8u * NIN * kRow, kRowBytes = 8u * kRow;

  // ask for span `sp` (if it exists) into stage `st` of this warp: lane 0 the bulk copies, lanes 1 ..
  // 2 NIN the head (the element in front of an odd base) and an odd tail element with 8-byte cp.async;
  // every lane commits one cp.async group
  auto request = [&](int sp, int st) {
    if (sp < nSpans && lane <= 2 * NIN) {
      const int x0 = sp * kSpan;
      const int cnt = min(kSpan, n - x0);
      const unsigned stageS = inBase + kStageBytes * (unsigned)st;
      if (lane == 0) {
        unsigned bytes = 0;
#pragma unroll
        for (int a = 0; a < NIN; ++a) bytes += 8u * (unsigned)max(0, (cnt - phase[a]) & ~1);
        const unsigned bar = barBase + 8u * (unsigned)st;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
        for (int a = 0; a < NIN; ++a) {
          const int m = max(0, (cnt - phase[a]) & ~1);
          if (m > 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             stageS + kRowBytes * a + 16u * (unsigned)phase[a]),
                         "l"(src[a] + x0 + phase[a]), "r"(8u * (unsigned)m), "r"(bar)
                         : "memory");
        }
      } else {
#pragma unroll
        for (int a = 0; a < NIN; ++a) {
          const int m = max(0, (cnt - phase[a]) & ~1);
          if (lane == 1 + 2 * a && phase[a] != 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stageS + kRowBytes * a + 8u), "l"(src[a] + x0)
                         : "memory");
          const int o = phase[a] + m;
          if (lane == 2 + 2 * a && o < cnt)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stageS + kRowBytes * a + 8u * (unsigned)(o + phase[a])),
                         "l"(src[a] + x0 + o)
                         : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int first = spanLo + (int)blockIdx.x * WARPS + w;
#pragma unroll
  for (int s = 0; s < kStages; ++s) request(first + s * stride, s);

  const int nbM1 = nBlocks - 1;
  const int *__restrict__ descInd = reinterpret_cast<const int *>(desc);  // ind of block j at descInd[8 j]
  int k = 0;
  for (int sp = first; sp < nSpans; sp += stride, ++k) {
    const int st = k % kStages;
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    mbarWait(barBase + 8u * (unsigned)st, (unsigned)((k / kStages) & 1));
    __syncwarp();  // the head / tail copies of the other lanes are visible
    const int x0 = sp * kSpan;
    const int cnt = min(kSpan, n - x0);
    const double *__restrict__ stage = &sm.in[w][st][0][0];
    const int *__restrict__ gf = groupFirst + (x0 >> 5);
#pragma unroll 1
    for (int u0 = 0; u0 < kPerLane; u0 += kBatch) {
      if (32 * u0 >= cnt) break;  // warp-uniform
      Elem<NIN> el[kBatch];
      int j[kBatch], o[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int og = 32 * (u0 + u);                 // first element of the group of 32
        o[u] = min(og + lane, cnt - 1);               // (clamped: lanes behind the end repeat the last element)
        j[u] = __ldg(gf + (min(og, cnt - 1) >> 5));
        el[u].x = x0 + o[u];
      }
      // the block of element x: walk from the block of the group's first element
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
        while (j[u] < nbM1 && __ldg(descInd + 8 * (j[u] + 1)) <= el[u].x) ++j[u];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        el[u].lo = __ldg(reinterpret_cast<const int4 *>(desc + j[u]));
        el[u].hi = __ldg(reinterpret_cast<const int4 *>(desc + j[u]) + 1);
        const int e = el[u].x - el[u].lo.x;
        const int nElem = (el[u].lo.w & 0xFF) * (el[u].lo.w >> 8);
        el[u].ok = 32 * (u0 + u) + lane < cnt && e >= 0 && e < nElem;
        el[u].e = el[u].ok ? e : 0;
#pragma unroll
        for (int a = 0; a < NIN; ++a) el[u].v[a] = stage[kRow * a + o[u] + phase[a]];
      }
      body(el);
    }
    __syncwarp();  // every lane has taken its values out of the stage
    request(sp + kStages * stride, st);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

#endif  // __CUDACC__

}  // namespace flat
}  // namespace dftbp_b200

// k3_density_matrix.cu -- fullDensityMatrix_{real,cmplx} / fullEnergyDensityMatrix_{real,cmplx}
// (dftb/densitymatrix.F90:530-806): dm = lower( C(:,1:nLevels) diag(w) C(:,1:nLevels)^H ).
//
// The reference scales the eigenvector columns by sqrt(f), calls BLAS xSYRK/xHERK and un-scales
// (or uses a shifted two-pass form when a weight is <= 0).  Here the weight is folded into the B
// operand while it is loaded from shared memory, so weights of any sign need one pass, the
// eigenvectors are never modified and no sqrt is taken.
//
// FP64 tensor-core contraction: DMMA (mma.sync.m8n8k4.f64; the only FP64 MMA shape sm_100a has
// in hardware -- m16n8k16 etc. lower to the same DMMA.8x8x4 SASS).  Persistent CTAs walk the
// lower-triangular tile list in super-blocks so concurrently running CTAs share row panels in L2;
// operands stream global -> shared through a 4-stage cp.async pipeline; the eigenvectors are
// column-major, i.e. "MN-major" for both operands, and the shared tiles use a pitch = 4 (mod 16)
// doubles so the (k = lane%4, mn = lane/4) fragment loads are bank-conflict free.
//
// Roofline: FP64 tensor pipe; flops = n^2 * nLevels (one triangle, FMA = 2), x4 for complex.
#include <cstdlib>

#include "common.cuh"

using namespace dftbp_b200;

namespace {

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cpAsync(void *smem, const void *gmem, int bytes, int srcBytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(srcBytes));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(srcBytes));
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// i-th lower-triangular tile in super-block order: super-blocks of G x G tiles, row-major over the
// lower triangle of super-blocks, row-major inside.  Returns false when (ti, tj) is above the diagonal.
struct TileOrder {
  int nT, G, nSB;
  int tjFirst, tjLast;  // column panel (in tiles, inclusive): tiles outside are skipped (multi-GPU split of one dm)
};

__device__ __forceinline__ bool tileAt(const TileOrder &o, long long idx, int *ti, int *tj) {
  // idx enumerates ALL positions of all lower super-blocks (including padding); callers skip misses
  const int per = o.G * o.G;
  const long long sb = idx / per;
  const int in = (int)(idx - sb * per);
  // sb -> (si, sj), si >= sj
  int si = (int)((sqrt(8.0 * (double)sb + 1.0) - 1.0) * 0.5);
  while ((long long)(si + 1) * (si + 2) / 2 <= sb) ++si;
  while ((long long)si * (si + 1) / 2 > sb) --si;
  const int sj = (int)(sb - (long long)si * (si + 1) / 2);
  *ti = si * o.G + in / o.G;
  *tj = sj * o.G + in % o.G;
  return *ti < o.nT && *tj <= *ti && *tj >= o.tjFirst && *tj <= o.tjLast;
}

// ------------------------------------------------------------------------------------------------
// real:  dm(i,j) = sum_k w_k C(i,k) C(j,k),  i >= j
// ------------------------------------------------------------------------------------------------
template <int BM, int BK, int STAGES, int WARPS_M, int WARPS_N, int VEC>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N)
k3_dm_real(const double *__restrict__ C, int ldc, const double *__restrict__ wgt, int n,
           const int *__restrict__ nLevels, double *__restrict__ dm, int ldd, TileOrder order,
           unsigned long long *tileCounter) {
  const int K = *nLevels;
  constexpr int NT = 32 * WARPS_M * WARPS_N;
  constexpr int PITCH = BM + 4;
  constexpr int WM = BM / WARPS_M, WN = BM / WARPS_N;  // warp tile
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(16) double smem[];
  double *As = smem;                             // [STAGES][BK][PITCH]
  double *Bs = As + STAGES * BK * PITCH;         // [STAGES][BK][PITCH]
  double *Ws = Bs + STAGES * BK * PITCH;         // [STAGES][BK]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int fk = lane & 3, fmn = lane >> 2;
  const int nK = (K + BK - 1) / BK;
  const long long nIdx = (long long)order.nSB * (order.nSB + 1) / 2 * order.G * order.G;

  __shared__ long long nextSlot;
  for (;;) {
    // dynamic tile scheduler: one atomic per tile keeps all SMs busy to the end (diagonal
    // super-blocks are half empty, so a static round-robin leaves SMs idle for ~30% of the time)
    __syncthreads();
    if (tid == 0) nextSlot = atomicAdd(tileCounter, 1ULL);
    __syncthreads();
    const long long idx = nextSlot;
    if (idx >= nIdx) break;
    int ti, tj;
    if (!tileAt(order, idx, &ti, &tj)) continue;
    const int i0 = ti * BM, j0 = tj * BM;

    auto loadStage = [&](int stage, int kt) {
      const int k0 = kt * BK;
      constexpr int CH = BM / VEC;  // chunks per k column
      for (int c = tid; c < CH * BK; c += NT) {
        const int col = c / CH, ch = c - col * CH;
        const int k = k0 + col;
        {
          const int i = i0 + ch * VEC;
          int valid = (k < K) ? min(max(n - i, 0), VEC) : 0;
          const double *src = valid > 0 ? C + (size_t)i + (size_t)ldc * k : C;
          cpAsync(As + (stage * BK + col) * PITCH + ch * VEC, src, VEC * 8, valid * 8);
        }
        {
          const int j = j0 + ch * VEC;
          int valid = (k < K) ? min(max(n - j, 0), VEC) : 0;
          const double *src = valid > 0 ? C + (size_t)j + (size_t)ldc * k : C;
          cpAsync(Bs + (stage * BK + col) * PITCH + ch * VEC, src, VEC * 8, valid * 8);
        }
      }
      if (tid < BK) {
        const int k = k0 + tid;
        cpAsync(Ws + stage * BK + tid, k < K ? wgt + k : wgt, 8, k < K ? 8 : 0);
      }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nK) loadStage(s, s);
      cpAsyncCommit();
    }
    for (int kt = 0; kt < nK; ++kt) {
      cpAsyncWait<STAGES - 2>();
      __syncthreads();
      if (kt + STAGES - 1 < nK) loadStage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
      cpAsyncCommit();
      const int st = kt % STAGES;
      const double *a_s = As + st * BK * PITCH + wm * WM + fmn;
      const double *b_s = Bs + st * BK * PITCH + wn * WN + fmn;
      const double *w_s = Ws + st * BK;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        const int kr = kk * 4 + fk;
        const double wk = w_s[kr];
        double a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = a_s[kr * PITCH + mi * 8];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = b_s[kr * PITCH + ni * 8] * wk;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
    cpAsyncWait<0>();
    __syncthreads();

    // epilogue: lane holds D(row = fmn, col = 2*fk + {0,1}) of each 8x8 tile
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int i = i0 + wm * WM + mi * 8 + fmn;
        const int j = j0 + wn * WN + ni * 8 + 2 * fk;
        if (i < n) {
          if (j < n && i >= j) dm[(size_t)i + (size_t)ldd * j] = acc[mi][ni][0];
          if (j + 1 < n && i >= j + 1) dm[(size_t)i + (size_t)ldd * (j + 1)] = acc[mi][ni][1];
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// real, TMA-fed: the same tile walk and fragment layout as k3_dm_real, but the operand panels come in
// through the bulk-copy (TMA) engine and the k loop has no CTA-wide barrier.  One producer warp asks
// for the BK columns of the A and the B panel of a stage with one cp.async.bulk per column (BM rows =
// 512 bytes each, landing at the padded pitch that keeps the fragment loads conflict free -- a
// tensor-map box would be dense in shared memory and 4-way conflicted, so the copies are the
// non-tensor form); their bytes complete the stage's `full` mbarrier.  The consumer warps wait on
// `full`, issue their DMMAs and arrive on the stage's `empty` mbarrier, which is all the producer waits
// for before it reuses the stage: the warps of a CTA drift apart by up to STAGES k-tiles instead of
// meeting at a __syncthreads every 16 levels.  Diagonal tiles load one panel only.
// Needs 16-byte aligned columns: even ld, even n, aligned base (else k3_dm_real).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned k3SmemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void k3MbarWait(unsigned bar, unsigned parity) {
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

template <int BM, int BK, int STAGES, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(32 * (WARPS_M * WARPS_N + 1))
k3_dm_real_tma(const double *__restrict__ C, int ldc, const double *__restrict__ wgt, int n,
               const int *__restrict__ nLevels, double *__restrict__ dm, int ldd, TileOrder order,
               unsigned long long *tileCounter) {
  static_assert(BK <= 16, "one lane of the producer warp per (panel, column)");
  const int K = *nLevels;
  constexpr int NCW = WARPS_M * WARPS_N;  // consumer warps; warp NCW is the producer
  constexpr int PITCH = BM + 4;
  constexpr int WM = BM / WARPS_M, WN = BM / WARPS_N;
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(16) double smem[];
  double *As = smem;                      // [STAGES][BK][PITCH]
  double *Bs = As + STAGES * BK * PITCH;  // [STAGES][BK][PITCH]
  double *Ws = Bs + STAGES * BK * PITCH;  // [STAGES][BK]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(Ws + STAGES * BK);  // full[STAGES], empty[STAGES]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nK = (K + BK - 1) / BK;
  const long long nIdx = (long long)order.nSB * (order.nSB + 1) / 2 * order.G * order.G;
  const unsigned fullBar = k3SmemAddr(bars), emptyBar = k3SmemAddr(bars + STAGES);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fullBar + 8u * s) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(emptyBar + 8u * s), "r"(NCW) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __shared__ long long nextSlot;
  unsigned it = 0;  // k-tiles handled so far by this CTA (all warps count alike): stage = it % STAGES
  for (;;) {
    __syncthreads();
    if (tid == 0) nextSlot = atomicAdd(tileCounter, 1ULL);
    __syncthreads();
    const long long idx = nextSlot;
    if (idx >= nIdx) break;
    int ti, tj;
    if (!tileAt(order, idx, &ti, &tj)) continue;
    const int i0 = ti * BM, j0 = tj * BM;
    const bool diag = ti == tj;

    if (warp == NCW) {
      // ---------------- producer ----------------
      const unsigned rowsA = 8u * (unsigned)min(BM, n - i0), rowsB = diag ? 0u : 8u * (unsigned)min(BM, n - j0);
      const unsigned bytes = (unsigned)BK * (rowsA + rowsB);
      const bool isB = lane >= 16;
      const int col = lane & 15;
      const unsigned myBytes = isB ? rowsB : rowsA;
      const double *base = C + (isB ? j0 : i0);
      const unsigned dst0 = k3SmemAddr(isB ? Bs : As) + 8u * (unsigned)(col * PITCH);
      for (int kt = 0; kt < nK; ++kt, ++it) {
        const unsigned st = it % STAGES;
        if (it >= (unsigned)STAGES) k3MbarWait(emptyBar + 8u * st, ((it / STAGES) & 1u) ^ 1u);
        const int k = kt * BK + col;
        // levels behind the last one: a valid column with weight zero
        if (lane < BK) Ws[st * BK + lane] = k < K ? __ldg(wgt + k) : 0.0;
        __syncwarp();
        const unsigned bar = fullBar + 8u * st;
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        __syncwarp();
        if (col < BK && myBytes > 0)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           dst0 + 8u * (unsigned)(st * BK * PITCH)),
                       "l"(base + (size_t)ldc * (size_t)min(k, K - 1)), "r"(myBytes), "r"(bar)
                       : "memory");
      }
    } else {
      // ---------------- consumers ----------------
      const int wm = warp % WARPS_M, wn = warp / WARPS_M;
      const int fk = lane & 3, fmn = lane >> 2;
      double acc[MI][NI][2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      const double *bPanel = diag ? As : Bs;
      for (int kt = 0; kt < nK; ++kt, ++it) {
        const unsigned st = it % STAGES;
        k3MbarWait(fullBar + 8u * st, (it / STAGES) & 1u);
        const double *a_s = As + st * BK * PITCH + wm * WM + fmn;
        const double *b_s = bPanel + st * BK * PITCH + wn * WN + fmn;
        const double *w_s = Ws + st * BK;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
          const int kr = kk * 4 + fk;
          const double wk = w_s[kr];
          double a[MI], b[NI];
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) a[mi] = a_s[kr * PITCH + mi * 8];
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) b[ni] = b_s[kr * PITCH + ni * 8] * wk;
#pragma unroll
          for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(emptyBar + 8u * st) : "memory");
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const int i = i0 + wm * WM + mi * 8 + fmn;
          const int j = j0 + wn * WN + ni * 8 + 2 * fk;
          if (i < n) {
            if (j < n && i >= j) dm[(size_t)i + (size_t)ldd * j] = acc[mi][ni][0];
            if (j + 1 < n && i >= j + 1) dm[(size_t)i + (size_t)ldd * (j + 1)] = acc[mi][ni][1];
          }
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// complex:  dm(i,j) = sum_k w_k C(i,k) conj(C(j,k)),  i >= j, imag(diag) = 0 (as xHERK)
//   re += ar*(w br) + ai*(w bi);  im += ai*(w br) + ar*(-w bi)
// ------------------------------------------------------------------------------------------------
template <int BM, int BK, int STAGES, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N)
k3_dm_cplx(const double2 *__restrict__ C, int ldc, const double *__restrict__ wgt, int n,
           const int *__restrict__ nLevels, double2 *__restrict__ dm, int ldd, TileOrder order,
           unsigned long long *tileCounter) {
  const int K = *nLevels;
  constexpr int NT = 32 * WARPS_M * WARPS_N;
  constexpr int PITCH = BM + 2;  // in complex (16-byte) units: = 2 (mod 8) for conflict-free LDS.128
  constexpr int WM = BM / WARPS_M, WN = BM / WARPS_N;
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(16) double smem[];
  double2 *As = reinterpret_cast<double2 *>(smem);
  double2 *Bs = As + STAGES * BK * PITCH;
  double *Ws = reinterpret_cast<double *>(Bs + STAGES * BK * PITCH);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int fk = lane & 3, fmn = lane >> 2;
  const int nK = (K + BK - 1) / BK;
  const long long nIdx = (long long)order.nSB * (order.nSB + 1) / 2 * order.G * order.G;

  __shared__ long long nextSlot;
  for (;;) {
    // dynamic tile scheduler: one atomic per tile keeps all SMs busy to the end (diagonal
    // super-blocks are half empty, so a static round-robin leaves SMs idle for ~30% of the time)
    __syncthreads();
    if (tid == 0) nextSlot = atomicAdd(tileCounter, 1ULL);
    __syncthreads();
    const long long idx = nextSlot;
    if (idx >= nIdx) break;
    int ti, tj;
    if (!tileAt(order, idx, &ti, &tj)) continue;
    const int i0 = ti * BM, j0 = tj * BM;

    auto loadStage = [&](int stage, int kt) {
      const int k0 = kt * BK;
      for (int c = tid; c < BM * BK; c += NT) {
        const int col = c / BM, ch = c - col * BM;
        const int k = k0 + col;
        const int i = i0 + ch, j = j0 + ch;
        const bool va = k < K && i < n, vb = k < K && j < n;
        cpAsync(As + (stage * BK + col) * PITCH + ch, va ? C + (size_t)i + (size_t)ldc * k : C, 16, va ? 16 : 0);
        cpAsync(Bs + (stage * BK + col) * PITCH + ch, vb ? C + (size_t)j + (size_t)ldc * k : C, 16, vb ? 16 : 0);
      }
      if (tid < BK) {
        const int k = k0 + tid;
        cpAsync(Ws + stage * BK + tid, k < K ? wgt + k : wgt, 8, k < K ? 8 : 0);
      }
    };

    double re[MI][NI][2], im[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) re[mi][ni][0] = re[mi][ni][1] = im[mi][ni][0] = im[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nK) loadStage(s, s);
      cpAsyncCommit();
    }
    for (int kt = 0; kt < nK; ++kt) {
      cpAsyncWait<STAGES - 2>();
      __syncthreads();
      if (kt + STAGES - 1 < nK) loadStage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
      cpAsyncCommit();
      const int st = kt % STAGES;
      const double2 *a_s = As + st * BK * PITCH + wm * WM + fmn;
      const double2 *b_s = Bs + st * BK * PITCH + wn * WN + fmn;
      const double *w_s = Ws + st * BK;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        const int kr = kk * 4 + fk;
        const double wk = w_s[kr];
        double2 a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = a_s[kr * PITCH + mi * 8];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const double2 v = b_s[kr * PITCH + ni * 8];
          b[ni] = make_double2(v.x * wk, v.y * wk);
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) {
            dmma884(re[mi][ni][0], re[mi][ni][1], a[mi].x, b[ni].x);
            dmma884(re[mi][ni][0], re[mi][ni][1], a[mi].y, b[ni].y);
            dmma884(im[mi][ni][0], im[mi][ni][1], a[mi].y, b[ni].x);
            dmma884(im[mi][ni][0], im[mi][ni][1], a[mi].x, -b[ni].y);
          }
      }
    }
    cpAsyncWait<0>();
    __syncthreads();

#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int i = i0 + wm * WM + mi * 8 + fmn;
        const int j = j0 + wn * WN + ni * 8 + 2 * fk;
        if (i < n) {
          if (j < n && i >= j)
            dm[(size_t)i + (size_t)ldd * j] = make_double2(re[mi][ni][0], i == j ? 0.0 : im[mi][ni][0]);
          if (j + 1 < n && i >= j + 1)
            dm[(size_t)i + (size_t)ldd * (j + 1)] = make_double2(re[mi][ni][1], i == j + 1 ? 0.0 : im[mi][ni][1]);
        }
      }
  }
}

// nLevels = index of the last |filling| >= epsilon (densitymatrix.F90:549-554); computed on the
// device so that device-resident fillings need no host round trip.  Result in *out (>= 1).
__global__ void k3_count_levels(const double *filling, int n, int *out) {
  __shared__ int best;
  if (threadIdx.x == 0) best = 1;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (fabs(filling[i]) >= 2.220446049250313e-16) mine = max(mine, i + 1);
  atomicMax(&best, mine);
  __syncthreads();
  if (threadIdx.x == 0) *out = best;
}

TileOrder makeOrder(int n, int BM, int colFirst, int colLast) {
  TileOrder o;
  o.nT = (n + BM - 1) / BM;
  o.G = 12;
  o.nSB = (o.nT + o.G - 1) / o.G;
  o.tjFirst = (colFirst - 1) / BM;  // 1-based inclusive dense columns -> tiles that overlap them
  o.tjLast = (colLast - 1) / BM;
  return o;
}

template <class Kern>
int setSmem(dftbp_b200_ctx *ctx, Kern k, size_t bytes) {
  DFTBP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int BM, int BK, int ST, int WMn, int WNn, int VEC>
int launchReal(dftbp_b200_ctx *ctx, int ctasPerSm, const double *dC, int ld, const double *dW, int n,
               const int *K, double *dDm, int ldd, unsigned long long *counter) {
  const size_t smemB = sizeof(double) * (2 * ST * BK * (BM + 4) + ST * BK);
  const TileOrder o = makeOrder(n, BM, ctx->dmColFirst, ctx->dmColLast);
  const long long nTiles = (long long)o.nT * (o.nT + 1) / 2;
  const int grid = (int)std::min<long long>(nTiles, (long long)ctasPerSm * ctx->smCount);
  auto kern = k3_dm_real<BM, BK, ST, WMn, WNn, VEC>;
  DFTBP_TRY(setSmem(ctx, kern, smemB));
  kern<<<grid, 32 * WMn * WNn, smemB, ctx->stream>>>(dC, ld, dW, n, K, dDm, ldd, o, counter);
  return 0;
}

template <int BM, int BK, int ST, int WMn, int WNn>
int launchRealTma(dftbp_b200_ctx *ctx, int ctasPerSm, const double *dC, int ld, const double *dW, int n,
                  const int *K, double *dDm, int ldd, unsigned long long *counter) {
  const size_t smemB = sizeof(double) * (2 * ST * BK * (BM + 4) + ST * BK) + sizeof(unsigned long long) * 2 * ST;
  const TileOrder o = makeOrder(n, BM, ctx->dmColFirst, ctx->dmColLast);
  const long long nTiles = (long long)o.nT * (o.nT + 1) / 2;
  const int grid = (int)std::min<long long>(nTiles, (long long)ctasPerSm * ctx->smCount);
  auto kern = k3_dm_real_tma<BM, BK, ST, WMn, WNn>;
  DFTBP_TRY(setSmem(ctx, kern, smemB));
  kern<<<grid, 32 * (WMn * WNn + 1), smemB, ctx->stream>>>(dC, ld, dW, n, K, dDm, ldd, o, counter);
  return 0;
}

// tile configuration of the real kernel for (n, ld, alignment); K and the tile counter live on the device
int launchRealAuto(dftbp_b200_ctx *ctx, const double *dC, int ld, const double *dW, int n, const int *K, double *dDm,
                   int ldd, unsigned long long *counter) {
  const bool aligned = (ld % 2 == 0) && ((reinterpret_cast<size_t>(dC) & 15) == 0);
  // tuning hook (development only): DFTBP_B200_K3_VARIANT selects a tile configuration
  const char *env = getenv("DFTBP_B200_K3_VARIANT");
  const int variant = env ? atoi(env) : 0;
  // TMA-fed kernel: 16-byte aligned columns of C (bulk copies) -- even ld, even n, aligned base
  const bool tmaOk = aligned && n % 2 == 0;
  if (tmaOk && (variant == 0 || variant >= 30)) {
    if (variant == 30 || (variant == 0 && n < 4096)) {
      // small matrices: 64 x 64 tiles, three CTAs per SM (more tiles than CTA slots from n = 1400 on)
      DFTBP_TRY((launchRealTma<64, 16, 4, 2, 2>(ctx, 3, dC, ld, dW, n, K, dDm, ldd, counter)));
    } else if (variant == 31) {
      DFTBP_TRY((launchRealTma<64, 16, 3, 2, 2>(ctx, 4, dC, ld, dW, n, K, dDm, ldd, counter)));
    } else if (variant == 34) {
      DFTBP_TRY((launchRealTma<128, 8, 6, 4, 2>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
    } else if (variant == 35) {
      DFTBP_TRY((launchRealTma<128, 16, 4, 4, 2>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
    } else if (variant == 36) {
      DFTBP_TRY((launchRealTma<128, 8, 8, 2, 4>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
    } else {
      // 128 x 128 tiles, eight consumer warps of 64 x 32 + the producer warp, one CTA per SM: 34.3 TFLOP/s
      // at n = 31744 (93 % of cuBLAS DGEMM; 64 x 64 tiles 32.7, the cp.async kernel of round 1 30.5).
      // Four stages: the same tile with three or six stages ran 14 times slower (2.4 TFLOP/s, 32 us =
      // 2^16 cycles per k-tile -- the time limit of mbarrier.try_wait, i.e. waiters that were only released
      // by their own time-out; not understood, results correct; profiles/r2b_k3_tma.md)
      DFTBP_TRY((launchRealTma<128, 16, 4, 2, 4>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
    }
  } else if (n < 1024) {
    DFTBP_TRY((launchReal<64, 16, 3, 2, 2, 1>(ctx, 2, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (!aligned) {
    DFTBP_TRY((launchReal<128, 16, 4, 2, 4, 1>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 1) {
    DFTBP_TRY((launchReal<128, 16, 4, 4, 4, 2>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 2) {
    DFTBP_TRY((launchReal<128, 32, 3, 2, 4, 2>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 3) {
    DFTBP_TRY((launchReal<128, 16, 4, 2, 4, 2>(ctx, 1, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 6) {
    DFTBP_TRY((launchReal<64, 16, 3, 2, 2, 2>(ctx, 4, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 7) {
    DFTBP_TRY((launchReal<64, 16, 4, 2, 2, 2>(ctx, 2, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 8) {
    DFTBP_TRY((launchReal<64, 32, 3, 2, 2, 2>(ctx, 2, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 9) {
    DFTBP_TRY((launchReal<64, 16, 4, 4, 2, 2>(ctx, 3, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else if (variant == 10) {
    DFTBP_TRY((launchReal<64, 8, 6, 2, 2, 2>(ctx, 4, dC, ld, dW, n, K, dDm, ldd, counter)));
  } else {
    // 64x64 tiles, 4 warps of 32x32, three CTAs per SM: independent barrier domains hide the
    // per-k-tile __syncthreads better than one 128x128 CTA (measured 30.3 vs 26.3 TFLOP/s)
    DFTBP_TRY((launchReal<64, 16, 4, 2, 2, 2>(ctx, 3, dC, ld, dW, n, K, dDm, ldd, counter)));
  }
  return 0;
}

__global__ void k3_set_levels(int K, int *out) { *out = K; }

}  // namespace

namespace dftbp_b200 {
// dm(ldd, n) = lower( C(:, 1:K) diag(w) C(:, 1:K)^T ) with C(ldc, K), w(K) on the device and K given by the host
// (the Ewald structure-factor form of n4_coulomb.cu: K may exceed n); all of dm is zero-filled first.
int weightedGramLower(dftbp_b200_ctx *ctx, const double *dC, int ldc, const double *dW, int n, int K, double *dDm,
                      int ldd) {
  DFTBP_CHECK(ctx, dC && dW && dDm && n > 0 && K > 0 && ldc >= n && ldd >= n, "weightedGramLower: bad arguments");
  ctx->dmColFirst = 1;
  ctx->dmColLast = n;
  DFTBP_CUDA(ctx, cudaMemsetAsync(dDm, 0, sizeof(double) * (size_t)ldd * n, ctx->stream));
  k3_set_levels<<<1, 1, 0, ctx->stream>>>(K, ctx->dInfo);
  unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->dInfo + 2);
  DFTBP_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
  DFTBP_TRY(launchRealAuto(ctx, dC, ldc, dW, n, ctx->dInfo, dDm, ldd, counter));
  ctx->launches += 2;
  DFTBP_CUDA(ctx, cudaGetLastError());
  return 0;
}
}  // namespace dftbp_b200

namespace {

int densityImpl(dftbp_b200_ctx *ctx, void *dm, const void *eigvecs, int n, int ld, const double *filling,
                const double *weights, bool cplx, int colFirst = 1, int colLast = 0) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, dm && eigvecs && filling && n > 0 && ld >= n, "density_matrix: bad arguments");
  if (colLast <= 0) colLast = n;
  DFTBP_CHECK(ctx, colFirst >= 1 && colFirst <= colLast && colLast <= n, "density_matrix: bad column panel");
  ctx->dmColFirst = colFirst;
  ctx->dmColLast = colLast;
  const size_t el = cplx ? 2 : 1;
  const double *dC, *dF, *dW;
  double *dDm;
  DFTBP_TRY(devIn(ctx, kSlotA, (const double *)eigvecs, (size_t)ld * n * el, &dC));
  DFTBP_TRY(devIn(ctx, kSlotC, filling, (size_t)n, &dF));
  dW = dF;
  if (weights) DFTBP_TRY(devIn(ctx, kSlotD, weights, (size_t)n, &dW));
  DFTBP_TRY(devOut(ctx, kSlotB, (double *)dm, (size_t)ld * n * el, false, &dDm));
  {
  ScopedTimer timer(ctx, kTimerDensityMatrix);
  // dm(:,:) = 0 (densitymatrix.F90:547); the kernel then fills the lower triangle
  DFTBP_CUDA(ctx, cudaMemsetAsync(dDm, 0, sizeof(double) * ld * n * el, ctx->stream));
  k3_count_levels<<<1, 256, 0, ctx->stream>>>(dF, n, ctx->dInfo);
  ctx->launches += 1;  // k3_count_levels
  const int *K = ctx->dInfo;  // stays on the device: no host round trip inside the SCC loop
  unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->dInfo + 2);
  DFTBP_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
  if (!cplx) {
    DFTBP_TRY(launchRealAuto(ctx, dC, ld, dW, n, K, dDm, ld, counter));
  } else {
    constexpr int BM = 64, BK = 8, ST = 4;
    const size_t smemB = sizeof(double2) * (2 * ST * BK * (BM + 2)) + sizeof(double) * ST * BK;
    const TileOrder o = makeOrder(n, BM, ctx->dmColFirst, ctx->dmColLast);
    const long long nTiles = (long long)o.nT * (o.nT + 1) / 2;
    const int grid = (int)std::min<long long>(nTiles, 2LL * ctx->smCount);
    auto kern = k3_dm_cplx<BM, BK, ST, 2, 2>;
    DFTBP_TRY(setSmem(ctx, kern, smemB));
    kern<<<grid, 128, smemB, ctx->stream>>>((const double2 *)dC, ld, dW, n, K, (double2 *)dDm, ld, o, counter);
  }
  ctx->launches++;
  }
  return finish(ctx);
}

}  // namespace

extern "C" {

int dftbp_b200_density_matrix_real(dftbp_b200_ctx *ctx, double *dm, const double *eigvecs, int n, int ld,
                                   const double *filling, const double *weights) {
  return densityImpl(ctx, dm, eigvecs, n, ld, filling, weights, false);
}

int dftbp_b200_density_matrix_cplx(dftbp_b200_ctx *ctx, void *dm, const void *eigvecs, int n, int ld,
                                   const double *filling, const double *weights) {
  return densityImpl(ctx, dm, eigvecs, n, ld, filling, weights, true);
}

int dftbp_b200_density_matrix_real_cols(dftbp_b200_ctx *ctx, double *dm, const double *eigvecs, int n, int ld,
                                        const double *filling, const double *weights, int colFirst, int colLast) {
  return densityImpl(ctx, dm, eigvecs, n, ld, filling, weights, false, colFirst, colLast);
}

int dftbp_b200_density_matrix_cplx_cols(dftbp_b200_ctx *ctx, void *dm, const void *eigvecs, int n, int ld,
                                        const double *filling, const double *weights, int colFirst, int colLast) {
  return densityImpl(ctx, dm, eigvecs, n, ld, filling, weights, true, colFirst, colLast);
}

}  // extern "C"

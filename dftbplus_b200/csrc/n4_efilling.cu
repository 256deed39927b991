// n4_efilling.cu -- Efilling (dftb/etemp.F90:59-191) on the device, Fermi distribution: the Fermi level
// for a given number of electrons and the fillings, band energy, electronic entropy and the energy
// extrapolated to 0 K (first piece).  In an SCC iteration it sits between the
// eigensolve and the density matrix, both on the device: with the eigenvalues and fillings staying
// there, the iteration needs no host round trip for them (caller: getFillingsAndBandEnergies,
// dftbplus/main.F90:4597-4648, which also applies the factor two of spin-unpolarised runs).
//
// One CTA.  Every pass of the reference over the eigenvalues (electronCount, etemp.F90:195-290) is a
// block-wide sum here: each thread adds its strided share with the reference's compensated summation
// (math/summation.F90:20-42), the partial sums are added in a fixed tree -- the result can differ from
// the serial sum in the last bits, which moves the Fermi level by an ulp at most; fillings agree with
// the serial code to ~1e-13 at kT >= 1e-3 and exactly for gapped spectra at the default (0 K -> minTemp).
// Control flow as in the reference: no electrons / all levels filled; the mid-gap shortcut for integer
// electron numbers (middleGap, etemp.F90:493-525 -- without its sort: the level at which the aufbau
// count reaches the electron number is found by bisection on the eigenvalue, the next level by a
// minimum search); bracket widening; bisection until the count is accurate to elecTol or the interval
// has collapsed; electronFill (etemp.F90:299-466) with the cut at elecTol; the final rescaling.
#include <cfloat>

#include "common.cuh"

using namespace dftbp_b200;

namespace {

constexpr int kThreads = 1024;
constexpr double kElecTol = 1.0e-15;     // common/accuracy.F90:55
constexpr double kElecTolMax = 1.0e-7;   // common/accuracy.F90:60
constexpr double kEpsSp = 1.1920928955078125e-07;  // epsilon(1.0_rsp)

struct FillParams {
  const double *eig;      // (nLev, nK, nSpin)
  const double *kWeight;  // (nK)
  int nLev, nK, nSpin;
  double nElectrons, kT;
  double *filling, *Ebs, *Ef, *TS, *E0;  // filling (nLev, nK, nSpin); Ebs, TS, E0 (nSpin); Ef scalar
  int *status;            // 0 ok, 1 "Fermi level search did not converge."
};

// math/summation.F90:20-42
__device__ __forceinline__ void kahan(double &sum, double input, double &comp) {
  const double tmp = sum + input;
  if (fabs(sum) >= fabs(input)) comp = comp + (sum - tmp) + input;
  else comp = comp + (input - tmp) + sum;
  sum = tmp;
}

// block-wide reductions (all threads get the result); red: kThreads / 32 doubles of shared memory
template <class Op>
__device__ __forceinline__ double blockReduce(double v, double *red, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}
struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

__device__ __forceinline__ double fermi(double e, double ef, double beta) {
  return 1.0 / (1.0 + exp((e - ef) * beta));  // exp -> inf gives 0, as in the reference without EXP_TRAP
}

__global__ void __launch_bounds__(kThreads) n4_efilling(const FillParams p) {
  __shared__ double red[kThreads / 32];
  const int tid = threadIdx.x;
  const long long nPerSpin = (long long)p.nLev * p.nK, n = nPerSpin * p.nSpin;
  const double beta = 1.0 / p.kT, nEl = p.nElectrons;
  const double eps = DBL_EPSILON, epsilon2 = 2.0 * DBL_EPSILON;

  auto weightOf = [&](long long i) { return __ldg(p.kWeight + (i / p.nLev) % p.nK); };
  // electronCount(Ef)
  auto count = [&](double ef) {
    double s = 0.0, c = 0.0;
    for (long long i = tid; i < n; i += kThreads) kahan(s, weightOf(i) * fermi(__ldg(p.eig + i), ef, beta), c);
    return blockReduce(s + c, red, OpAdd());
  };
  // sum of the weights of the levels with eigenvalue <= e (aufbau count)
  auto countBelow = [&](double e) {
    double s = 0.0;
    for (long long i = tid; i < n; i += kThreads) s += __ldg(p.eig + i) <= e ? weightOf(i) : 0.0;
    return blockReduce(s, red, OpAdd());
  };
  // electronFill(Ef) + optional rescaling
  auto fill = [&](double ef, double scale) {
    for (int iS = 0; iS < p.nSpin; ++iS) {
      double eb = 0.0, cb = 0.0, ts = 0.0, ct = 0.0;
      for (long long i = tid; i < nPerSpin; i += kThreads) {
        const long long g = i + nPerSpin * iS;
        const int lev = (int)(i % p.nLev);
        const double e = __ldg(p.eig + g), w = weightOf(g);
        double f = fermi(e, ef, beta);
        // the level loop of the reference stops at the first filling <= elecTol: that value stays,
        // the levels above keep zero (eigenvalues ascend within a (k, spin) column)
        const bool prevCut = lev > 0 && fermi(__ldg(p.eig + g - 1), ef, beta) <= kElecTol;
        const bool counted = !prevCut && f > kElecTol;
        if (prevCut) f = 0.0;
        if (counted) {
          if (f > eps && f < 1.0 - eps) kahan(ts, -w * (f * log(f) + (1.0 - f) * log(1.0 - f)), ct);
          kahan(eb, w * (f * e), cb);
        }
        p.filling[g] = f * scale;
      }
      const double ebT = blockReduce(eb + cb, red, OpAdd()), tsT = blockReduce(ts + ct, red, OpAdd()) * p.kT;
      if (tid == 0) {
        p.Ebs[iS] = ebT;
        p.TS[iS] = tsT;
        p.E0[iS] = ebT - 0.5 * tsT;
      }
    }
  };

  if (tid == 0) *p.status = 0;
  double vmin = DBL_MAX, vmax = -DBL_MAX;
  for (long long i = tid; i < n; i += kThreads) {
    const double e = __ldg(p.eig + i);
    vmin = fmin(vmin, e);
    vmax = fmax(vmax, e);
  }
  const double eMinAll = blockReduce(vmin, red, OpMin()), eMaxAll = blockReduce(vmax, red, OpMax());

  // no electrons: Fermi level well below the lowest eigenvalue, everything zero
  if (nEl < eps) {
    for (long long i = tid; i < n; i += kThreads) p.filling[i] = 0.0;
    if (tid < p.nSpin) p.Ebs[tid] = p.TS[tid] = p.E0[tid] = 0.0;
    if (tid == 0) *p.Ef = eMinAll - 1000.0 * (p.kT + kEpsSp);
    return;
  }
  // all levels filled: Fermi level well above the highest eigenvalue
  if ((double)p.nLev * p.nSpin <= nEl) {
    const double ef = eMaxAll + 1000.0 * (p.kT + kEpsSp);
    fill(ef, 1.0);
    if (tid == 0) *p.Ef = ef;
    return;
  }
  // integer number of electrons: try the middle of the gap
  if (fabs(nEl - rint(nEl)) <= kElecTol) {
    // e_i = the level at which the aufbau count reaches nEl: bisection on the value
    double lo = eMinAll - 1.0, hi = eMaxAll;  // count(lo) = 0 < nEl <= count(hi)
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (!(mid > lo && mid < hi)) break;
      if (countBelow(mid) >= nEl) hi = mid;
      else lo = mid;
    }
    double ei = -DBL_MAX, ej = DBL_MAX, tiesW = 0.0, tieMax = 0.0;
    for (long long i = tid; i < n; i += kThreads) {
      const double e = __ldg(p.eig + i);
      if (e <= hi) ei = fmax(ei, e);
    }
    ei = blockReduce(ei, red, OpMax());
    for (long long i = tid; i < n; i += kThreads) {
      const double e = __ldg(p.eig + i);
      if (e > ei) ej = fmin(ej, e);
      if (e == ei) {
        tiesW += weightOf(i);
        tieMax = fmax(tieMax, weightOf(i));
      }
    }
    ej = blockReduce(ej, red, OpMin());
    tiesW = blockReduce(tiesW, red, OpAdd());
    tieMax = blockReduce(tieMax, red, OpMax());
    // levels equal to e_i left over once the count has reached nEl: the next level is e_i itself
    const double wEi = countBelow(ei);
    if (wEi - nEl >= tieMax * (1.0 - 1.0e-12) && tiesW > tieMax * (1.0 + 1.0e-12)) ej = ei;
    if (ej == DBL_MAX) ej = ei;  // "just in case the system has all levels filled"
    const double ef = 0.5 * (ej + ei);
    const double ne = count(ef);
    if (fabs(nEl - ne) <= kElecTolMax) {
      fill(ef, 1.0);
      if (tid == 0) *p.Ef = ef;
      return;
    }
  }
  // bracket: lowest eigenvalue of the first level, highest of the last level
  double e1 = DBL_MAX, eN = -DBL_MAX;
  for (long long c = tid; c < (long long)p.nK * p.nSpin; c += kThreads) {
    e1 = fmin(e1, __ldg(p.eig + c * p.nLev));
    eN = fmax(eN, __ldg(p.eig + c * p.nLev + p.nLev - 1));
  }
  double upper = blockReduce(eN, red, OpMax()) + 0.01, lower = blockReduce(e1, red, OpMin()) - 0.01;
  double nMax = count(upper), nMin = count(lower);
  for (int guard = 0; nMin > nEl && guard < 200; ++guard) {
    lower = 2.0 * (lower - upper) + lower;
    nMin = count(lower);
  }
  for (int guard = 0; nMax < nEl && guard < 200; ++guard) {
    upper = 2.0 * (upper - lower) + lower;
    nMax = count(upper);
  }
  double ef = 0.5 * (upper + lower);
  double ne = count(ef);
  while (fabs(nEl - ne) > kElecTol && fabs(upper - lower) >= fmax(fabs(ef) * epsilon2, epsilon2)) {
    if ((nMax >= nMin) == (nEl >= ne)) {
      lower = ef;
      nMin = ne;
    } else {
      upper = ef;
      nMax = ne;
    }
    ef = 0.5 * (upper + lower);
    ne = count(ef);
  }
  if (fabs(nEl - ne) > kElecTolMax) {
    if (tid == 0) *p.status = 1;
    return;
  }
  const double scale = (nEl > epsilon2 && fabs(nEl - ne) > epsilon2) ? nEl / ne : 1.0;
  fill(ef, scale);
  if (tid == 0) *p.Ef = ef;
}

}  // namespace

extern "C" {

int dftbp_b200_efilling(dftbp_b200_ctx *ctx, double *Ebs, double *Ef, double *TS, double *E0, double *filling,
                        const double *eigenvals, int nLevels, int nKpt, int nSpin, double nElectrons, double kT,
                        const double *kWeight, int distrib) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, Ebs && Ef && TS && E0 && filling && eigenvals && kWeight, "efilling: null argument");
  DFTBP_CHECK(ctx, nLevels > 0 && nKpt > 0 && (nSpin == 1 || nSpin == 2), "efilling: bad sizes");
  DFTBP_CHECK(ctx, kT > 0.0 && nElectrons >= 0.0, "efilling: kT must be positive, nElectrons non-negative");
  if (distrib != DFTBP_B200_FILLING_FERMI)
    return fail(ctx, DFTBP_B200_ENOTIMPL, "efilling: only the Fermi distribution is built (Gaussian / Methfessel-Paxton: host)");
  const size_t n = (size_t)nLevels * nKpt * nSpin;
  FillParams p;
  p.nLev = nLevels; p.nK = nKpt; p.nSpin = nSpin; p.nElectrons = nElectrons; p.kT = kT;
  DFTBP_TRY(devIn(ctx, kSlotA, eigenvals, n, &p.eig));
  DFTBP_TRY(devIn(ctx, kSlotB, kWeight, (size_t)nKpt, &p.kWeight));
  DFTBP_TRY(devOut(ctx, kSlotC, filling, n, false, &p.filling));
  // the four small outputs share one staging block when they are host pointers
  double *small = nullptr;
  const bool smallOnDevice = isDevicePtr(Ebs) && isDevicePtr(Ef) && isDevicePtr(TS) && isDevicePtr(E0);
  if (smallOnDevice) {
    p.Ebs = Ebs; p.Ef = Ef; p.TS = TS; p.E0 = E0;
  } else {
    void *blk;
    DFTBP_TRY(stagingGet(ctx, kSlotD, sizeof(double) * 8, &blk));
    small = static_cast<double *>(blk);
    p.Ebs = small; p.TS = small + 2; p.E0 = small + 4; p.Ef = small + 6;
  }
  p.status = ctx->dInfo + 2;
  n4_efilling<<<1, kThreads, 0, ctx->stream>>>(p);
  ctx->launches++;
  if (!smallOnDevice) {
    double h[8];
    DFTBP_CUDA(ctx, cudaMemcpyAsync(h, small, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    int st = 0;
    DFTBP_CUDA(ctx, cudaMemcpyAsync(&st, p.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (st != 0) {
      ctx->pending.clear();
      return fail(ctx, DFTBP_B200_EINVAL, "Fermi level search did not converge.");
    }
    for (int i = 0; i < nSpin; ++i) {
      auto put = [&](double *dst, double v) {
        if (isDevicePtr(dst)) cudaMemcpyAsync(dst + i, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        else dst[i] = v;
      };
      put(Ebs, h[i]); put(TS, h[2 + i]); put(E0, h[4 + i]);
    }
    if (isDevicePtr(Ef)) DFTBP_CUDA(ctx, cudaMemcpyAsync(Ef, &h[6], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else *Ef = h[6];
  }
  return finish(ctx);
}

}  // extern "C"

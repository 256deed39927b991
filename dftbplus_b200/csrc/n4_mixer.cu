// n4_mixer.cu -- the modified Broyden mixer of the SCC loop with its vectors on the device:
//   TBroydenMixerReal_init   mixer/broydenmixer.F90:110-139
//   TBroydenMixerReal_reset  mixer/broydenmixer.F90:143-170
//   TBroydenMixerReal_mix -> modifiedBroydenMixing   mixer/broydenmixer.F90:180-342
// (D. D. Johnson, PRB 38, 12807 (1988); parser defaults mixParam 0.2, omega0 0.01, minWeight 1, maxWeight 1e5,
// weightFac 1e-2: dftbplus/parser.F90:5786-5790).
// The charge vectors (qInp, qDiff, the history |DF(i)>, |u(i)>) stay on the device between the Mulliken kernel
// and the shift kernels of the next iteration; what crosses to the host per call is the handful of scalar
// products the reference forms with dot_product (2 + 2 (iter - 2) + 1 doubles), and back the iter - 1
// coefficients w_i gamma_i: the (iter - 1)^2 system beta gamma = c is solved on the host (the reference calls
// LAPACK getrf / getrs there).  Element-wise operations are in the reference's order; the scalar products are
// fixed-order tree sums of one CTA each (deterministic; they differ from a serial dot_product by rounding).
#include <cmath>
#include <vector>

#include "common.cuh"

using namespace dftbp_b200;

namespace {

struct MixerState {
  int mIter = 0, nElem = 0, iIter = 0;
  double alpha = 0.2, omega0 = 0.01, minWeight = 1.0, maxWeight = 1.0e5, weightFac = 1.0e-2;
  std::vector<double> ww, aa;  // ww(mIter - 1), aa(mIter - 1, mIter - 1) column-major
  double *qInpLast = nullptr, *qDiffLast = nullptr, *dF = nullptr, *uu = nullptr;  // device: (n), (n), (n, m), (n, m)
  double *dots = nullptr, *coef = nullptr;                                        // device scratch: (2 m + 2), (m)
  size_t capQ = 0, capD = 0, capDF = 0, capU = 0, capDots = 0, capCoef = 0;
  bool initialised = false;
};

MixerState *mixerOf(dftbp_b200_ctx *ctx) {
  if (!ctx->mixerState) ctx->mixerState = new MixerState();
  return static_cast<MixerState *>(ctx->mixerState);
}

// fixed-order sum over the CTA (256 threads): out = sum_j f(j)
template <class F>
__device__ __forceinline__ double ctaSum(int n, F &&f) {
  __shared__ double part[256];
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) s += f(j);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  const double r = part[0];
  __syncthreads();
  return r;
}

// first iteration: store qInp / qDiff, simple mixing (broydenmixer.F90:268-273)
__global__ void __launch_bounds__(256) mix_first(int n, double alpha, double *__restrict__ q, const double *__restrict__ dq,
                                                 double *__restrict__ qLast, double *__restrict__ dLast) {
  for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
    qLast[j] = q[j];
    dLast[j] = dq[j];
    q[j] = q[j] + alpha * dq[j];
  }
}

// dots[0] = <qDiff|qDiff>, dots[1] = <qDiff - qDiffLast|qDiff - qDiffLast>
__global__ void __launch_bounds__(256) mix_norms(int n, const double *__restrict__ dq, const double *__restrict__ dLast,
                                                 double *__restrict__ dots) {
  const bool second = blockIdx.x == 1;
  const double r = ctaSum(n, [&](int j) {
    const double v = second ? dq[j] - dLast[j] : dq[j];
    return v * v;
  });
  if (threadIdx.x == 0) dots[blockIdx.x] = r;
}

// |DF(m-1)> = invNorm (qDiff - qDiffLast) -> dF(:, m1); CTA i < m1: a(i) = <DF(i)|DF(m-1)>, c(i) = <DF(i)|qDiff>;
// CTA m1: c(m1) = <DF(m-1)|qDiff> and the store of the new column
__global__ void __launch_bounds__(256) mix_dots(int n, int m1, double invNorm, const double *__restrict__ dq,
                                                const double *__restrict__ dLast, double *__restrict__ dF,
                                                double *__restrict__ dots) {
  const int i = blockIdx.x;
  if (i < m1) {
    const double *__restrict__ col = dF + (size_t)n * i;
    const double a = ctaSum(n, [&](int j) { return col[j] * (invNorm * (dq[j] - dLast[j])); });
    const double c = ctaSum(n, [&](int j) { return col[j] * dq[j]; });
    if (threadIdx.x == 0) {
      dots[2 + i] = a;
      dots[2 + m1 + 1 + i] = c;
    }
  } else {
    double *__restrict__ col = dF + (size_t)n * m1;
    const double c = ctaSum(n, [&](int j) {
      const double v = invNorm * (dq[j] - dLast[j]);
      col[j] = v;
      return v * dq[j];
    });
    if (threadIdx.x == 0) dots[2 + m1] = c;
  }
}

// |u(m-1)>, the stores of the last vectors and the new input (broydenmixer.F90:318-338), coef(i) = w(i) gamma(i)
__global__ void __launch_bounds__(256) mix_update(int n, int m1, double alpha, double invNorm, const double *__restrict__ coef,
                                                  double *__restrict__ q, const double *__restrict__ dq,
                                                  double *__restrict__ qLast, double *__restrict__ dLast,
                                                  const double *__restrict__ dF, double *__restrict__ uu) {
  for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
    const double u = alpha * dF[(size_t)n * m1 + j] + invNorm * (q[j] - qLast[j]);
    qLast[j] = q[j];
    dLast[j] = dq[j];
    double acc = q[j] + alpha * dq[j];
    for (int i = 0; i < m1; ++i) acc = acc - coef[i] * uu[(size_t)n * i + j];
    acc = acc - coef[m1] * u;
    uu[(size_t)n * m1 + j] = u;
    q[j] = acc;
  }
}

template <class T>
int grow(dftbp_b200_ctx *ctx, T **p, size_t *cap, size_t need) {
  if (*cap >= need && *p) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (need ? need : 1));
  if (e != cudaSuccess) return fail(ctx, DFTBP_B200_ENOMEM, "mixer: cudaMalloc failed: %s", cudaGetErrorString(e));
  *cap = need;
  return 0;
}

// solve A x = b (n x n, column-major) by Gaussian elimination with partial pivoting; false if singular
bool solveDense(int n, std::vector<double> &A, std::vector<double> &b) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i + (size_t)n * k]) > std::fabs(A[p + (size_t)n * k])) p = i;
    if (A[p + (size_t)n * k] == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k + (size_t)n * j], A[p + (size_t)n * j]);
      std::swap(b[k], b[p]);
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i + (size_t)n * k] / A[k + (size_t)n * k];
      if (f == 0.0) continue;
      for (int j = k + 1; j < n; ++j) A[i + (size_t)n * j] -= f * A[k + (size_t)n * j];
      b[i] -= f * b[k];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[k + (size_t)n * j] * b[j];
    b[k] = s / A[k + (size_t)n * k];
  }
  return true;
}

}  // namespace

namespace dftbp_b200 {
void freeMixer(dftbp_b200_ctx *ctx) {
  MixerState *st = static_cast<MixerState *>(ctx->mixerState);
  if (!st) return;
  cudaFree(st->qInpLast); cudaFree(st->qDiffLast); cudaFree(st->dF); cudaFree(st->uu);
  cudaFree(st->dots); cudaFree(st->coef);
  delete st;
  ctx->mixerState = nullptr;
}
}  // namespace dftbp_b200

extern "C" {

int dftbp_b200_mixer_init(dftbp_b200_ctx *ctx, int maxSccIter, double mixParam, double omega0, double minWeight,
                          double maxWeight, double weightFac) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, maxSccIter > 0 && mixParam > 0.0 && omega0 > 0.0, "mixer_init: maxSccIter, mixParam and omega0 must be positive");
  MixerState *st = mixerOf(ctx);
  st->mIter = maxSccIter; st->alpha = mixParam; st->omega0 = omega0;
  st->minWeight = minWeight; st->maxWeight = maxWeight; st->weightFac = weightFac;
  st->nElem = 0; st->iIter = 0;
  st->initialised = true;
  return 0;
}

int dftbp_b200_mixer_reset(dftbp_b200_ctx *ctx, int nElem) {
  if (!ctx) return DFTBP_B200_EINVAL;
  MixerState *st = mixerOf(ctx);
  DFTBP_CHECK(ctx, st->initialised, "mixer_reset: call mixer_init first");
  DFTBP_CHECK(ctx, nElem > 0, "mixer_reset: nElem must be positive");
  const size_t m = (size_t)std::max(1, st->mIter - 1);
  st->nElem = nElem;
  st->iIter = 0;
  st->ww.assign(m, 0.0);
  st->aa.assign(m * m, 0.0);
  DFTBP_TRY(grow(ctx, &st->qInpLast, &st->capQ, (size_t)nElem));
  DFTBP_TRY(grow(ctx, &st->qDiffLast, &st->capD, (size_t)nElem));
  DFTBP_TRY(grow(ctx, &st->dF, &st->capDF, (size_t)nElem * m));
  DFTBP_TRY(grow(ctx, &st->uu, &st->capU, (size_t)nElem * m));
  DFTBP_TRY(grow(ctx, &st->dots, &st->capDots, 2 * m + 2));
  DFTBP_TRY(grow(ctx, &st->coef, &st->capCoef, m));
  return 0;
}

int dftbp_b200_mixer_mix(dftbp_b200_ctx *ctx, double *qInpResult, const double *qDiff, int nElem) {
  if (!ctx) return DFTBP_B200_EINVAL;
  MixerState *st = mixerOf(ctx);
  DFTBP_CHECK(ctx, st->initialised && st->nElem > 0, "mixer_mix: call mixer_init and mixer_reset first");
  DFTBP_CHECK(ctx, qInpResult && qDiff && nElem == st->nElem, "mixer_mix: vectors of the length given to mixer_reset expected");
  st->iIter++;
  if (st->iIter > st->mIter) return fail(ctx, DFTBP_B200_EINVAL, "Broyden mixer: Maximal nr. of steps exceeded");
  const int n = nElem, nn = st->iIter;
  double *dQ;
  const double *dD;
  DFTBP_TRY(devOut(ctx, kSlotA, qInpResult, (size_t)n, true, &dQ));
  DFTBP_TRY(devIn(ctx, kSlotB, qDiff, (size_t)n, &dD));
  const int grid = std::max(1, std::min((n + 255) / 256, ctx->smCount * 4));
  if (nn == 1) {
    mix_first<<<grid, 256, 0, ctx->stream>>>(n, st->alpha, dQ, dD, st->qInpLast, st->qDiffLast);
    ctx->launches++;
    return finish(ctx);
  }
  const int m1 = nn - 2;  // 0-based position of iteration nn - 1
  const int mm = st->mIter - 1;
  double h2[2];
  mix_norms<<<2, 256, 0, ctx->stream>>>(n, dD, st->qDiffLast, st->dots);
  DFTBP_CUDA(ctx, cudaMemcpyAsync(h2, st->dots, sizeof(h2), cudaMemcpyDeviceToHost, ctx->stream));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // weight of the current iteration (broydenmixer.F90:278-287)
  double w = std::sqrt(h2[0]);
  w = w > st->weightFac / st->maxWeight ? st->weightFac / w : st->maxWeight;
  if (w < st->minWeight) w = st->minWeight;
  st->ww[m1] = w;
  // 1 / max(|qDiff - qDiffLast|, epsilon)
  double invNorm = std::sqrt(h2[1]);
  invNorm = 1.0 / std::max(invNorm, 2.220446049250313e-16);
  mix_dots<<<m1 + 1, 256, 0, ctx->stream>>>(n, m1, invNorm, dD, st->qDiffLast, st->dF, st->dots);
  std::vector<double> hd(2 * (size_t)m1 + 1);
  DFTBP_CUDA(ctx, cudaMemcpyAsync(hd.data(), st->dots + 2, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost, ctx->stream));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<double> cc((size_t)m1 + 1);
  for (int i = 0; i < m1; ++i) {
    st->aa[i + (size_t)mm * m1] = hd[i];
    st->aa[m1 + (size_t)mm * i] = hd[i];
    cc[i] = st->ww[i] * hd[m1 + 1 + i];
  }
  st->aa[m1 + (size_t)mm * m1] = 1.0;
  cc[m1] = st->ww[m1] * hd[m1];
  const int k = m1 + 1;
  std::vector<double> beta((size_t)k * k);
  for (int j = 0; j < k; ++j) {
    for (int i = 0; i < k; ++i) beta[i + (size_t)k * j] = st->ww[i] * st->ww[j] * st->aa[i + (size_t)mm * j];
    beta[j + (size_t)k * j] += st->omega0 * st->omega0;
  }
  // gamma = c beta^-1 (beta is symmetric: getrs with trans = 't' in the reference)
  if (!solveDense(k, beta, cc)) return fail(ctx, DFTBP_B200_EINVAL, "Broyden mixer: singular system");
  for (int i = 0; i < k; ++i) cc[i] = st->ww[i] * cc[i];
  DFTBP_CUDA(ctx, cudaMemcpyAsync(st->coef, cc.data(), sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
  mix_update<<<grid, 256, 0, ctx->stream>>>(n, m1, st->alpha, invNorm, st->coef, dQ, dD, st->qInpLast, st->qDiffLast,
                                            st->dF, st->uu);
  ctx->launches += 3;
  // cc is a host temporary staged asynchronously: the copy has been consumed once the stream is idle
  DFTBP_TRY(finish(ctx));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

// solver.cu -- the dense generalized eigensolve between unpackHS and the density matrix.
// NOT rewritten (BASELINE.json north_star): a thin wrapper over cuSOLVER's divide-and-conquer
// xSYGVD / xHEGVD, the counterpart of diagDenseMtx -> hegvd (dftbplus/eigenvects.F90:85-100,
// math/eigensolver.F90).  Kept here so that H, S and the eigenvectors stay device-resident.
#include <cusolverDn.h>

#include "common.cuh"

using namespace dftbp_b200;

namespace {

int ensureHandle(dftbp_b200_ctx *ctx, cusolverDnHandle_t *h) {
  if (!ctx->solverHandle) {
    cusolverDnHandle_t hh;
    if (cusolverDnCreate(&hh) != CUSOLVER_STATUS_SUCCESS)
      return fail(ctx, DFTBP_B200_ECUDA, "cusolverDnCreate failed");
    ctx->solverHandle = hh;
  }
  *h = (cusolverDnHandle_t)ctx->solverHandle;
  if (cusolverDnSetStream(*h, ctx->stream) != CUSOLVER_STATUS_SUCCESS)
    return fail(ctx, DFTBP_B200_ECUDA, "cusolverDnSetStream failed");
  return 0;
}

int ensureWork(dftbp_b200_ctx *ctx, size_t bytes) {
  if (ctx->solverWorkBytes >= bytes) return 0;
  cudaFree(ctx->solverWork);
  ctx->solverWork = nullptr;
  ctx->solverWorkBytes = 0;
  cudaError_t e = cudaMalloc(&ctx->solverWork, bytes);
  if (e != cudaSuccess)
    return fail(ctx, DFTBP_B200_ENOMEM, "cudaMalloc of %zu solver workspace bytes failed: %s", bytes,
                cudaGetErrorString(e));
  ctx->solverWorkBytes = bytes;
  return 0;
}

int hegvImpl(dftbp_b200_ctx *ctx, void *H, void *S, double *eigen, int n, int ld, char jobz, bool cplx) {
  if (!ctx) return DFTBP_B200_EINVAL;
  DFTBP_CHECK(ctx, H && S && eigen && n > 0 && ld >= n, "hegv: bad arguments");
  DFTBP_CHECK(ctx, jobz == 'V' || jobz == 'N' || jobz == 'v' || jobz == 'n', "hegv: jobz must be 'V' or 'N'");
  cusolverDnHandle_t h;
  DFTBP_TRY(ensureHandle(ctx, &h));
  const size_t el = cplx ? 2 : 1, nSq = (size_t)ld * n * el;
  double *dH, *dS, *dE;
  DFTBP_TRY(devOut(ctx, kSlotA, (double *)H, nSq, true, &dH));
  DFTBP_TRY(devOut(ctx, kSlotB, (double *)S, nSq, true, &dS));
  DFTBP_TRY(devOut(ctx, kSlotC, eigen, (size_t)n, false, &dE));
  const cusolverEigMode_t mode = (jobz == 'V' || jobz == 'v') ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR;
  int lwork = 0;
  cusolverStatus_t st;
  if (cplx)
    st = cusolverDnZhegvd_bufferSize(h, CUSOLVER_EIG_TYPE_1, mode, CUBLAS_FILL_MODE_LOWER, n,
                                     (cuDoubleComplex *)dH, ld, (cuDoubleComplex *)dS, ld, dE, &lwork);
  else
    st = cusolverDnDsygvd_bufferSize(h, CUSOLVER_EIG_TYPE_1, mode, CUBLAS_FILL_MODE_LOWER, n, dH, ld, dS,
                                     ld, dE, &lwork);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    ctx->pending.clear();
    return fail(ctx, DFTBP_B200_ECUDA, "cusolver hegvd bufferSize failed (%d)", (int)st);
  }
  DFTBP_TRY(ensureWork(ctx, (size_t)lwork * sizeof(double) * el));
  {
  ScopedTimer timer(ctx, kTimerDiagonalization);
  if (cplx)
    st = cusolverDnZhegvd(h, CUSOLVER_EIG_TYPE_1, mode, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex *)dH, ld,
                          (cuDoubleComplex *)dS, ld, dE, (cuDoubleComplex *)ctx->solverWork, lwork, ctx->dInfo);
  else
    st = cusolverDnDsygvd(h, CUSOLVER_EIG_TYPE_1, mode, CUBLAS_FILL_MODE_LOWER, n, dH, ld, dS, ld, dE,
                          (double *)ctx->solverWork, lwork, ctx->dInfo);
  }
  if (st != CUSOLVER_STATUS_SUCCESS) {
    ctx->pending.clear();
    return fail(ctx, DFTBP_B200_ECUDA, "cusolver hegvd failed (%d)", (int)st);
  }
  int info = 0;
  DFTBP_CUDA(ctx, cudaMemcpyAsync(&info, ctx->dInfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DFTBP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (info != 0) {
    ctx->pending.clear();
    // same wording class as the reference's LAPACK info handling (math/eigensolver.F90)
    return fail(ctx, DFTBP_B200_ECUDA, "hegvd: info = %d (%s)", info,
                info > n ? "overlap matrix not positive definite" : "eigensolver did not converge");
  }
  return finish(ctx);
}

}  // namespace

namespace dftbp_b200 {
void solverDestroy(dftbp_b200_ctx *ctx) {
  if (ctx->solverHandle) cusolverDnDestroy((cusolverDnHandle_t)ctx->solverHandle);
  ctx->solverHandle = nullptr;
}
}  // namespace dftbp_b200

extern "C" {

int dftbp_b200_hegv_real(dftbp_b200_ctx *ctx, double *H, double *S, double *eigen, int n, int ld, char jobz) {
  return hegvImpl(ctx, H, S, eigen, n, ld, jobz, false);
}

int dftbp_b200_hegv_cplx(dftbp_b200_ctx *ctx, void *H, void *S, double *eigen, int n, int ld, char jobz) {
  return hegvImpl(ctx, H, S, eigen, n, ld, jobz, true);
}

}  // extern "C"

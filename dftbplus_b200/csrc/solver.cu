// solver.cu -- the dense generalized eigensolve between unpackHS and the density matrix.
// NOT rewritten (BASELINE.json north_star): a thin wrapper over cuSOLVER's divide-and-conquer
// xSYGVD / xHEGVD, the counterpart of diagDenseMtx -> hegvd (dftbplus/eigenvects.F90:85-100,
// math/eigensolver.F90).  Kept here so that H, S and the eigenvectors stay device-resident.
#include <cublas_v2.h>
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

// upper triangle <- conjugate transpose of the lower one (32 x 32 tiles through shared memory)
template <bool CPLX>
__global__ void __launch_bounds__(256) sv_fill_upper(double *a, long long n, long long ld) {
  __shared__ double tile[CPLX ? 2 : 1][32][33];
  const long long bi = blockIdx.y, bj = blockIdx.x;  // tile row / column of the source (lower triangle: bi >= bj)
  if (bi < bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long i = bi * 32 + tx, j = bj * 32 + r;  // source element (i, j), column-major
    if (i < n && j < n) {
      if (CPLX) {
        tile[0][r][tx] = a[2 * (i + ld * j)];
        tile[CPLX ? 1 : 0][r][tx] = a[2 * (i + ld * j) + 1];
      } else {
        tile[0][r][tx] = a[i + ld * j];
      }
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long i = bi * 32 + r, j = bj * 32 + tx;  // source (i, j) -> destination (j, i)
    if (i < n && j < n && i > j) {
      if (CPLX) {
        a[2 * (j + ld * i)] = tile[0][tx][r];
        a[2 * (j + ld * i) + 1] = -tile[CPLX ? 1 : 0][tx][r];
      } else {
        a[j + ld * i] = tile[0][tx][r];
      }
    }
  }
}

// The same reduction LAPACK's xSYGVD / xHEGVD performs, from cuSOLVER's 64-bit building blocks: the
// 32-bit lwork of cusolverDn?sygvd overflows at the BASELINE size n = 31744.
//   S = L L^H (Xpotrf);  A = L^-1 H L^-H (two triangular solves on the full matrix);  A y = e y (Xsyevd);
//   c = L^-H y (triangular solve).  On exit H holds the eigenvectors, S the Cholesky factor, like hegvd.
int hegvLarge(dftbp_b200_ctx *ctx, cusolverDnHandle_t h, double *dH, double *dS, double *dE, int n, int ld,
              cusolverEigMode_t mode, bool cplx) {
  if (!ctx->blasHandle) {
    cublasHandle_t bh;
    if (cublasCreate(&bh) != CUBLAS_STATUS_SUCCESS) return fail(ctx, DFTBP_B200_ECUDA, "cublasCreate failed");
    ctx->blasHandle = bh;
  }
  cublasHandle_t bh = (cublasHandle_t)ctx->blasHandle;
  if (cublasSetStream(bh, ctx->stream) != CUBLAS_STATUS_SUCCESS) return fail(ctx, DFTBP_B200_ECUDA, "cublasSetStream failed");
  cusolverDnParams_t params;
  if (cusolverDnCreateParams(&params) != CUSOLVER_STATUS_SUCCESS) return fail(ctx, DFTBP_B200_ECUDA, "cusolverDnCreateParams failed");
  const cudaDataType dt = cplx ? CUDA_C_64F : CUDA_R_64F;
  size_t wDev1 = 0, wHost1 = 0, wDev2 = 0, wHost2 = 0;
  cusolverStatus_t st = cusolverDnXpotrf_bufferSize(h, params, CUBLAS_FILL_MODE_LOWER, n, dt, dS, ld, dt, &wDev1, &wHost1);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    cusolverDnDestroyParams(params);
    return fail(ctx, DFTBP_B200_ECUDA, "cusolverDnXpotrf_bufferSize failed (%d)", (int)st);
  }
  if (st == CUSOLVER_STATUS_SUCCESS)
    st = cusolverDnXsyevd_bufferSize(h, params, mode, CUBLAS_FILL_MODE_LOWER, n, dt, dH, ld, CUDA_R_64F, dE, dt, &wDev2, &wHost2);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    cusolverDnDestroyParams(params);
    return fail(ctx, DFTBP_B200_ECUDA, "cusolverDnXsyevd_bufferSize failed (%d; n %d ld %d mode %d cplx %d, potrf workspace %zu bytes)",
                (int)st, n, ld, (int)mode, (int)cplx, wDev1);
  }
  int rc = ensureWork(ctx, std::max(wDev1, wDev2));
  std::vector<char> hostWork(std::max(wHost1, wHost2) + 1);
  if (rc) {
    cusolverDnDestroyParams(params);
    return rc;
  }
  {
    ScopedTimer timer(ctx, kTimerDiagonalization);
    st = cusolverDnXpotrf(h, params, CUBLAS_FILL_MODE_LOWER, n, dt, dS, ld, dt, ctx->solverWork, wDev1, hostWork.data(), wHost1, ctx->dInfo);
    if (st == CUSOLVER_STATUS_SUCCESS) {
      int info = 0;
      cudaMemcpyAsync(&info, ctx->dInfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      if (info != 0) {
        cusolverDnDestroyParams(params);
        return fail(ctx, DFTBP_B200_ECUDA, "hegvd: info = %d (overlap matrix not positive definite)", n + info);
      }
      const dim3 grid((n + 31) / 32, (n + 31) / 32);
      cublasStatus_t bs;
      if (cplx) {
        sv_fill_upper<true><<<grid, 256, 0, ctx->stream>>>(dH, n, ld);
        const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
        bs = cublasZtrsm_64(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one,
                            (const cuDoubleComplex *)dS, ld, (cuDoubleComplex *)dH, ld);
        if (bs == CUBLAS_STATUS_SUCCESS)
          bs = cublasZtrsm_64(bh, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, n, n, &one,
                              (const cuDoubleComplex *)dS, ld, (cuDoubleComplex *)dH, ld);
      } else {
        sv_fill_upper<false><<<grid, 256, 0, ctx->stream>>>(dH, n, ld);
        const double one = 1.0;
        bs = cublasDtrsm_64(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, dS, ld, dH, ld);
        if (bs == CUBLAS_STATUS_SUCCESS)
          bs = cublasDtrsm_64(bh, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n, n, &one, dS, ld, dH, ld);
      }
      ctx->launches++;
      if (bs != CUBLAS_STATUS_SUCCESS) {
        cusolverDnDestroyParams(params);
        return fail(ctx, DFTBP_B200_ECUDA, "cublas trsm failed (%d)", (int)bs);
      }
      st = cusolverDnXsyevd(h, params, mode, CUBLAS_FILL_MODE_LOWER, n, dt, dH, ld, CUDA_R_64F, dE, dt, ctx->solverWork, wDev2,
                            hostWork.data(), wHost2, ctx->dInfo);
      if (st == CUSOLVER_STATUS_SUCCESS && mode == CUSOLVER_EIG_MODE_VECTOR) {
        if (cplx) {
          const cuDoubleComplex one = make_cuDoubleComplex(1.0, 0.0);
          bs = cublasZtrsm_64(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_C, CUBLAS_DIAG_NON_UNIT, n, n, &one,
                              (const cuDoubleComplex *)dS, ld, (cuDoubleComplex *)dH, ld);
        } else {
          const double one = 1.0;
          bs = cublasDtrsm_64(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, n, n, &one, dS, ld, dH, ld);
        }
        if (bs != CUBLAS_STATUS_SUCCESS) {
          cusolverDnDestroyParams(params);
          return fail(ctx, DFTBP_B200_ECUDA, "cublas trsm failed (%d)", (int)bs);
        }
      }
    }
  }
  cusolverDnDestroyParams(params);
  if (st != CUSOLVER_STATUS_SUCCESS) return fail(ctx, DFTBP_B200_ECUDA, "cusolver X potrf / syevd failed (%d)", (int)st);
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
  // workspace of the 32-bit API: about 2 n^2 elements -> beyond INT_MAX near n = 23 000 (real)
  const bool large = (double)n * n * (cplx ? 2.2 : 2.1) > 2.0e9 || tuneIntNow("DFTBP_B200_HEGV_LARGE", 0);
  if (large) {
    int rc = hegvLarge(ctx, h, dH, dS, dE, n, ld, mode, cplx);
    if (rc) {
      ctx->pending.clear();
      return rc;
    }
  } else {
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
  if (ctx->blasHandle) cublasDestroy((cublasHandle_t)ctx->blasHandle);
  ctx->blasHandle = nullptr;
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

// which n does cusolverDnXsyevd_bufferSize accept? (development aid)
#include <cstdio>
#include <cusolverDn.h>
int main() {
  cusolverDnHandle_t h; cusolverDnCreate(&h);
  cusolverDnParams_t p; cusolverDnCreateParams(&p);
  double *A, *W; cudaMalloc(&A, 1 << 20); cudaMalloc(&W, 1 << 20);
  for (long long n : {4096LL, 16384LL, 23000LL, 23170LL, 23171LL, 24000LL, 31744LL, 32768LL, 46000LL}) {
    size_t d = 0, hh = 0;
    cusolverStatus_t s1 = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &d, &hh);
    size_t d2 = 0, h2 = 0;
    cusolverStatus_t s2 = cusolverDnXsyevd_bufferSize(h, p, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &d2, &h2);
    int lw = 0;
    cusolverStatus_t s3 = cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A, (int)n, W, &lw);
    int lw4 = 0;
    cusolverStatus_t s4 = cusolverDnDsygvd_bufferSize(h, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A, (int)n, A, (int)n, W, &lw4);
    printf("n %lld: Xsyevd V status %d dev %zu host %zu | N status %d dev %zu | Dsyevd status %d lwork %d | Dsygvd status %d lwork %d\n", n, (int)s1, d, hh, (int)s2, d2, (int)s3, lw, (int)s4, lw4);
  }
  return 0;
}

// fp64_latency.cu -- DFMA / DMUL dependent-issue latency and throughput per SM on this GPU, LDS latency
// (development aid; nvcc -gencode arch=compute_100a,code=sm_100a -O3)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, int iters, int ilp, long long *cycles) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c);
    if (ilp > 1) a1 = fma(a1, b, c);
    if (ilp > 2) { a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    if (ilp > 4) { a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
__global__ void ldsChain(int *out, int iters, long long *cycles) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) j = s[j];
  long long t1 = clock64();
  out[threadIdx.x] = j;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  double *out; long long *cyc; int *iout;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&iout, 4096); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  for (int ilp : {1, 2, 4, 8})
    for (int warps : {1, 2, 4, 8, 16, 32}) {
      chain<<<1, 32 * warps>>>(out, iters, ilp, cyc);
      cudaDeviceSynchronize();
      printf("DFMA ilp %d warps/SM %2d: %6.2f cycles per dependent step, %6.2f warp-DFMA per cycle per SM\n", ilp, warps,
             (double)*cyc / iters, (double)ilp * warps * iters / (double)*cyc);
    }
  ldsChain<<<1, 32>>>(iout, iters, cyc);
  cudaDeviceSynchronize();
  printf("LDS dependent latency: %.1f cycles\n", (double)*cyc / iters);
  return 0;
}

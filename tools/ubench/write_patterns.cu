// write_patterns.cu -- micro-benchmark: what HBM write rate do the store patterns of the pair build
// reach on a B200?  (development aid; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3)
//   A  coalesced grid-stride 8-byte stores over one array
//   B  same, 16-byte stores
//   C  runs of 81 doubles at arbitrary 8-byte offsets, one warp per run, runs visited in a shuffled order
//   D  like C, a warp writes 32 runs per "item" from shared memory (the K1 write-out loop)
//   E  like D with runs visited in memory order
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void kA(double *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)i;
}
__global__ void kB(double2 *out, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_double2((double)i, 1.0);
}
__global__ void kC(double *out, const int *start, int nRuns, int len) {
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
  for (int r = w; r < nRuns; r += nW) {
    double *d = out + start[r] + lane;
    for (int e = lane; e < len; e += 32) d[e - lane] = (double)e;
  }
}
template <int PITCH>
__global__ void kD(double *out, const int *start, int nRuns, int len) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nWc = blockDim.x >> 5;
  double *stage = sm + (size_t)wIn * 32 * PITCH;
  for (int e = 0; e < PITCH; ++e) stage[lane * PITCH + e] = (double)(e + lane);
  __syncwarp();
  const int w = blockIdx.x * nWc + wIn, nW = gridDim.x * nWc;
  const bool m0 = lane < len, m1 = lane + 32 < len, m2 = lane + 64 < len;
  for (int it = w; it * 32 < nRuns; it += nW) {
    const int ind = it * 32 + lane < nRuns ? start[it * 32 + lane] : start[it * 32];
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const int base = __shfl_sync(0xffffffffu, ind, q);
      double *d = out + base + lane;
      const double *s = stage + q * PITCH + lane;
      const double v0 = s[0], v1 = s[32], v2 = s[64];
      if (m0) d[0] = v0;
      if (m1) d[32] = v1;
      if (m2) d[64] = v2;
    }
  }
}

//   F  a warp fills a linear image of 32 consecutive runs in shared memory and one lane sends it with
//      cp.async.bulk (TMA 1-D bulk store), 16-byte interior + scalar edges
//   G  same image, copied out by the 32 lanes with coalesced 8-byte stores
template <bool BULK>
__global__ void kF(double *out, int nRuns, int len, int imgCap) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, wIn = threadIdx.x >> 5, nWc = blockDim.x >> 5;
  double *img = sm + (size_t)wIn * imgCap;
  const int w = blockIdx.x * nWc + wIn, nW = gridDim.x * nWc;
  for (int it = w; it * 32 < nRuns; it += nW) {
    const long long first = (long long)it * 32 * len;
    const int nR = min(32, nRuns - it * 32), total = nR * len;
    const int shift = (int)(first & 1);  // smem element j <-> global element first + j, same 16-byte phase
    if (BULK) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (lane < nR)
      for (int e = 0; e < len; ++e) img[shift + lane * len + e] = (double)(e + lane);
    __syncwarp();
    if (BULK) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) {
        const int head = shift, body = (total - head) & ~1, tail = total - head - body;
        if (head) out[first] = img[shift];
        if (body > 0) {
          const unsigned src = (unsigned)__cvta_generic_to_shared(img + shift + head);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + first + head), "r"(src),
                       "r"(body * 8)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tail) out[first + total - 1] = img[shift + total - 1];
      }
    } else {
      for (int e = lane; e < total; e += 32) out[first + e] = img[shift + e];
    }
  }
  if (BULK) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
  const int len = argc > 1 ? atoi(argv[1]) : 81;
  const int nRuns = 2 * 131072;  // H0 + S of ~131k pairs
  const size_t n = (size_t)nRuns * len;
  double *out;
  CK(cudaMalloc(&out, (n + 64) * sizeof(double)));
  std::vector<int> start(nRuns);
  for (int r = 0; r < nRuns; ++r) start[r] = (int)((size_t)r * len);
  std::vector<int> shuffled = start;
  std::mt19937 rng(1);
  std::shuffle(shuffled.begin(), shuffled.end(), rng);
  int *dStart, *dShuf;
  CK(cudaMalloc(&dStart, nRuns * sizeof(int)));
  CK(cudaMalloc(&dShuf, nRuns * sizeof(int)));
  CK(cudaMemcpy(dStart, start.data(), nRuns * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dShuf, shuffled.data(), nRuns * sizeof(int), cudaMemcpyHostToDevice));
  char *flush;
  const size_t flushBytes = 256u << 20;
  CK(cudaMalloc(&flush, flushBytes));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](const char *name, auto launch) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaMemsetAsync(flush, rep, flushBytes));
      cudaEventRecord(a);
      launch();
      cudaEventRecord(b);
      CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0) best = std::min(best, ms);
    }
    printf("%-58s %8.1f us  %7.1f GB/s\n", name, best * 1e3, n * 8.0 / best / 1e6);
  };
  printf("runs %d x %d doubles = %.1f MB\n", nRuns, len, n * 8.0 / 1e6);
  for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
    char nm[128];
    snprintf(nm, sizeof nm, "A coalesced 8-byte stores, %d CTAs x 256", ctas);
    timeit(nm, [&] { kA<<<ctas, 256>>>(out, n); });
    snprintf(nm, sizeof nm, "B coalesced 16-byte stores, %d CTAs x 256", ctas);
    timeit(nm, [&] { kB<<<ctas, 256>>>((double2 *)out, n / 2); });
    snprintf(nm, sizeof nm, "C one warp per run, shuffled, %d CTAs x 256", ctas);
    timeit(nm, [&] { kC<<<ctas, 256>>>(out, dShuf, nRuns, len); });
    snprintf(nm, sizeof nm, "C one warp per run, memory order, %d CTAs x 256", ctas);
    timeit(nm, [&] { kC<<<ctas, 256>>>(out, dStart, nRuns, len); });
  }
  CK(cudaFuncSetAttribute(kD<97>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  for (int warps : {4, 8, 9}) {
    const size_t smem = sizeof(double) * 32 * 97 * warps;
    char nm[128];
    snprintf(nm, sizeof nm, "D smem tile write-out, shuffled, 148 CTAs x %d warps", warps);
    timeit(nm, [&] { kD<97><<<148, 32 * warps, smem>>>(out, dShuf, nRuns, len); });
    snprintf(nm, sizeof nm, "E smem tile write-out, memory order, 148 CTAs x %d warps", warps);
    timeit(nm, [&] { kD<97><<<148, 32 * warps, smem>>>(out, dStart, nRuns, len); });
  }
  const int imgCap = 32 * len + 2;
  CK(cudaFuncSetAttribute(kF<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(kF<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  for (int warps : {4, 8, 10}) {
    const size_t smem = sizeof(double) * (size_t)imgCap * warps;
    if (smem > 227 * 1024) continue;
    char nm[128];
    snprintf(nm, sizeof nm, "F linear image + cp.async.bulk store, 148 CTAs x %d warps", warps);
    timeit(nm, [&] { kF<true><<<148, 32 * warps, smem>>>(out, nRuns, len, imgCap); });
    snprintf(nm, sizeof nm, "G linear image + coalesced STG.64, 148 CTAs x %d warps", warps);
    timeit(nm, [&] { kF<false><<<148, 32 * warps, smem>>>(out, nRuns, len, imgCap); });
  }
  return 0;
}

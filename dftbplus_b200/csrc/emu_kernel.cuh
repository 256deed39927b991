// emu_kernel.cuh -- qualifiers and index macros of the kernels that are also run on the CPU by the
// test suite (tests/host_eval/): grid-stride loops that communicate through atomics only, no shared
// memory, no warp collectives.  Under nvcc these are ordinary CUDA kernels; under g++ the includer
// defines NL_TID / NL_STRIDE, NL_BITS / NL_FROM_BITS and the atomics, and runs the emulated threads one
// after the other in a shuffled order.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define NL_KERNEL __global__
#define NL_DEV __device__ __forceinline__
#define NL_TID ((long long)blockIdx.x * blockDim.x + threadIdx.x)
#define NL_STRIDE ((long long)gridDim.x * blockDim.x)
// IEEE operations that must not be contracted into FMAs (the reference is compiled without contraction
// across statements; the oracle with -ffp-contract=off)
#define NL_ADD(a, b) __dadd_rn((a), (b))
#define NL_SUB(a, b) __dsub_rn((a), (b))
#define NL_MUL(a, b) __dmul_rn((a), (b))
#define NL_BITS(x) ((unsigned long long)__double_as_longlong(x))
#define NL_FROM_BITS(u) __longlong_as_double((long long)(u))
#else
// host emulation: the includer defines NL_TID / NL_STRIDE and the atomics (tests/host_eval/host_nl.cpp)
#define NL_KERNEL
#define NL_DEV inline
#define NL_ADD(a, b) ((a) + (b))
#define NL_SUB(a, b) ((a) - (b))
#define NL_MUL(a, b) ((a) * (b))
#endif

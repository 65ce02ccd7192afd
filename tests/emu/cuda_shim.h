// Host shim that lets the *device* sources (sb_bdf.cuh, sb_kernels.cuh and the generated
// __device__ functions) be compiled by g++ and run on the CPU.  TEST INFRASTRUCTURE ONLY: it
// exists so that the integrator logic can be checked against the oracle in the CPU-only test
// tier (`-m "not gpu"`), where no GPU is available.  The product never loads this.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __grid_constant__

using std::fma; using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::pow;
using std::max; using std::min; using std::exp; using std::log; using std::log1p; using std::cbrt;

static inline double __ldg(const double* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }

struct EmuDim { unsigned x, y, z; };
static thread_local EmuDim blockIdx, blockDim, threadIdx, gridDim;

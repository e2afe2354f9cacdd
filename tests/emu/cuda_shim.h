// cuda_shim.h — TEST INFRASTRUCTURE.  Lets g++ compile watergap2_b200/csrc/wgk_kernels.cuh for
// the host so that the *kernel source itself* can be executed thread by thread on the CPU
// with glibc's libm and compared BIT-FOR-BIT with the oracle (tests/test_kernel_logic_cpu.py).
// This is a checker of the kernel logic only; nothing here is linked into libwgk.so and the
// product has no CPU execution path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline void __syncthreads() {}
static inline int __syncthreads_or(int p) { return p; }  // only in kernel drivers, which the emulation replaces
static inline void __pipeline_memcpy_async(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
static inline void __pipeline_commit() {}
static inline void __pipeline_wait_prior(int) {}
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
using std::exp; using std::pow; using std::sqrt; using std::fabs; using std::log; using std::cbrt; using std::fma;

// device_on_host.hpp — just enough of the CUDA device vocabulary for a HOST compiler to parse
// raftsql_b200/csrc/mrq_kernels.cuh, so that the per-group tick functions of that header (fast_group_tick,
// general_group_tick: the arithmetic the kernels run per thread) can be compiled for the host and run over host
// arrays against the oracle (tests/cpp/tick_host_test.cpp).  The kernels themselves (thread indexing, warp
// collectives, atomics, TMA) are parsed but never called here; the stand-ins below only have to make them compile.
// One "lane" per call: a warp-wide reduction over one lane is the lane's own value.
#pragma once
#define MRQ_HOST_EMULATION 1

#include <cuda_runtime.h>  // vector types and the qualifier macros as the host compiler sees them

#include <algorithm>
#include <cstdint>
#include <cstring>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

static const uint3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0};
static const dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline int __all_sync(unsigned, int p) { return p != 0; }
static inline uint32_t __reduce_or_sync(unsigned, uint32_t v) { return v; }
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v) { return v; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) {}
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
template <class T, class U>
static inline T atomicAdd(T *p, U v) {
  const T old = *p;
  *p = (T)(old + (T)v);
  return old;
}
template <class T, class U>
static inline T atomicMax(T *p, U v) {
  const T old = *p;
  if ((T)v > old) *p = (T)v;
  return old;
}

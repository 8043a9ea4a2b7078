// streams_bench.cu — development microbenchmark (not part of the product): what HBM rate does a
// "many parallel column streams" access pattern reach on this GPU, as a function of the number of read
// and write streams, the bytes per lane per stream (8 / 16 / 32), and whether half of the loads depend on the
// first half (the tick kernel's two load phases)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/streams_bench tools/streams_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int V>
struct Vec;
template <>
struct Vec<1> { typedef unsigned long long T; };
template <>
struct Vec<2> { typedef ulonglong2 T; };
template <>
struct Vec<4> { typedef ulonglong4 T; };

__device__ __forceinline__ unsigned long long fold(unsigned long long v) { return v; }
__device__ __forceinline__ unsigned long long fold(ulonglong2 v) { return v.x ^ v.y; }
__device__ __forceinline__ unsigned long long fold(ulonglong4 v) { return v.x ^ v.y ^ v.z ^ v.w; }
__device__ __forceinline__ void setv(unsigned long long &d, unsigned long long s) { d = s; }
__device__ __forceinline__ void setv(ulonglong2 &d, unsigned long long s) { d.x = s; d.y = s + 1; }
__device__ __forceinline__ void setv(ulonglong4 &d, unsigned long long s) { d.x = s; d.y = s + 1; d.z = s + 2; d.w = s + 3; }

template <int NR, int NW, int V, int PHASES, int THREADS>
__global__ void __launch_bounds__(THREADS) k(const unsigned long long *__restrict__ rd, unsigned long long *__restrict__ wr,
                                             size_t n, size_t stride) {
  typedef typename Vec<V>::T T;
  const size_t i = ((size_t)blockIdx.x * THREADS + threadIdx.x);
  if (i * V >= n) return;
  unsigned long long acc = 0;
  T v[NR];
  constexpr int H = PHASES == 2 ? (NR + 1) / 2 : NR;
#pragma unroll
  for (int s = 0; s < H; ++s) v[s] = *reinterpret_cast<const T *>(rd + (size_t)s * stride + i * V);
#pragma unroll
  for (int s = 0; s < H; ++s) acc ^= fold(v[s]);
  if (PHASES == 2) {
    const size_t off = (acc == 0x1234567ull) ? 1 : 0;  // data-dependent (never true): serialises the second wave
#pragma unroll
    for (int s = H; s < NR; ++s) v[s] = *reinterpret_cast<const T *>(rd + (size_t)s * stride + (i + off) * V);
#pragma unroll
    for (int s = H; s < NR; ++s) acc ^= fold(v[s]);
  }
#pragma unroll
  for (int s = 0; s < NW; ++s) {
    T o;
    setv(o, acc + s);
    *reinterpret_cast<T *>(wr + (size_t)s * stride + i * V) = o;
  }
  if (NW == 0 && acc == 0xdeadbeefull) wr[0] = acc;
}

template <int NR, int NW, int V, int PHASES, int THREADS>
void run(const char *name, unsigned long long *rd, unsigned long long *wr, size_t n, size_t stride, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t threads = (n + V - 1) / V;
  const unsigned grid = (unsigned)((threads + THREADS - 1) / THREADS);
  k<NR, NW, V, PHASES, THREADS><<<grid, THREADS>>>(rd, wr, n, stride);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) k<NR, NW, V, PHASES, THREADS><<<grid, THREADS>>>(rd, wr, n, stride);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)n * 8.0 * (NR + NW);
  printf("%-34s n=%zu NR=%2d NW=%2d V=%d ph=%d thr=%d : %8.2f us/launch  %7.1f GB/s\n", name, n, NR, NW, V, PHASES, THREADS,
         ms * 1e3 / iters, bytes * iters / (ms * 1e-3) / 1e9);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
}

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : (size_t)1 << 22;
  const size_t pad = argc > 2 ? strtoull(argv[2], 0, 10) : 64;  // column skew in 8-byte elements
  const size_t stride = n + pad;
  printf("--- stride = n + %zu elements\n", pad);
  unsigned long long *rd, *wr;
  cudaMalloc(&rd, stride * 8 * 28);
  cudaMalloc(&wr, stride * 8 * 12);
  cudaMemset(rd, 1, stride * 8 * 28);
  cudaMemset(wr, 0, stride * 8 * 12);
  const int it = 20;
  if (argc > 3) {  // short form: only the tick-like and K3-like patterns
    run<7, 0, 2, 1, 256>("K3-like 7r v2", rd, wr, n, stride, it);
    run<20, 8, 1, 2, 128>("20r 8w 2-phase (tick-like)", rd, wr, n, stride, it);
    run<12, 8, 1, 1, 128>("12r 8w", rd, wr, n, stride, it);
    return 0;
  }
  run<7, 0, 1, 1, 256>("K3-like 7r", rd, wr, n, stride, it);
  run<7, 0, 2, 1, 256>("K3-like 7r v2", rd, wr, n, stride, it);
  run<1, 1, 2, 1, 256>("copy 1r1w v2", rd, wr, n, stride, it);
  run<4, 4, 2, 1, 256>("copy 4r4w v2", rd, wr, n, stride, it);
  run<20, 0, 1, 1, 128>("20r", rd, wr, n, stride, it);
  run<20, 0, 1, 2, 128>("20r 2-phase", rd, wr, n, stride, it);
  run<20, 8, 1, 1, 128>("20r 8w", rd, wr, n, stride, it);
  run<20, 8, 1, 2, 128>("20r 8w 2-phase (tick-like)", rd, wr, n, stride, it);
  run<20, 8, 2, 2, 128>("20r 8w 2-phase v2", rd, wr, n, stride, it);
  run<20, 8, 2, 1, 128>("20r 8w 1-phase v2", rd, wr, n, stride, it);
  run<20, 8, 4, 1, 64>("20r 8w 1-phase v4", rd, wr, n, stride, it);
  run<12, 8, 1, 1, 128>("12r 8w", rd, wr, n, stride, it);
  run<12, 8, 2, 1, 128>("12r 8w v2", rd, wr, n, stride, it);
  run<12, 4, 2, 1, 128>("12r 4w v2", rd, wr, n, stride, it);
  run<8, 8, 1, 1, 256>("8r 8w", rd, wr, n, stride, it);
  run<8, 8, 2, 1, 256>("8r 8w v2", rd, wr, n, stride, it);
  run<20, 8, 1, 2, 256>("20r 8w 2-phase thr256", rd, wr, n, stride, it);
  run<20, 8, 1, 2, 512>("20r 8w 2-phase thr512", rd, wr, n, stride, it);
  return 0;
}

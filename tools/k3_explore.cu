// k3_explore.cu — development tool (not part of the product): variants of the standalone quorum kernel (K3) at the
// headline shape, 1,048,576 groups x 5 replicas, to find out where its fixed ~2.7 us goes and what moves it.
// It includes the product's kernel header and reuses its loads and its per-group arithmetic (quorum_commit_one),
// so every variant computes exactly what quorum_kernel_ldg256 computes; only the SHAPE of the kernel changes:
//
//   base      4 groups per thread, 128-thread CTAs, one CTA per 512 groups              (= the product's ldg256 form)
//   g8        8 groups per thread (two 256-bit loads per column in flight)
//   t256/t512 the base with larger CTAs
//   pers<k>   persistent grid, 148 * k CTAs, grid-stride over 512-group tiles (equal work per SM, no tail wave)
//   nowait    the base launched with programmatic stream serialisation and NO griddepcontrol.wait: consecutive
//             launches on independent column sets overlap their ramp-up and drain (only valid when the caller
//             guarantees independence, as this tool does: every launch reads a different set)
//
// Every timed launch reads a column set that has not been touched since an L2 flush; times are CUDA events over the
// timed launches on one stream.  Build and run on a B200:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/k3_explore tools/k3_explore.cu && /tmp/k3_explore
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../raftsql_b200/csrc/mrq_kernels.cuh"

using namespace mrq;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t st_ = (x);                                                                 \
    if (st_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(st_), __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int R = 5;

// 4 groups starting at i (i % 4 == 0, i + 3 < G): the body of the product's ldg256 form
__device__ __forceinline__ unsigned quad(const QuorumArgs &a, uint64_t i) {
  u64x4 mv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) mv[r] = ld_stream_v4(a.match + (uint64_t)r * a.gs + i);
  const u64x4 cm = ld_plain_v4(a.committed + i);
  const u64x4 gt = ld_stream_v4(a.term_start + i);
  unsigned nmoved = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint64_t m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) m[r] = mv[r].v[k];
    bool moved;
    const uint64_t c = quorum_commit_one<R>(m, cm.v[k], gt.v[k], moved);
    if (moved) st_state(a.committed + i + k, c);
    nmoved += moved;
  }
  return nmoved;
}

// ---- diagnostic variants: what bounds the 7-stream read + 1-stream write at this size? -----------------------------
// MODE 0: the product body; 1: no store (read-only); 2: loads + xor only (no selection network, no store);
// 3: loads carry the .L2::256B prefetch hint; 4: stores carry an evict-first policy (st.global.cs);
// 5: 3 + 4 + one 256-bit store per quad when all four groups moved (what the product's ldg256 form now does)
__device__ __forceinline__ u64x4 ld_v4_pf256(const uint64_t *p) {
  u64x4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.b64 {%0, %1, %2, %3}, [%4];"
               : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3])
               : "l"(p));
  return r;
}
template <int MODE>
__device__ __forceinline__ unsigned quad_x(const QuorumArgs &a, uint64_t i) {
  u64x4 mv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) mv[r] = (MODE == 3 || MODE == 5) ? ld_v4_pf256(a.match + (uint64_t)r * a.gs + i) : ld_stream_v4(a.match + (uint64_t)r * a.gs + i);
  const u64x4 cm = ld_plain_v4(a.committed + i);
  const u64x4 gt = (MODE == 3 || MODE == 5) ? ld_v4_pf256(a.term_start + i) : ld_stream_v4(a.term_start + i);
  unsigned nmoved = 0;
  if (MODE == 2) {
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int r = 0; r < R; ++r) x ^= mv[r].v[k];
      x ^= cm.v[k] ^ gt.v[k];
    }
    return (unsigned)(x == 0x123456789ull);
  }
  if (MODE == 5) {
    u64x4 out;
    bool mvd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t m[R];
#pragma unroll
      for (int r = 0; r < R; ++r) m[r] = mv[r].v[k];
      out.v[k] = quorum_commit_one<R>(m, cm.v[k], gt.v[k], mvd[k]);
      nmoved += mvd[k];
    }
    if (nmoved == 4u) {
      st_stream_v4(a.committed + i, out);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (mvd[k]) st_stream_u64(a.committed + i + k, out.v[k]);
    }
    return nmoved;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint64_t m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) m[r] = mv[r].v[k];
    bool moved;
    const uint64_t c = quorum_commit_one<R>(m, cm.v[k], gt.v[k], moved);
    if (MODE == 1) {
      nmoved += moved;
      continue;
    }
    if (moved) {
      if (MODE == 4)
        asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(a.committed + i + k), "l"(c) : "memory");
      else
        st_state(a.committed + i + k, c);
    }
    nmoved += moved;
  }
  return nmoved;
}
template <int THREADS, int MINB, int MODE>
__global__ void __launch_bounds__(THREADS, MINB) k3_diag(const QuorumArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t i = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * 4;
  unsigned n = 0;
  if (i + 3 < a.G) n = quad_x<MODE>(a, i);
  count_moved(a.ctr, n);
}
// one contiguous array of the same 58.7 MB, 256-bit loads: the pure read-stream ceiling of this GPU at this size
__global__ void __launch_bounds__(128) k3_onestream(const QuorumArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t n4 = a.G * (R + 2) / 4;  // 256-bit words in match[R] + 2 columns' worth
  uint64_t x = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * 128 + threadIdx.x; q < n4; q += (uint64_t)gridDim.x * 128) {
    const uint64_t m4 = a.G * R / 4, g4 = a.G / 4;  // match is [R][G]; then term_start, then committed: three arrays, each contiguous
    const uint64_t *src = q < m4 ? a.match + q * 4 : (q < m4 + g4 ? a.term_start + (q - m4) * 4 : a.committed + (q - m4 - g4) * 4);
    const u64x4 v = ld_stream_v4(src);
    x ^= v.v[0] ^ v.v[1] ^ v.v[2] ^ v.v[3];
  }
  count_moved(a.ctr, (unsigned)(x == 0x123456789ull));
}

template <int THREADS, int QUADS, bool WAIT>
__global__ void __launch_bounds__(THREADS) k3_tile(const QuorumArgs a) {
  pdl_launch_dependents();
  if (WAIT) pdl_wait();
  const uint64_t base = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * 4;
  const uint64_t stride = (uint64_t)gridDim.x * THREADS * 4;
  unsigned n = 0;
#pragma unroll
  for (int q = 0; q < QUADS; ++q) {
    const uint64_t i = base + q * stride;  // interleaved: consecutive threads stay on consecutive quads
    if (i + 3 < a.G) n += quad(a, i);
  }
  count_moved(a.ctr, n);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k3_persistent(const QuorumArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  unsigned n = 0;
  for (uint64_t i = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * 4; i + 3 < a.G; i += (uint64_t)gridDim.x * THREADS * 4) n += quad(a, i);
  count_moved(a.ctr, n);
}

// persistent grid over 512-group tiles with the shipped body (prefetch hint + streaming 256-bit store)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k3_persistent5(const QuorumArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  unsigned n = 0;
  for (uint64_t i = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) * 4; i + 3 < a.G; i += (uint64_t)gridDim.x * THREADS * 4) n += quad_x<5>(a, i);
  count_moved(a.ctr, n);
}

struct Set {
  uint64_t *match, *committed, *gate, *committed0;
};

template <typename K>
float time_variant(const char *name, K kern, unsigned grid, unsigned block, bool pdl, std::vector<Set> &sets, uint64_t G, Counters *ctr,
                   uint8_t *flush, size_t flush_bytes, cudaStream_t st, double peak_gbs) {
  const int warm = 5, timed = (int)sets.size() - warm;
  float best = 1e9f, sum = 0;
  for (int rep = 0; rep < 3; ++rep) {
    for (auto &s : sets) CK(cudaMemcpyAsync(s.committed, s.committed0, G * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(flush, rep + 1, flush_bytes, st));  // push everything out of L2
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int k = 0; k < (int)sets.size(); ++k) {
      if (k == warm) CK(cudaEventRecord(e0, st));
      QuorumArgs a{sets[k].match, sets[k].committed, sets[k].gate, ctr, G, G};
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(block);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 1 : 0;
      CK(cudaLaunchKernelEx(&cfg, kern, a));
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const float us = 1e3f * ms / timed;
    best = us < best ? us : best;
    sum += us;
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
  }
  const double bytes = (8.0 * R + 16.0) * (double)G;
  const float mean = sum / 3;
  printf("%-28s grid %6u x %3u  pdl %d   mean %6.2f us  best %6.2f us   %7.1f GB/s  frac %.3f\n", name, grid, block, (int)pdl, mean, best,
         bytes / (mean * 1e-6) / 1e9, bytes / (mean * 1e-6) / 1e9 / peak_gbs);
  return mean;
}

int main(int argc, char **argv) {
  const uint64_t G = 1u << 20;
  const double peak = argc > 1 ? atof(argv[1]) : 6583.5;  // MEASURED_PEAKS.json hbm_gbs
  int sm = 148;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  sm = prop.multiProcessorCount;
  printf("%s, %d SMs; %llu groups x %d replicas, %.1f MB per launch, peak %.1f GB/s\n", prop.name, sm, (unsigned long long)G, R,
         (8.0 * R + 16) * G / 1e6, peak);
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int nsets = 25;
  std::vector<Set> sets(nsets);
  std::vector<uint64_t> h_match((size_t)R * G), h_c(G), h_g(G);
  uint64_t x = 12345;
  auto rnd = [&]() { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; return x * 0x2545F4914F6CDD1Dull; };
  for (auto &s : sets) {
    for (uint64_t g = 0; g < G; ++g) {
      const uint64_t li = (1ull << 20) + rnd() % ((1ull << 40) - (1ull << 20));
      for (int r = 0; r < R; ++r) h_match[(size_t)r * G + g] = li - rnd() % 12;
      h_c[g] = li - 40;
      h_g[g] = li - 45;
    }
    CK(cudaMalloc(&s.match, (size_t)R * G * 8));
    CK(cudaMalloc(&s.committed, G * 8));
    CK(cudaMalloc(&s.committed0, G * 8));
    CK(cudaMalloc(&s.gate, G * 8));
    CK(cudaMemcpy(s.match, h_match.data(), (size_t)R * G * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.committed0, h_c.data(), G * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.gate, h_g.data(), G * 8, cudaMemcpyHostToDevice));
  }
  Counters *ctr;
  CK(cudaMalloc(&ctr, sizeof(Counters) * kCtrShards));
  CK(cudaMemset(ctr, 0, sizeof(Counters) * kCtrShards));
  const size_t flush_bytes = 512u << 20;
  uint8_t *flush;
  CK(cudaMalloc(&flush, flush_bytes));

  const unsigned quads = (unsigned)(G / 4);
  auto blocks = [&](unsigned threads, unsigned qpt) { return (quads + threads * qpt - 1) / (threads * qpt); };
#define RUN(name, kern, grid, block, pdl) time_variant(name, kern, grid, block, pdl, sets, G, ctr, flush, flush_bytes, st, peak)
  RUN("base (4/thread, 128 thr)", (k3_tile<128, 1, true>), blocks(128, 1), 128, true);
  RUN("base, no PDL attribute", (k3_tile<128, 1, true>), blocks(128, 1), 128, false);
  RUN("t256", (k3_tile<256, 1, true>), blocks(256, 1), 256, true);
  RUN("t512", (k3_tile<512, 1, true>), blocks(512, 1), 512, true);
  RUN("g8 (8/thread, 128 thr)", (k3_tile<128, 2, true>), blocks(128, 2), 128, true);
  RUN("g16 (16/thread, 128 thr)", (k3_tile<128, 4, true>), blocks(128, 4), 128, true);
  for (unsigned k : {2u, 4u, 8u, 16u}) {
    char name[64];
    snprintf(name, sizeof name, "pers%u (148*%u CTAs, 128 thr)", k, k);
    RUN(name, (k3_persistent<128>), (unsigned)sm * k, 128, true);
  }
  RUN("pers4, 256 thr", (k3_persistent<256>), (unsigned)sm * 4, 256, true);
  RUN("diag: read-only (no store)", (k3_diag<128, 1, 1>), blocks(128, 1), 128, true);
  RUN("diag: loads + xor only", (k3_diag<128, 1, 2>), blocks(128, 1), 128, true);
  RUN("diag: .L2::256B prefetch hint", (k3_diag<128, 1, 3>), blocks(128, 1), 128, true);
  RUN("diag: st.global.cs stores", (k3_diag<128, 1, 4>), blocks(128, 1), 128, true);
  RUN("diag: pf256 + cs + v4 store", (k3_diag<128, 1, 5>), blocks(128, 1), 128, true);
  RUN("diag: same, 256 threads", (k3_diag<256, 1, 5>), blocks(256, 1), 256, true);
  RUN("diag: shipped body, persistent 148*8", (k3_persistent5<128>), (unsigned)sm * 8, 128, true);
  RUN("diag: shipped body, persistent 148*7", (k3_persistent5<128>), (unsigned)sm * 7, 128, true);
  RUN("diag: shipped body, pers 148*4 x 256", (k3_persistent5<256>), (unsigned)sm * 4, 256, true);
  RUN("diag: <=64 regs (8 CTAs/SM)", (k3_diag<128, 8, 0>), blocks(128, 1), 128, true);
  RUN("diag: one contiguous stream*", (k3_onestream), (unsigned)sm * 8, 128, true);
  RUN("nowait (independent sets)", (k3_tile<128, 1, false>), blocks(128, 1), 128, true);
  RUN("nowait g8", (k3_tile<128, 2, false>), blocks(128, 2), 128, true);
  unsigned long long moved = 0;
  std::vector<Counters> hc(kCtrShards);
  CK(cudaMemcpy(hc.data(), ctr, sizeof(Counters) * kCtrShards, cudaMemcpyDeviceToHost));
  for (auto &c : hc) moved += c.commits_advanced;
  printf("commits advanced in all (sanity: every launch advances every group): %llu\n", moved);
  return 0;
}

// tools/gather_probe.cu -- microbenchmark behind the "random 32-byte sector" roofline denominator:
// how many DRAM bytes does one random 32-byte load really cost on B200, and which knobs change it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gather_probe tools/gather_probe.cu
//   ./tools/gather_probe [GiB]          (under ncu: dram__bytes_read.sum per kernel)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t& x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
  return z;
}

enum Mode { kNc256 = 0, kPlain256 = 1, kNc128 = 2, kNc64 = 3, kCg128 = 4, kLine128B = 5 };

template <int kMode, int kMlp>
__global__ void __launch_bounds__(256) k_probe(const uint8_t* __restrict__ base, uint64_t n_sectors, int iters, uint64_t seed, unsigned long long* sink) {
  uint64_t x = seed + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0xD1B54A32D192ED03ull;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    const uint8_t* p[kMlp];
#pragma unroll
    for (int b = 0; b < kMlp; ++b) p[b] = base + __umul64hi(mix(x), n_sectors) * 32;
#pragma unroll
    for (int b = 0; b < kMlp; ++b) {
      unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      if (kMode == kNc256) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p[b]));
      if (kMode == kPlain256) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p[b]));
      if (kMode == kNc128) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p[b]));
      if (kMode == kNc64) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(w0) : "l"(p[b]));
      if (kMode == kCg128) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p[b]));
      if (kMode == kLine128B) {   // whole 128-byte line: four 256-bit loads
        const uint8_t* q = (const uint8_t*)((uintptr_t)p[b] & ~(uintptr_t)127);
#pragma unroll
        for (int o = 0; o < 128; o += 32) {
          unsigned long long a, b2, c, d;
          asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b2), "=l"(c), "=l"(d) : "l"(q + o));
          w0 ^= a ^ b2 ^ c ^ d;
        }
      }
      acc += w0 ^ w1 ^ w2 ^ w3;
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <int kMode, int kMlp>
int run(const char* name, const uint8_t* buf, uint64_t n_sectors, int blocks_per_sm, unsigned long long* sink) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm, iters = 2048 / kMlp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_probe<kMode, kMlp><<<blocks, 256>>>(buf, n_sectors, iters, 77 + rep, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double probes = (double)blocks * 256 * iters * kMlp;
  printf("%-28s mlp %2d ctas/sm %d : %7.2f G probes/s  (%.3f ms)\n", name, kMlp, blocks_per_sm, probes / best / 1e6, best);
  return 0;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 8.0;
  const uint64_t bytes = (uint64_t)(gib * (1ull << 30));
  uint8_t* buf; unsigned long long* sink;
  CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(buf, 1, bytes));
  const uint64_t n_sectors = bytes / 32;
  size_t lim = 0;
  cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("table %.1f GiB; default cudaLimitMaxL2FetchGranularity = %zu\n", gib, lim);
  for (int g : {0, 32, 128}) {
    if (g) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
      cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
      printf("--- set granularity %d -> %s, now %zu\n", g, cudaGetErrorString(e), lim);
    }
    run<kNc256, 4>("nc.256b", buf, n_sectors, 8, sink);
    run<kNc256, 8>("nc.256b", buf, n_sectors, 8, sink);
    run<kNc256, 16>("nc.256b", buf, n_sectors, 8, sink);
    run<kPlain256, 8>("plain.256b", buf, n_sectors, 8, sink);
    run<kNc128, 8>("nc.128b", buf, n_sectors, 8, sink);
    run<kNc64, 8>("nc.64b", buf, n_sectors, 8, sink);
    run<kCg128, 8>("cg.128b", buf, n_sectors, 8, sink);
    run<kLine128B, 4>("line128B(4x256b)", buf, n_sectors, 8, sink);
  }
  return 0;
}

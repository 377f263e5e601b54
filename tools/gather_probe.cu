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
// r2: whole-line probes issued COOPERATIVELY, so that the lanes' accesses of one warp instruction coalesce into one
// request per 128-byte line: kCoop lanes share a probe, each loads 128 / kCoop ... bytes of the line (or of its first half).
enum CoopMode { kCoop4x32 = 0, kCoop8x16 = 1, kCoop2x32Half = 2, kCoop4x16Half = 3, kCoop16x8 = 4, kThread2x32Half = 5 };

template <int kMode, int kMlp>
__global__ void __launch_bounds__(256) k_probe_coop(const uint8_t* __restrict__ base, uint64_t n_lines, int iters, uint64_t seed, unsigned long long* sink) {
  constexpr int kLanes = kMode == kCoop4x32 ? 4 : kMode == kCoop8x16 ? 8 : kMode == kCoop2x32Half ? 2 : kMode == kCoop4x16Half ? 4 : kMode == kCoop16x8 ? 16 : 1;
  const int lane = threadIdx.x & 31;
  const int sub = lane % kLanes;
  // every lane owns a probe stream; in round r of an iteration the group adopts the probe of its lane r
  uint64_t x = seed + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0xD1B54A32D192ED03ull;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint64_t mine[kMlp];
#pragma unroll
    for (int b = 0; b < kMlp; ++b) mine[b] = __umul64hi(mix(x), n_lines) * 128;
#pragma unroll
    for (int r = 0; r < kLanes; ++r) {
      unsigned long long w[kMlp][4];
#pragma unroll
      for (int b = 0; b < kMlp; ++b) {
        const uint64_t off = kLanes == 1 ? mine[b] : __shfl_sync(0xffffffffu, mine[b], (lane & ~(kLanes - 1)) + r);
        const uint8_t* q = base + off;
        w[b][0] = w[b][1] = w[b][2] = w[b][3] = 0;
        if (kMode == kCoop4x32 || kMode == kCoop2x32Half)
          asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(q + sub * 32));
        if (kMode == kCoop8x16 || kMode == kCoop4x16Half)
          asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(w[b][0]), "=l"(w[b][1]) : "l"(q + sub * 16));
        if (kMode == kCoop16x8)
          asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(w[b][0]) : "l"(q + sub * 8));
        if (kMode == kThread2x32Half) {
          unsigned long long a2, b2, c2, d2;
          asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(q));
          asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a2), "=l"(b2), "=l"(c2), "=l"(d2) : "l"(q + 32));
          w[b][0] ^= a2 ^ b2 ^ c2 ^ d2;
        }
      }
#pragma unroll
      for (int b = 0; b < kMlp; ++b) acc += w[b][0] ^ w[b][1] ^ w[b][2] ^ w[b][3];
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <int kMode, int kMlp>
int run_coop(const char* name, const uint8_t* buf, uint64_t n_lines, int blocks_per_sm, unsigned long long* sink) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm, iters = 1024 / kMlp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_probe_coop<kMode, kMlp><<<blocks, 256>>>(buf, n_lines, iters, 77 + rep, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double probes = (double)blocks * 256 * iters * kMlp;     // every lane's probes are all served (kLanes rounds)
  printf("%-28s mlp %2d ctas/sm %d : %7.2f G probes/s  (%.3f ms)\n", name, kMlp, blocks_per_sm, probes / best / 1e6, best);
  return 0;
}

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
  const bool coop_only = argc > 2;
  const uint64_t bytes = (uint64_t)(gib * (1ull << 30));
  uint8_t* buf; unsigned long long* sink;
  CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(buf, 1, bytes));
  const uint64_t n_sectors = bytes / 32;
  size_t lim = 0;
  cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("table %.2f GiB; default cudaLimitMaxL2FetchGranularity = %zu\n", gib, lim);
  if (!coop_only) {
    run<kNc256, 4>("nc.256b", buf, n_sectors, 8, sink);
    run<kNc256, 8>("nc.256b", buf, n_sectors, 8, sink);
    run<kNc128, 8>("nc.128b", buf, n_sectors, 8, sink);
    run<kNc64, 8>("nc.64b", buf, n_sectors, 8, sink);
    run<kLine128B, 4>("line128B(4x256b)", buf, n_sectors, 8, sink);
  }
  const uint64_t n_lines = bytes / 128;
  run_coop<kCoop4x32, 2>("coop 4 lanes x 32B (line)", buf, n_lines, 8, sink);
  run_coop<kCoop4x32, 4>("coop 4 lanes x 32B (line)", buf, n_lines, 8, sink);
  run_coop<kCoop8x16, 2>("coop 8 lanes x 16B (line)", buf, n_lines, 8, sink);
  run_coop<kCoop8x16, 4>("coop 8 lanes x 16B (line)", buf, n_lines, 8, sink);
  run_coop<kCoop16x8, 4>("coop 16 lanes x 8B (line)", buf, n_lines, 8, sink);
  run_coop<kCoop2x32Half, 4>("coop 2 lanes x 32B (64B)", buf, n_lines, 8, sink);
  run_coop<kCoop4x16Half, 4>("coop 4 lanes x 16B (64B)", buf, n_lines, 8, sink);
  run_coop<kThread2x32Half, 4>("1 thread 2 x 32B (64B)", buf, n_lines, 8, sink);
  run_coop<kThread2x32Half, 8>("1 thread 2 x 32B (64B)", buf, n_lines, 8, sink);
  return 0;
}

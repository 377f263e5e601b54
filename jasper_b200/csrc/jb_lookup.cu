// jb_lookup.cu -- lookup entry points: jb_lookup / jb_lookup_flat (exact
// qf[jf.MerDNA(s).get_canonical()], JF!swig/mer_file.i:41 + JF!swig/mer_dna.i:15,71) and
// jb_count_positions (the same for every window of a sequence), plus the random-sector gather
// micro-benchmark that provides the roofline denominator for the probe kernels.
#include <algorithm>
#include <cstring>

#include "jb_scan.cuh"

namespace jb {

// one thread per string: MerDNA(const char*) -> canonical -> probe
__global__ void k_lookup_strings(TableView t, const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                                 uint64_t n, uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Key m = key_from_chars(buf + off[i], (int64_t)(off[i + 1] - off[i]), t.g.k, t.mask);
  out[i] = table_count(t, key_canonical(m, t.g.k));
}

// Random sector gather: every thread issues `per_thread` independent loads of `sector` bytes at
// pseudo-random sector-aligned offsets of the table (the access pattern of a perfect probe).
template <int kSectorBytes>
__global__ void k_gather(const uint8_t* __restrict__ base, uint64_t n_sectors, uint64_t per_thread, uint64_t seed,
                         unsigned long long* sink) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = seed + tid * 0x9E3779B97F4A7C15ull;
  unsigned long long acc = 0;
  for (uint64_t it = 0; it < per_thread; it += 4) {
    const uint8_t* p[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
      p[b] = base + mulhi64(x, n_sectors) * kSectorBytes;
      x += 0x9E3779B97F4A7C15ull;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int o = 0; o < kSectorBytes; o += 32) {
        unsigned long long w0, w1, w2, w3;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p[b] + o));
        acc += w0 ^ w1 ^ w2 ^ w3;
      }
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

// Random LINE gather the way k_scan probes a compact table: four adjacent lanes share a probe and load 32 bytes each of
// one 128-byte line in a single warp instruction (one request per line); every lane owns a stream of probes.
__global__ void k_gather_line(const uint8_t* __restrict__ base, uint64_t n_lines, uint64_t per_thread, uint64_t seed,
                              unsigned long long* sink) {
  const int lane = threadIdx.x & 31, sub = lane & 3;
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = seed + tid * 0x9E3779B97F4A7C15ull;
  unsigned long long acc = 0;
  for (uint64_t it = 0; it < per_thread; it += 4) {
    uint64_t mine[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
      mine[b] = mulhi64(x, n_lines) * 128;
      x += 0x9E3779B97F4A7C15ull;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      unsigned long long w[4][4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint64_t off = __shfl_sync(0xffffffffu, mine[b], (lane & 28) | r);
        ldg256(base + off + sub * 32, w[b][0], w[b][1], w[b][2], w[b][3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) acc += w[b][0] ^ w[b][1] ^ w[b][2] ^ w[b][3];
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

}  // namespace jb

using namespace jb;

extern "C" int jb_lookup_flat(const jb_db* db, const char* buf, const uint64_t* offsets, size_t n, uint32_t* counts) {
  if (int rc_ready = jb_check_ready(db)) return rc_ready;
  if (!db || (!buf && n && offsets[n]) || !offsets || !counts) return fail(JB_E_ARG, "jb_lookup_flat: null argument");
  if (n == 0) return JB_OK;
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  const uint64_t nbytes = offsets[n];
  uint8_t* d_buf = nullptr; uint64_t* d_off = nullptr; uint32_t* d_out = nullptr;
  JB_CUDA(cudaMalloc(&d_buf, nbytes + 16));
  JB_CUDA(cudaMalloc(&d_off, (n + 1) * 8));
  JB_CUDA(cudaMalloc(&d_out, n * 4));
  JB_CUDA(cudaMemcpyAsync(d_buf, buf, nbytes, cudaMemcpyHostToDevice, s));
  JB_CUDA(cudaMemcpyAsync(d_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
  k_lookup_strings<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(db->view(), d_buf, d_off, n, d_out);
  JB_CUDA(cudaGetLastError());
  JB_CUDA(cudaMemcpyAsync(counts, d_out, n * 4, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_buf); cudaFree(d_off); cudaFree(d_out);
  add_counter("kernel_launches", 1);
  return JB_OK;
}

extern "C" int jb_lookup(const jb_db* db, const char* const* strs, size_t n, uint32_t* counts) {
  if (!db || !strs || !counts) return fail(JB_E_ARG, "jb_lookup: null argument");
  const int k = db->hdr.k();
  std::vector<uint64_t> off(n + 1, 0);
  std::string flat;
  for (size_t i = 0; i < n; ++i) {
    size_t len = strnlen(strs[i], (size_t)k);       // MerDNA never reads past k bytes or the NUL
    flat.append(strs[i], len);
    off[i + 1] = flat.size();
  }
  return jb_lookup_flat(db, flat.data(), off.data(), n, counts);
}

extern "C" int jb_count_positions(const jb_db* db, const char* seq, size_t len, uint32_t* counts) {
  if (int rc_ready = jb_check_ready(db)) return rc_ready;
  if (!db || (!seq && len) || !counts) return fail(JB_E_ARG, "jb_count_positions: null argument");
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  std::vector<ScanUnit> units{ScanUnit{0, (uint64_t)len, 0}};
  std::vector<uint32_t> tiles = make_tile_prefix(units);
  uint8_t* d_seq = nullptr; uint32_t* d_cnt = nullptr; ScanUnit* d_units = nullptr; uint32_t* d_tiles = nullptr;
  cudaEvent_t e0, e1, e2, e3;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1)); JB_CUDA(cudaEventCreate(&e2)); JB_CUDA(cudaEventCreate(&e3));
  JB_CUDA(cudaMalloc(&d_seq, len + kSeqSlice + 16));      // (k_scan fetches whole tile slices: room behind the last base)
  JB_CUDA(cudaMalloc(&d_cnt, (len + 1) * 4));
  JB_CUDA(cudaMalloc(&d_units, sizeof(ScanUnit)));
  JB_CUDA(cudaMalloc(&d_tiles, tiles.size() * sizeof(uint32_t)));
  JB_CUDA(cudaEventRecord(e0, s));
  if (len) JB_CUDA(cudaMemcpyAsync(d_seq, seq, len, cudaMemcpyHostToDevice, s));
  JB_CUDA(cudaMemcpyAsync(d_units, units.data(), sizeof(ScanUnit), cudaMemcpyHostToDevice, s));
  JB_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  ScanArgs a{};
  a.bases = d_seq; a.units = d_units; a.tile_first = d_tiles; a.n_units = 1; a.n_tiles = tiles.back(); a.counts = d_cnt;
  JB_CUDA(cudaEventRecord(e1, s));
  if (int rc = launch_scan(db->view(), a, s)) return rc;
  JB_CUDA(cudaEventRecord(e2, s));
  JB_CUDA(cudaMemcpyAsync(counts, d_cnt, (len + 1) * 4, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaEventRecord(e3, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e1, e2); note_timing("count_positions", ms);
  cudaEventElapsedTime(&ms, e0, e1); note_timing("h2d", ms);
  cudaEventElapsedTime(&ms, e2, e3); note_timing("d2h", ms);
  note_counter("scan_lookups", (int64_t)len + 1);
  note_counter("kernel_launches", 1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
  cudaFree(d_seq); cudaFree(d_cnt); cudaFree(d_units); cudaFree(d_tiles);
  return JB_OK;
}

extern "C" int jb_gather_bench(const jb_db* db, uint64_t n_probes, int sector_bytes, int iters, double* gbps, double* gprobes) {
  if (!db || !gbps || !gprobes) return fail(JB_E_ARG, "jb_gather_bench: null argument");
  if (sector_bytes != 32 && sector_bytes != 64 && sector_bytes != 128) return fail(JB_E_ARG, "sector_bytes must be 32, 64 or 128");
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  const uint64_t table_bytes = db->my_units() * (uint64_t)db->unit_bytes();
  const uint64_t n_sectors = table_bytes / sector_bytes;
  const int threads = 256, blocks = 148 * 8;
  uint64_t per_thread = std::max<uint64_t>(4, (n_probes / ((uint64_t)threads * blocks) + 3) / 4 * 4);
  unsigned long long* d_sink = nullptr;
  JB_CUDA(cudaMalloc(&d_sink, 8));
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < iters + 1; ++it) {
    JB_CUDA(cudaEventRecord(e0, s));
    const uint8_t* base = db->mem;
    if (sector_bytes == 32) k_gather<32><<<blocks, threads, 0, s>>>(base, n_sectors, per_thread, 1234 + it, d_sink);
    else if (sector_bytes == 64) k_gather<64><<<blocks, threads, 0, s>>>(base, n_sectors, per_thread, 1234 + it, d_sink);
    else k_gather_line<<<blocks, threads, 0, s>>>(base, n_sectors, per_thread, 1234 + it, d_sink);      // whole lines, four lanes each
    JB_CUDA(cudaEventRecord(e1, s));
    JB_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0) best = std::min(best, ms);
  }
  const double total = (double)per_thread * threads * blocks;
  *gprobes = total / (best * 1e-3) / 1e9;
  *gbps = total * sector_bytes / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_sink);
  return JB_OK;
}

// jb_index.cu -- builds the scan index (jb_index.cuh) on the device from the resident table: two sweeps over the table's
// records (bucket sizes; insertion) around one prefix sum.  No file I/O, no sort.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <vector>

#include "jb_geom.hpp"
#include "jb_index.cuh"
#include "jb_internal.cuh"

namespace jb {
namespace {

struct IndexBuild {
  ScanIndex ix;              // being built (slots / offsets writable through the casts below)
  unsigned int* bucket_n;    // records per bucket (first sweep)
  int k;
};

__device__ __forceinline__ void index_visit(const IndexBuild& b, const Key& key, uint32_t count, int phase) {
  const Key rc = key_revcomp(key, b.k);
  const uint64_t bucket = index_bucket(window_min_hash(key, rc, b.k, b.ix.m, nullptr), b.ix.bucket_bits);
  if (phase == 0) { atomicAdd(b.bucket_n + bucket, 1u); return; }
  uint64_t beg; uint32_t cap;
  index_bounds(b.ix, bucket, &beg, &cap);
  uint4* slots = const_cast<uint4*>(b.ix.slots);
  uint32_t s = 2 * __umulhi(index_slot_hash(key), cap / 2);
  // claim the first empty slot from the home pair on (the count word is the claim; the keys of a table are distinct, so nobody
  // compares keys during the build)
  while (atomicCAS(&slots[beg + s].w, 0u, count) != 0u) s = s + 1 == cap ? 0 : s + 1;
  slots[beg + s].x = (uint32_t)key.lo; slots[beg + s].y = (uint32_t)(key.lo >> 32); slots[beg + s].z = (uint32_t)key.hi;
}

// wide slots: a wide table, or the stash list of a compact one
__global__ void k_index_wide(const Slot* __restrict__ slots, uint64_t n, IndexBuild b, int phase) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const Slot s = slots[i];
    if (s.count) index_visit(b, Key{s.lo, s.hi}, s.count, phase);
  }
}

// compact slots: the key comes back from (sector, tag, remainder) through the inverse of the hash matrix's low block, the
// way dump_compact (jb_table.cu) does it on the host.  Escaped counts and displaced records are taken from the stash list.
struct CompactDecode { TableGeom g; const uint64_t* lut; const uint64_t* inv; uint64_t n_units; };
__global__ void k_index_compact(const unsigned long long* __restrict__ words, uint64_t n_words, CompactDecode d, IndexBuild b, int phase,
                                unsigned long long* n_bad) {
  const TableGeom& g = d.g;
  const uint32_t cmask = cval_mask(g);
  const int r = g.r_bits, bits = 64 - g.left_shift;
  const uint64_t tag_mask = g.tag_bits >= 64 ? ~0ull : (1ull << g.tag_bits) - 1;
  const uint64_t rmask = r >= 64 ? ~0ull : (1ull << r) - 1;
  for (uint64_t sidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < n_words; sidx += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long w = words[sidx];
    if (w == 0) continue;
    const uint32_t cv = (uint32_t)w & cmask;
    if (cv >= cmask - 1u) continue;                                   // tombstone / escaped count: the record is in the stash list
    const uint64_t pat = w >> g.cnt_bits, tag = pat & tag_mask, rem = g.tag_bits >= 64 ? 0 : pat >> g.tag_bits;
    // the posx in the window of home sectors [line - kMaxHops, line] that ends in `tag`
    const uint64_t line = sidx / kCompactSlots, first_line = line > (uint64_t)kMaxHops ? line - kMaxHops : 0;
    const unsigned __int128 span = (unsigned __int128)1 << bits;
    const uint64_t lo_px = (uint64_t)(((unsigned __int128)first_line * span + g.units - 1) / g.units);   // smallest posx whose home sector >= first_line
    uint64_t px = (lo_px & ~tag_mask) | tag;
    if (px < lo_px) px += tag_mask + 1;
    const uint64_t pos = px >> g.ext_bits;
    Key key;                                                          // high bits from the remainder, low r bits solved
    if (r == 0) { key.lo = rem; key.hi = 0; }
    else if (r >= 64) { key.lo = 0; key.hi = rem << (r - 64); }
    else { key.lo = rem << r; key.hi = rem >> (64 - r); }
    const uint64_t v = (pos ^ hash_lut(d.lut, g.key_bytes, key)) & rmask;
    uint64_t low = 0;
    for (int q = 0; q < r; ++q) low |= (uint64_t)(__popcll(d.inv[q] & v) & 1) << q;
    key.lo |= low;
    if (phase == 0 && key_posx(g, hash_lut(d.lut, g.key_bytes, key), key) != px) { atomicAdd(n_bad, 1ull); continue; }   // cannot happen
    index_visit(b, key, cv, phase);
  }
}

// ---- exclusive prefix sums of the bucket capacities (64-bit), three launches
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(256) k_cap_block_sums(const unsigned int* __restrict__ n, uint64_t n_buckets, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long part[256];
  const uint64_t first = (uint64_t)blockIdx.x * kScanBlock;
  unsigned long long s = 0;
  for (int q = threadIdx.x; q < kScanBlock; q += 256) if (first + q < n_buckets) s += index_bucket_capacity(n[first + q]);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d; d >>= 1) { if ((int)threadIdx.x < d) part[threadIdx.x] += part[threadIdx.x + d]; __syncthreads(); }
  if (threadIdx.x == 0) sums[blockIdx.x] = part[0];
}
__global__ void __launch_bounds__(1024) k_cap_scan_sums(unsigned long long* sums, uint64_t n_blocks, unsigned long long* total) {
  __shared__ unsigned long long part[1024];
  const uint64_t per = (n_blocks + 1023) / 1024, lo = min(n_blocks, threadIdx.x * per), hi = min(n_blocks, lo + per);
  unsigned long long s = 0;
  for (uint64_t q = lo; q < hi; ++q) s += sums[q];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned long long add = (int)threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  unsigned long long run = part[threadIdx.x] - s;
  for (uint64_t q = lo; q < hi; ++q) { const unsigned long long v = sums[q]; sums[q] = run; run += v; }
  if (threadIdx.x == 1023) *total = part[1023];
}
template <class Off>
__global__ void __launch_bounds__(256) k_cap_offsets(const unsigned int* __restrict__ n, uint64_t n_buckets, const unsigned long long* __restrict__ sums, Off* __restrict__ off) {
  // one thread per 4 buckets inside the block's 1024: serial inside the thread, a scan across the 256 threads
  __shared__ unsigned long long part[256];
  const uint64_t first = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * 4;
  unsigned long long c[4], s = 0;
  for (int q = 0; q < 4; ++q) { c[q] = first + q < n_buckets ? index_bucket_capacity(n[first + q]) : 0; s += c[q]; }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    const unsigned long long add = (int)threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  unsigned long long run = sums[blockIdx.x] + part[threadIdx.x] - s;
  for (int q = 0; q < 4; ++q) { if (first + q < n_buckets) off[first + q] = (Off)run; run += c[q]; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) off[n_buckets] = (Off)run;
}

}  // namespace

void free_scan_index(jb_db* db) {
  if (db->idx_slots) cudaFree(db->idx_slots);
  if (db->idx_off) cudaFree(db->idx_off);
  db->idx_slots = nullptr; db->idx_off = nullptr; db->idx_bits = 0; db->idx_n_slots = 0; db->idx_wide_off = false;
}

// JB_OK also when the index is not built (switched off, k too small, partitioned table, not enough memory beside the table)
int build_scan_index(jb_db* db) {
  db->idx_bits = 0;
  const int k = db->hdr.k();
  // Off unless asked for (JB_SCAN_INDEX=1, or 2 = build it even when little memory is left): measured on the bench workload
  // (profiles/r2j_scan_index_ab.txt) the index scan takes 3.2 ms against the hash table's 1.88 ms.  It does cut DRAM traffic (5.9 GB
  // against 8.2 GB per launch) but not to the ~3 GB the layout promises -- a bucket mixes several minimizers and a run's windows
  // land on 4-5 lines of it -- and it pays for that with 1.5x the instructions and twice the exposed load latency (bucket bounds,
  // then the slot).  Kept as an option with its tests: the records and the counts are the table's either way.
  const char* env = getenv("JB_SCAN_INDEX");
  if (!env || atoi(env) == 0) return JB_OK;
  if (db->n_parts > 1 || k < 15 || db->n_records == 0) return JB_OK;
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  const auto t0 = std::chrono::steady_clock::now();
  const uint64_t N = db->n_records;
  int bits = 8;
  while (bits < 28 && (N >> (bits + 1)) >= 12) ++bits;            // ~12-24 records per bucket
  const uint64_t n_buckets = 1ull << bits;
  const uint64_t slots_bound = N + N / 4 + 2 * n_buckets + 64;    // sum of (n + n/4 + 2) over the non-empty buckets
  const bool wide_off = slots_bound >= (1ull << 32);
  size_t free_b = 0, total_b = 0;
  JB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t need = slots_bound * 16 + (n_buckets + 1) * (wide_off ? 8 : 4) + n_buckets * 4 + (64ull << 20);
  // keep room for the polishing work buffers: a fifth of the device, unless the caller insists (JB_SCAN_INDEX=2)
  const uint64_t reserve = env && atoi(env) >= 2 ? (2ull << 30) : std::max<uint64_t>(8ull << 30, total_b / 5);
  if (need + reserve > free_b) { note_counter("index_bytes", 0); return JB_OK; }

  unsigned int* d_n = nullptr;
  unsigned long long* d_sums = nullptr;
  uint64_t* d_inv = nullptr;
  unsigned long long* d_misc = nullptr;                           // [0] total slots, [1] undecodable compact slots
  const uint64_t n_blocks = (n_buckets + kScanBlock - 1) / kScanBlock;
  auto cleanup = [&]() { cudaFree(d_n); cudaFree(d_sums); cudaFree(d_inv); cudaFree(d_misc); };
  auto bail = [&](int rc, const char* what) { cleanup(); free_scan_index(db); return rc == JB_OK ? JB_OK : fail(rc, what); };
  if (cudaMalloc(&d_n, n_buckets * 4) != cudaSuccess || cudaMalloc(&d_sums, n_blocks * 8) != cudaSuccess ||
      cudaMalloc(&d_misc, 16) != cudaSuccess || cudaMalloc(&db->idx_off, (n_buckets + 1) * (wide_off ? 8 : 4)) != cudaSuccess) {
    cudaGetLastError();
    return bail(JB_OK, "");
  }
  JB_CUDA(cudaMemsetAsync(d_n, 0, n_buckets * 4, s));
  JB_CUDA(cudaMemsetAsync(d_misc, 0, 16, s));
  IndexBuild b{};
  b.ix.bucket_bits = bits; b.ix.m = index_mmer_len(k); b.bucket_n = d_n; b.k = k;
  b.ix.off32 = wide_off ? nullptr : reinterpret_cast<const uint32_t*>(db->idx_off);
  b.ix.off64 = wide_off ? reinterpret_cast<const uint64_t*>(db->idx_off) : nullptr;
  const bool compact = db->geom.layout == kLayoutCompact;
  CompactDecode dec{};
  if (compact) {
    const std::vector<uint64_t> inv = low_block_inverse(db->hdr, db->geom.r_bits);
    if (inv.empty()) return bail(JB_OK, "");
    if (cudaMalloc(&d_inv, std::max<size_t>(inv.size(), 1) * 8) != cudaSuccess) { cudaGetLastError(); return bail(JB_OK, ""); }
    JB_CUDA(cudaMemcpyAsync(d_inv, inv.data(), inv.size() * 8, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaStreamSynchronize(s));                            // (inv is a local)
    dec.g = db->geom; dec.lut = db->lut; dec.inv = d_inv; dec.n_units = db->n_units;
  }
  const unsigned grid = 148 * 8;
  auto sweep = [&](int phase) -> int {
    if (compact) {
      k_index_compact<<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(db->mem), db->n_units * kCompactSlots, dec, b, phase, d_misc + 1);
      if (db->n_stash) k_index_wide<<<148, 256, 0, s>>>(db->stash_list, db->n_stash, b, phase);
    } else {
      k_index_wide<<<grid, 256, 0, s>>>(reinterpret_cast<const Slot*>(db->mem), db->n_units * 2, b, phase);
    }
    JB_CUDA(cudaGetLastError());
    return JB_OK;
  };
  if (int rc = sweep(0)) return bail(rc, "scan index: first sweep");
  k_cap_block_sums<<<(unsigned)n_blocks, 256, 0, s>>>(d_n, n_buckets, d_sums);
  k_cap_scan_sums<<<1, 1024, 0, s>>>(d_sums, n_blocks, d_misc);
  if (wide_off) k_cap_offsets<uint64_t><<<(unsigned)n_blocks, 256, 0, s>>>(d_n, n_buckets, d_sums, reinterpret_cast<uint64_t*>(db->idx_off));
  else k_cap_offsets<uint32_t><<<(unsigned)n_blocks, 256, 0, s>>>(d_n, n_buckets, d_sums, reinterpret_cast<uint32_t*>(db->idx_off));
  JB_CUDA(cudaGetLastError());
  unsigned long long misc[2] = {0, 0};
  JB_CUDA(cudaMemcpyAsync(misc, d_misc, 16, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  if (misc[1] != 0 || misc[0] > slots_bound) return bail(JB_OK, "");           // (a table this code cannot read back: no index, the hash table serves)
  if (cudaMalloc(&db->idx_slots, std::max<uint64_t>(misc[0], 1) * 16) != cudaSuccess) { cudaGetLastError(); return bail(JB_OK, ""); }
  JB_CUDA(cudaMemsetAsync(db->idx_slots, 0, std::max<uint64_t>(misc[0], 1) * 16, s));
  b.ix.slots = reinterpret_cast<const uint4*>(db->idx_slots);
  if (int rc = sweep(1)) return bail(rc, "scan index: second sweep");
  JB_CUDA(cudaStreamSynchronize(s));
  cleanup();
  db->idx_bits = bits; db->idx_m = b.ix.m; db->idx_n_slots = misc[0]; db->idx_wide_off = wide_off;
  db->idx_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  note_counter("index_bytes", (int64_t)(misc[0] * 16 + (n_buckets + 1) * (wide_off ? 8 : 4)));
  note_counter("index_buckets", (int64_t)n_buckets);
  note_timing("index_build", db->idx_seconds * 1e3);
  return JB_OK;
}

}  // namespace jb

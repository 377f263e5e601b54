// jb_table.cu -- .jf loader: streams the "binary/sorted" records of a Jellyfish database into a
// GPU-resident open-addressing table (replaces QueryMerFile's ctor: JF!swig/mer_file.i:18-36,
// JF!include/jellyfish/binary_dumper.hpp:126-146, JF!include/jellyfish/mapped_file.hpp:30-50).
//
// The file is sorted by (pos, key), pos = M*key & (size-1) (JF!include/jellyfish/sorted_dumper.hpp:72-103,
// JF!include/jellyfish/mer_heap.hpp:26-30).  home(key) is monotone in that order, so linear-probing
// placement needs neither atomics nor re-hashing of occupied slots:
//     slot_i = max(home_i, slot_{i-1} + 1) = i + max_{j<=i}(home_j - j)
// i.e. one prefix-max scan over the record stream (K0a unpack+hash, K0b block carry, K0c place).
// Layouts (jb_kmer.cuh): compact = four 8-byte slots per 32-byte sector, wide = two 16-byte slots; home = first slot of
// the key's sector.  Reader threads fill a ring of pinned buffers (pread, optionally
// O_DIRECT) while the submitting thread queues the copy and the three kernels of the previous chunk.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "jb_geom.hpp"
#include "jb_internal.cuh"

namespace jb {
namespace {

constexpr int kRecPerBlock = 1024;   // 256 threads x 4 records
constexpr int kThreads = 256;

struct BuildState {          // device-resident carry between chunks
  long long carry;           // max_j (home_j - j) so far
  unsigned long long prev_posx;       // posx of the last record of the previous chunk
  unsigned long long next_prev_posx;  // written by k_place, promoted by the next k_block_carry
  unsigned long long max_disp;
  unsigned long long last_slot;
  unsigned long long n_stash;         // compact: records appended to the stash list
  int err_unsorted;
  int err_overflow;
  int err_zero;                       // a record whose 32-bit count is 0 cannot be told from an empty slot
  int err_stash;                      // the stash list is full
};

// K0a: raw records -> (key, count, posx); per-block max of (home slot - global index).
__global__ void __launch_bounds__(kThreads)
k_unpack(const uint8_t* __restrict__ raw, uint64_t n, uint64_t base_index, int rec_bytes, int key_bytes,
         int cnt_bytes, TableView t, int spu_shift, uint64_t* __restrict__ key_lo, uint64_t* __restrict__ key_hicnt,
         uint64_t* __restrict__ posx_out, long long* __restrict__ block_max) {
  extern __shared__ uint8_t s_raw[];
  __shared__ long long s_warp[kThreads / 32];
  const uint64_t first = (uint64_t)blockIdx.x * kRecPerBlock;
  const uint64_t cnt = min((uint64_t)kRecPerBlock, n - first);
  const uint64_t nbytes = cnt * rec_bytes;
  const uint8_t* src = raw + first * rec_bytes;           // block offset is a multiple of 1024 bytes
  for (uint64_t w = threadIdx.x; w * 4 < nbytes; w += kThreads) {
    uint32_t v;
    if (w * 4 + 4 <= nbytes) v = reinterpret_cast<const uint32_t*>(src)[w];
    else { v = 0; for (uint64_t b = w * 4; b < nbytes; ++b) v |= (uint32_t)src[b] << (8 * (b - w * 4)); }
    reinterpret_cast<uint32_t*>(s_raw)[w] = v;
  }
  __syncthreads();
  long long tmax = LLONG_MIN;
  for (int r = threadIdx.x; r < (int)cnt; r += kThreads) {
    const uint8_t* p = s_raw + (size_t)r * rec_bytes;
    Key key = {0, 0};
    for (int b = 0; b < key_bytes; ++b) {
      uint64_t v = p[b];
      if (b < 8) key.lo |= v << (8 * b); else key.hi |= v << (8 * (b - 8));
    }
    key.lo &= t.mask.lo; key.hi &= t.mask.hi;                        // clean_msw (binary_dumper.hpp:204)
    uint32_t c = 0;
    for (int b = 0; b < cnt_bytes && b < 4; ++b) c |= (uint32_t)p[key_bytes + b] << (8 * b);   // (unsigned int) cast, mer_file.i:41
    const uint64_t px = key_posx(t.g, hash_lut(t.lut, key_bytes, key), key);
    const uint64_t h = home_unit_of_posx(t.g, px) << spu_shift;
    uint64_t gi = first + r;
    key_lo[gi] = key.lo;
    key_hicnt[gi] = (key.hi & 0xFFFFFFFFull) | ((uint64_t)c << 32);
    posx_out[gi] = px;
    long long tv = (long long)h - (long long)(base_index + gi);
    tmax = tv > tmax ? tv : tmax;
  }
  for (int o = 16; o; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, tmax, o); tmax = v > tmax ? v : tmax; }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = tmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = s_warp[0];
    for (int w = 1; w < kThreads / 32; ++w) m = s_warp[w] > m ? s_warp[w] : m;
    block_max[blockIdx.x] = m;
  }
}

// K0b: exclusive prefix-max over the block aggregates, seeded with the carry of earlier chunks.
__global__ void k_block_carry(long long* block_max, int nblocks, BuildState* st) {
  long long run = st->carry;                  // one warp, 32 entries per step
  const int lane = threadIdx.x;
  if (lane == 0) st->prev_posx = st->next_prev_posx;
  for (int b0 = 0; b0 < nblocks; b0 += 32) {
    int b = b0 + lane;
    long long v = b < nblocks ? block_max[b] : LLONG_MIN;
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o && u > inc) inc = u; }
    long long exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = LLONG_MIN;
    if (b < nblocks) block_max[b] = exc > run ? exc : run;
    long long tot = __shfl_sync(0xffffffffu, inc, 31);
    run = tot > run ? tot : run;
  }
  if (lane == 0) st->carry = run;
}

// Where k_place writes: one array, or the parts of a partitioned table (peer stores for the foreign ones; a slot inside
// the first `overlap` units of part p is also written behind the end of part p-1).
struct PlaceTarget {
  uint8_t* base;
  int n_parts;
  uint64_t overlap;                    // units
  uint8_t* part[kMaxParts];
  uint64_t part_first[kMaxParts + 1];  // units
  int spu_shift;                       // log2(slots per unit): wide 1, compact 2
  int slot_shift;                      // log2(slot bytes): wide 4, compact 3
  Slot* stash_list;                    // compact
  unsigned long long stash_cap;
};
// the addresses slot `slot` lives at (one, or two inside an overlap); returns how many
__device__ __forceinline__ int slot_addrs(const PlaceTarget& T, uint64_t slot, uint8_t* out[2]) {
  if (T.n_parts <= 1) { out[0] = T.base + (slot << T.slot_shift); return 1; }
  const uint64_t unit = slot >> T.spu_shift;
  int p = 0;
#pragma unroll 1
  for (int q = 1; q < T.n_parts; ++q) p += unit >= T.part_first[q] ? 1 : 0;
  int n = 0;
  if (T.part[p]) out[n++] = T.part[p] + ((slot - (T.part_first[p] << T.spu_shift)) << T.slot_shift);     // (a rank of jb_db_open_part keeps its own part only)
  if (p > 0 && unit < T.part_first[p] + T.overlap && T.part[p - 1]) out[n++] = T.part[p - 1] + ((slot - (T.part_first[p - 1] << T.spu_shift)) << T.slot_shift);
  return n;
}

// K0c: inclusive prefix-max inside the block, then scatter the slots.
__global__ void __launch_bounds__(kThreads)
k_place(uint64_t n, uint64_t base_index, const uint64_t* __restrict__ key_lo, const uint64_t* __restrict__ key_hicnt,
        const uint64_t* __restrict__ posx_in, const long long* __restrict__ block_carry, TableGeom g, PlaceTarget target,
        uint64_t n_slots, BuildState* st) {
  __shared__ long long s_warp[kThreads / 32];
  const uint64_t first = (uint64_t)blockIdx.x * kRecPerBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread owns 4 consecutive records so the scan is a plain blocked scan
  long long t[4];
  uint64_t h[4], px[4];
  long long run = LLONG_MIN;
  bool unsorted = false;
  for (int q = 0; q < 4; ++q) {
    uint64_t gi = first + (uint64_t)threadIdx.x * 4 + q;
    if (gi < n) {
      px[q] = posx_in[gi];
      h[q] = home_unit_of_posx(g, px[q]) << target.spu_shift;
      uint64_t prev = gi > 0 ? posx_in[gi - 1] : (base_index ? st->prev_posx : 0);
      unsorted |= px[q] < prev;
      t[q] = (long long)h[q] - (long long)(base_index + gi);
    } else { h[q] = 0; px[q] = 0; t[q] = LLONG_MIN; }
    run = t[q] > run ? t[q] : run;
    t[q] = run;
  }
  long long inc = run;
  for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o && u > inc) inc = u; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  long long pre = block_carry[blockIdx.x];
  for (int w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
  long long exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane > 0 && exc > pre) pre = exc;
  unsigned long long disp = 0, last = 0;
  bool overflow = false;
  const uint32_t cmask = g.layout == kLayoutCompact ? cval_mask(g) : 0u;
  for (int q = 0; q < 4; ++q) {
    uint64_t gi = first + (uint64_t)threadIdx.x * 4 + q;
    if (gi >= n) break;
    long long m = t[q] > pre ? t[q] : pre;
    uint64_t slot = (uint64_t)((long long)(base_index + gi) + m);
    if (slot + 2 >= n_slots) { overflow = true; continue; }
    const uint64_t hc = key_hicnt[gi];
    const uint32_t c = (uint32_t)(hc >> 32);
    if (c == 0) st->err_zero = 1;
    uint8_t* at[2];
    const int na = slot_addrs(target, slot, at);
    if (g.layout == kLayoutCompact) {
      const Key key = {key_lo[gi], hc & 0xFFFFFFFFull};
      const uint64_t line = slot >> 2, home_line = h[q] >> 2;              // sectors
      uint32_t cv = c;
      bool to_stash = false;
      if (line - home_line > (uint64_t)kMaxHops) { cv = cmask - 1u; to_stash = true; }        // tombstone
      else if (c > cmask - 2u) { cv = cmask; to_stash = true; }                                    // escape
      if (to_stash) {
        const unsigned long long idx = atomicAdd(&st->n_stash, 1ull);
        if (idx < target.stash_cap) { Slot r; r.lo = key.lo; r.hi = (uint32_t)key.hi; r.count = c; target.stash_list[idx] = r; }
        else st->err_stash = 1;
      }
      const unsigned long long word = compact_pattern(g, px[q], key) | cv;
      // the last slot of a sector may also receive the "run continues" flag from the thread that places its successor
      for (int a = 0; a < na; ++a) {
        if ((slot & 3) == 3) atomicOr(reinterpret_cast<unsigned long long*>(at[a]), word);
        else *reinterpret_cast<unsigned long long*>(at[a]) = word;
      }
      if ((slot & 3) == 0 && home_line < line) {
        uint8_t* bt[2];
        const int nb2 = slot_addrs(target, slot - 1, bt);
        for (int a = 0; a < nb2; ++a) atomicOr(reinterpret_cast<unsigned long long*>(bt[a]), (unsigned long long)spill_bit(g));
      }
    } else {
      Slot s2; s2.lo = key_lo[gi]; s2.hi = (uint32_t)hc; s2.count = c;
      for (int a = 0; a < na; ++a) *reinterpret_cast<Slot*>(at[a]) = s2;
    }
    disp = max(disp, (unsigned long long)(slot - h[q]));
    last = slot;
  }
  if (disp) atomicMax(&st->max_disp, disp);
  if (last) atomicMax(&st->last_slot, last);
  if (unsorted) st->err_unsorted = 1;
  if (overflow) st->err_overflow = 1;
  for (int q = 0; q < 4; ++q)
    if (first + (uint64_t)threadIdx.x * 4 + q == n - 1) st->next_prev_posx = px[q];
}

// the stash of a compact table: the listed records into a small wide table (claim a slot by its count word)
__global__ void k_stash_insert(const Slot* __restrict__ list, uint64_t n, Slot* table, uint64_t units, uint64_t n_slots, TableView t, int* err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Slot r = list[i];
  const Key key = {r.lo, r.hi};
  const uint64_t px = key_posx(t.g, hash_lut(t.lut, t.g.key_bytes, key), key);
  uint64_t s = 2 * mulhi64(px << t.g.left_shift, units);
  while (true) {
    if (s + 2 >= n_slots) { *err = 1; return; }
    if (atomicCAS(&table[s].count, 0u, r.count) == 0u) { table[s].lo = r.lo; table[s].hi = r.hi; return; }
    ++s;
  }
}

__global__ void k_dump(const Slot* __restrict__ slots, uint64_t n_slots, uint64_t* key_lo, uint64_t* key_hi,
                       uint32_t* count, unsigned long long* cursor_by_block, int phase) {
  // phase 0: per-block occupied count; phase 1: write using exclusive offsets in cursor_by_block
  __shared__ unsigned int s_cnt;
  __shared__ unsigned int s_warp_off[kThreads / 32];
  const uint64_t first = (uint64_t)blockIdx.x * kThreads;
  uint64_t i = first + threadIdx.x;
  bool occ = i < n_slots && slots[i].count != 0;
  unsigned int ball = __ballot_sync(0xffffffffu, occ);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp_off[warp] = __popc(ball);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int w = 0; w < kThreads / 32; ++w) { unsigned int c = s_warp_off[w]; s_warp_off[w] = run; run += c; }
    s_cnt = run;
  }
  __syncthreads();
  if (phase == 0) { if (threadIdx.x == 0) cursor_by_block[blockIdx.x] = s_cnt; return; }
  if (occ) {
    uint64_t o = cursor_by_block[blockIdx.x] + s_warp_off[warp] + __popc(ball & ((1u << lane) - 1));
    Slot s = slots[i];
    key_lo[o] = s.lo; key_hi[o] = s.hi; count[o] = s.count;
  }
}

// wide slots (a wide table, or the stash LIST of a compact one: every entry occupied)
__global__ void k_histogram(const Slot* __restrict__ slots, uint64_t n_slots, unsigned long long high,
                            unsigned long long* hist) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = slots[i].count;
    if (c) atomicAdd(&hist[c > high ? high + 1 : c], 1ull);
  }
}
// compact slots: direct counts only (escapes and tombstones are counted from the stash list)
__global__ void k_histogram_compact(const unsigned long long* __restrict__ slots, uint64_t n_slots, uint32_t cmask, unsigned long long high,
                                    unsigned long long* hist) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long w = slots[i];
    const uint32_t c = (uint32_t)w & cmask;
    if (w != 0 && c < cmask - 1u) atomicAdd(&hist[c > high ? high + 1 : c], 1ull);
  }
}

}  // namespace
}  // namespace jb

using namespace jb;

namespace {

// ---- reading the file: a few threads fill a ring of pinned buffers ------------------------------------------------
struct ChunkRing {
  int fd = -1;
  bool odirect = false;
  uint64_t data_offset = 0, rec = 0, n = 0, chunk_rec = 0, n_chunks = 0;
  int n_buf = 0;
  std::vector<uint8_t*> pinned;
  std::vector<uint32_t> lead;               // bytes before the chunk's first record in its buffer (O_DIRECT alignment)
  std::vector<cudaEvent_t> copied;          // H2D out of buffer b has finished
  std::mutex mu;
  std::condition_variable cv;
  std::vector<int64_t> filled;              // chunk held by buffer b, or -1
  int64_t submitted = 0;                    // chunks whose H2D has been queued (their `copied` event is recorded)
  bool io_error = false, stop = false;
  std::vector<double> busy;                 // per reader: seconds inside pread
  std::vector<std::thread> threads;

  void reader(int id, int n_readers, int device) {
    cudaSetDevice(device);
    for (uint64_t c = (uint64_t)id; c < n_chunks; c += (uint64_t)n_readers) {
      const int b = (int)(c % (uint64_t)n_buf);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || c < (uint64_t)n_buf || submitted > (int64_t)(c - (uint64_t)n_buf); });
        if (stop) return;
      }
      if (c >= (uint64_t)n_buf) cudaEventSynchronize(copied[(size_t)b]);      // the copy that last used this buffer is done
      const uint64_t m = std::min(chunk_rec, n - c * chunk_rec);
      const uint64_t off = data_offset + c * chunk_rec * rec, want = m * rec;
      const uint64_t a_off = odirect ? off & ~4095ull : off;
      const uint64_t ld = off - a_off;
      const uint64_t a_want = odirect ? (ld + want + 4095) & ~4095ull : want;
      auto t0 = std::chrono::steady_clock::now();
      uint64_t got = 0;
      bool bad = false;
      while (got < ld + want) {
        ssize_t rd = pread(fd, pinned[(size_t)b] + got, a_want - got, (off_t)(a_off + got));
        if (rd <= 0) { bad = true; break; }
        got += (uint64_t)rd;
      }
      busy[(size_t)id] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (bad) io_error = true;
        lead[(size_t)b] = (uint32_t)ld;
        filled[(size_t)b] = (int64_t)c;
      }
      cv.notify_all();
      if (bad) return;
    }
  }
  // blocks until chunk c sits in its buffer; false on a read error
  bool wait_chunk(uint64_t c) {
    const int b = (int)(c % (uint64_t)n_buf);
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return io_error || filled[(size_t)b] == (int64_t)c; });
    return !io_error;
  }
  void mark_submitted() {
    { std::lock_guard<std::mutex> lk(mu); ++submitted; }
    cv.notify_all();
  }
  void shutdown() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    for (auto& t : threads) if (t.joinable()) t.join();
    threads.clear();
  }
};

// everything open_impl allocates besides the table itself; freed on every exit path
struct LoaderScratch {
  ChunkRing ring;
  std::vector<uint8_t*> d_raw;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, kdone[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  uint64_t *d_lo = nullptr, *d_hc = nullptr, *d_px = nullptr;
  long long* d_bmax = nullptr;
  BuildState* d_st = nullptr;
  int* d_err = nullptr;
  std::vector<jb_db*> dbs;
  bool keep = false;                        // success: the handles go to the caller
  ~LoaderScratch() {
    ring.shutdown();
    if (ring.fd >= 0) ::close(ring.fd);
    for (uint8_t* p : ring.pinned) if (p) cudaFreeHost(p);
    for (cudaEvent_t e : ring.copied) if (e) cudaEventDestroy(e);
    for (uint8_t* p : d_raw) if (p) cudaFree(p);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (cudaEvent_t e : kdone) if (e) cudaEventDestroy(e);
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    cudaFree(d_lo); cudaFree(d_hc); cudaFree(d_px); cudaFree(d_bmax); cudaFree(d_st); cudaFree(d_err);
    if (!keep) for (jb_db* d : dbs) jb_db_close(d);
    cudaGetLastError();
  }
};

constexpr int kStashOverflow = -1000;       // internal: the compact layout's stash list was too small -> retry wide

// One database, n >= 1 handles.  n == 1: the whole table on devices[0].  n > 1: the units partitioned over the devices
// (SURVEY.md §8e fallback): the records are unpacked and placed by kernels on devices[0], whose stores go straight into
// the owning part (peer stores), so the table never has to fit one GPU; every handle then probes all parts.
// only_part >= 0 (jb_db_open_part, one process per GPU): n_dev parts in all, this process builds and keeps part only_part
// on devices[0]; the other parts are attached later through CUDA IPC handles.
int open_impl(const char* path, int n_dev, const int* devices, jb_db** outs, int only_part, int layout_pref) {
  auto t0 = std::chrono::steady_clock::now();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(JB_E_CUDA, "no CUDA device: libjasper_b200 has no CPU path");
  if (n_dev < 1 || n_dev > kMaxParts) return fail(JB_E_ARG, "jb_db_open: between 1 and " + std::to_string(kMaxParts) + " devices");
  const int n_local = only_part >= 0 ? 1 : n_dev;       // handles (and devices) of this process
  for (int i = 0; i < n_local; ++i) {
    if (devices[i] < 0 || devices[i] >= ndev) return fail(JB_E_ARG, "jb_db_open: bad device index");
    // (tests on a one-GPU box: JB_PARTS_SAME_DEVICE lets the parts share a device -- separate allocations, the same partition
    //  arithmetic and probe routing, "peer" loads that happen to be local)
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i] && getenv("JB_PARTS_SAME_DEVICE") == nullptr) return fail(JB_E_ARG, "jb_db_open_partitioned: a device is listed twice");
  }
  const int device = devices[0];
  if (n_dev > 1 && only_part < 0) {
    for (int i = 0; i < n_dev; ++i)
      for (int j = 0; j < n_dev; ++j) {
        if (i == j || devices[i] == devices[j]) continue;
        int can = 0;
        JB_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
        if (!can) return fail(JB_E_UNSUPPORTED, "jb_db_open_partitioned: devices " + std::to_string(devices[i]) + " and " + std::to_string(devices[j]) + " have no peer access");
        JB_CUDA(cudaSetDevice(devices[i]));
        cudaError_t pe = cudaDeviceEnablePeerAccess(devices[j], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(JB_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
        cudaGetLastError();
      }
  }
  JB_CUDA(cudaSetDevice(device));

  LoaderScratch L;
  ChunkRing& ring = L.ring;
  ring.odirect = getenv("JB_ODIRECT") && atoi(getenv("JB_ODIRECT")) != 0;
  ring.fd = ::open(path, O_RDONLY);
  if (ring.fd < 0) return fail(JB_E_IO, std::string("Can't open file '") + path + "'");
  struct stat st;
  fstat(ring.fd, &st);
  const uint64_t fsize = (uint64_t)st.st_size;
  std::vector<char> head(9);
  if (fsize < 9 || pread(ring.fd, head.data(), 9, 0) != 9) return fail(JB_E_FORMAT, "file too short for a Jellyfish header");
  uint64_t hlen = 0;
  for (int i = 0; i < 9; ++i) hlen = hlen * 10 + (uint64_t)((head[i] >= '0' && head[i] <= '9') ? head[i] - '0' : 0);
  head.resize(9 + std::min<uint64_t>(hlen, fsize - 9));
  if ((uint64_t)pread(ring.fd, head.data(), head.size(), 0) != head.size()) return fail(JB_E_IO, "short read of header");
  L.dbs.assign((size_t)n_local, nullptr);
  for (int i = 0; i < n_local; ++i) L.dbs[i] = new jb_db();
  std::vector<jb_db*>& dbs = L.dbs;
  jb_db* db = dbs[0];
  std::string err = parse_jf_header(head.data(), head.size(), &db->hdr);
  if (!err.empty()) return fail(JB_E_FORMAT, err);
  const JfHeader& h = db->hdr;
  if (h.k() > JB_MAX_K || h.k() < 1) return fail(JB_E_UNSUPPORTED, "k = " + std::to_string(h.k()) + " outside [1, " + std::to_string(JB_MAX_K) + "]");
  const uint64_t body = fsize - h.data_offset;
  const int rec = h.record_bytes();
  if (body % rec) return fail(JB_E_FORMAT, "Size of database (" + std::to_string(body) + ") must be a multiple of the length of a record (" + std::to_string(rec) + ")");
  const uint64_t n = body / rec;
  if (ring.odirect) {                                     // reopen unbuffered; fall back silently where the file system refuses
    int fd2 = ::open(path, O_RDONLY | O_DIRECT);
    if (fd2 >= 0) { ::close(ring.fd); ring.fd = fd2; } else ring.odirect = false;
  }

  // ---- geometry -------------------------------------------------------------------------------
  double load = 0;
  if (const char* e = getenv("JB_TABLE_LOAD")) { double v = atof(e); if (v > 0.05 && v <= 0.95) load = v; }
  if (const char* e = getenv("JB_TABLE_LAYOUT")) {
    if (layout_pref < 0 && !strcmp(e, "wide")) layout_pref = kLayoutWide;
    if (layout_pref < 0 && !strcmp(e, "compact")) layout_pref = kLayoutCompact;
  }
  size_t free_b = ~(size_t)0, total_b = 0;
  for (int i = 0; i < n_local; ++i) {                     // the tightest device decides
    size_t f = 0;
    JB_CUDA(cudaSetDevice(devices[i]));
    JB_CUDA(cudaMemGetInfo(&f, &total_b));
    free_b = std::min(free_b, f);
  }
  JB_CUDA(cudaSetDevice(device));
  GeomChoice gc;
  uint64_t stash_cap = 0, overlap = 0;
  for (int attempt = 0; ; ++attempt) {
    gc = choose_geometry(h, n, load, layout_pref);
    const bool compact = gc.g.layout == kLayoutCompact;
    overlap = n_dev > 1 ? (compact ? kPartOverlapCompact : kPartOverlapWide) : 0;
    double frac = 1.0 / 64;
    if (const char* e = getenv("JB_STASH_FRAC")) { double v = atof(e); if (v > 0 && v <= 1) frac = v; }
    stash_cap = compact ? std::max<uint64_t>(4096, (uint64_t)((double)n * frac)) : 0;
    const uint64_t need = (gc.n_units / (uint64_t)n_dev + overlap + gc.tail_units) * (uint64_t)unit_bytes_of(gc.g) + stash_cap * 48 + (512ull << 20);
    if (need < free_b) break;
    if (gc.load >= 0.9) return fail(JB_E_NOMEM, "k-mer table (" + std::to_string(n) + " records) does not fit in device memory");
    load = std::min(0.9, gc.load + 0.05);
  }
  if (const char* e = getenv("JB_COMPACT_CVAL_BITS"))      // test knob: a narrow count field sends most records through the stash
    if (gc.g.layout == kLayoutCompact && atoi(e) >= 2) gc.g.cnt_bits = std::min(gc.g.cnt_bits, atoi(e) + 1);
  const TableGeom g = gc.g;
  const bool compact = g.layout == kLayoutCompact;
  const uint64_t n_units = gc.n_units;
  const int spu_shift = compact ? 2 : 1, slot_shift = compact ? 3 : 4, ub = unit_bytes_of(g);
  const uint64_t n_slots = n_units << spu_shift;

  // parts: equal shares of the home range, the last one also takes the overflow tail
  uint64_t part_first[kMaxParts + 1] = {};
  for (int p2 = 0; p2 <= n_dev; ++p2) part_first[p2] = p2 == n_dev ? n_units : g.units / (uint64_t)n_dev * (uint64_t)p2;
  for (int p2 = 0; p2 + 1 < n_dev; ++p2) overlap = std::min<uint64_t>(overlap, part_first[p2 + 2] - part_first[p2 + 1]);
  std::vector<uint64_t> lut = build_hash_lut(h);
  uint8_t* parts[kMaxParts] = {};
  for (int li = 0; li < n_local; ++li) {
    const int i = only_part >= 0 ? only_part : li;        // which part this handle owns
    jb_db* d = dbs[li];
    if (li) d->hdr = db->hdr;
    d->n_records = n; d->device = devices[li]; d->geom = g; d->home_units = g.units; d->n_units = n_units;
    JB_CUDA(cudaSetDevice(devices[li]));
    JB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    const uint64_t mine = part_first[i + 1] - part_first[i] + (i + 1 < n_dev ? overlap : 0);
    if (cudaMalloc(&d->mem, mine * (uint64_t)ub) != cudaSuccess) { cudaGetLastError(); return fail(JB_E_NOMEM, "cannot allocate the k-mer table part on device " + std::to_string(devices[li])); }
    JB_CUDA(cudaMemsetAsync(d->mem, 0, mine * (uint64_t)ub, d->stream));
    JB_CUDA(cudaMalloc(&d->lut, lut.size() * 8));
    JB_CUDA(cudaMemcpyAsync(d->lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, d->stream));
    JB_CUDA(cudaStreamSynchronize(d->stream));
    parts[i] = d->mem;
  }
  for (int li = 0; li < n_local; ++li) {
    jb_db* d = dbs[li];
    d->n_parts = n_dev; d->my_part = only_part >= 0 ? only_part : li;
    d->parts_ready = only_part < 0 || n_dev == 1;
    d->part_overlap = overlap;
    for (int q = 0; q < n_dev; ++q) d->parts[q] = parts[q];
    for (int q = 0; q <= n_dev; ++q) d->part_first[q] = part_first[q];
  }
  JB_CUDA(cudaSetDevice(device));
  cudaStream_t s = db->stream;
  if (compact) JB_CUDA(cudaMalloc(&db->stash_list, stash_cap * sizeof(Slot)));
  PlaceTarget target{};
  target.base = db->mem; target.n_parts = n_dev; target.overlap = overlap;
  target.spu_shift = spu_shift; target.slot_shift = slot_shift;
  target.stash_list = db->stash_list; target.stash_cap = stash_cap;
  for (int q = 0; q < n_dev; ++q) target.part[q] = parts[q];
  for (int q = 0; q <= n_dev; ++q) target.part_first[q] = part_first[q];

  // ---- stream the records ---------------------------------------------------------------------
  const uint64_t chunk_rec = std::min<uint64_t>(std::max<uint64_t>(n, 1), 4ull << 20);
  const uint64_t chunk_bytes = chunk_rec * rec;
  const uint64_t n_chunks = (n + chunk_rec - 1) / chunk_rec;
  int n_readers = 4;
  if (const char* e = getenv("JB_READERS")) n_readers = std::max(1, std::min(16, atoi(e)));
  n_readers = (int)std::min<uint64_t>((uint64_t)n_readers, std::max<uint64_t>(n_chunks, 1));
  ring.data_offset = h.data_offset; ring.rec = (uint64_t)rec; ring.n = n; ring.chunk_rec = chunk_rec; ring.n_chunks = n_chunks;
  ring.n_buf = (int)std::min<uint64_t>(std::max<uint64_t>(n_chunks, 1), (uint64_t)n_readers + 2);
  ring.pinned.assign((size_t)ring.n_buf, nullptr); ring.copied.assign((size_t)ring.n_buf, nullptr);
  ring.lead.assign((size_t)ring.n_buf, 0); ring.filled.assign((size_t)ring.n_buf, -1); ring.busy.assign((size_t)n_readers, 0.0);
  L.d_raw.assign(2, nullptr);
  const int max_blocks = (int)((chunk_rec + kRecPerBlock - 1) / kRecPerBlock);
  for (int b = 0; b < ring.n_buf; ++b) {
    JB_CUDA(cudaMallocHost(&ring.pinned[(size_t)b], chunk_bytes + 8192));
    JB_CUDA(cudaEventCreateWithFlags(&ring.copied[(size_t)b], cudaEventDisableTiming));
  }
  for (int b = 0; b < 2; ++b) JB_CUDA(cudaMalloc(&L.d_raw[(size_t)b], chunk_bytes + 16));
  JB_CUDA(cudaEventCreate(&L.ev0)); JB_CUDA(cudaEventCreate(&L.ev1));
  JB_CUDA(cudaEventCreateWithFlags(&L.kdone[0], cudaEventDisableTiming)); JB_CUDA(cudaEventCreateWithFlags(&L.kdone[1], cudaEventDisableTiming));
  JB_CUDA(cudaStreamCreateWithFlags(&L.copy_stream, cudaStreamNonBlocking));
  JB_CUDA(cudaMalloc(&L.d_lo, chunk_rec * 8)); JB_CUDA(cudaMalloc(&L.d_hc, chunk_rec * 8)); JB_CUDA(cudaMalloc(&L.d_px, chunk_rec * 8));
  JB_CUDA(cudaMalloc(&L.d_bmax, (size_t)max_blocks * 8));
  JB_CUDA(cudaMalloc(&L.d_st, sizeof(BuildState)));
  JB_CUDA(cudaMalloc(&L.d_err, sizeof(int)));
  JB_CUDA(cudaMemsetAsync(L.d_err, 0, sizeof(int), s));
  BuildState init{};
  init.carry = LLONG_MIN;
  JB_CUDA(cudaMemcpyAsync(L.d_st, &init, sizeof init, cudaMemcpyHostToDevice, s));
  TableView view = db->view();
  JB_CUDA(cudaEventRecord(L.ev0, s));
  for (int t = 0; t < n_readers && n > 0; ++t) ring.threads.emplace_back(&ChunkRing::reader, &ring, t, n_readers, device);
  // two device staging buffers: the copy of chunk c (copy stream) overlaps the kernels of chunk c-1 (table stream)
  bool io_error = false;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    if (!ring.wait_chunk(c)) { io_error = true; break; }
    const uint64_t done = c * chunk_rec, m = std::min(chunk_rec, n - done);
    const int b = (int)(c % (uint64_t)ring.n_buf), rb = (int)(c & 1);
    if (c >= 2) JB_CUDA(cudaStreamWaitEvent(L.copy_stream, L.kdone[rb], 0));        // the kernels of chunk c-2 have read d_raw[rb]
    JB_CUDA(cudaMemcpyAsync(L.d_raw[(size_t)rb], ring.pinned[(size_t)b] + ring.lead[(size_t)b], m * rec, cudaMemcpyHostToDevice, L.copy_stream));
    JB_CUDA(cudaEventRecord(ring.copied[(size_t)b], L.copy_stream));
    ring.mark_submitted();
    JB_CUDA(cudaStreamWaitEvent(s, ring.copied[(size_t)b], 0));
    int nb = (int)((m + kRecPerBlock - 1) / kRecPerBlock);
    k_unpack<<<nb, kThreads, (size_t)kRecPerBlock * rec + 16, s>>>(L.d_raw[(size_t)rb], m, done, rec, h.key_bytes(), h.counter_len, view, spu_shift,
                                                                    L.d_lo, L.d_hc, L.d_px, L.d_bmax);
    k_block_carry<<<1, 32, 0, s>>>(L.d_bmax, nb, L.d_st);
    k_place<<<nb, kThreads, 0, s>>>(m, done, L.d_lo, L.d_hc, L.d_px, L.d_bmax, g, target, n_slots, L.d_st);
    JB_CUDA(cudaEventRecord(L.kdone[rb], s));
  }
  ring.shutdown();
  JB_CUDA(cudaEventRecord(L.ev1, s));
  BuildState fin{};
  JB_CUDA(cudaMemcpyAsync(&fin, L.d_st, sizeof fin, cudaMemcpyDeviceToHost, s));
  cudaError_t ce = cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, L.ev0, L.ev1);
  if (ce != cudaSuccess) return fail(JB_E_CUDA, std::string("table build: ") + cudaGetErrorString(ce));
  if (io_error || ring.io_error) return fail(JB_E_IO, "short read while streaming records");
  if (fin.err_unsorted) return fail(JB_E_FORMAT, "records are not sorted by (hash position, key): not a jellyfish binary/sorted file");
  if (fin.err_zero) return fail(JB_E_UNSUPPORTED, "database holds a record whose count is 0 modulo 2^32");
  if (fin.err_overflow) return fail(JB_E_NOMEM, "table overflow tail exhausted (hash positions are badly skewed); lower JB_TABLE_LOAD");
  if (fin.err_stash) { set_error("compact table: more than " + std::to_string(stash_cap) + " records need the stash (raise JB_STASH_FRAC or set JB_TABLE_LAYOUT=wide)"); return kStashOverflow; }
  // ---- the stash of a compact table: a wide mini-table on every device --------------------------
  if (compact && fin.n_stash > 0) {
    const uint64_t su = std::max<uint64_t>(64, fin.n_stash), sn = (su + std::max<uint64_t>(4096, su / 16) + 4) * 2;
    for (int li = 0; li < n_local; ++li) {
      jb_db* d = dbs[li];
      JB_CUDA(cudaSetDevice(d->device));
      JB_CUDA(cudaMalloc(&d->stash, sn * sizeof(Slot)));
      d->stash_units = su; d->n_stash = fin.n_stash;
      if (li == 0) {
        JB_CUDA(cudaMemsetAsync(d->stash, 0, sn * sizeof(Slot), s));
        k_stash_insert<<<(unsigned)((fin.n_stash + 255) / 256), 256, 0, s>>>(d->stash_list, fin.n_stash, d->stash, su, sn, view, L.d_err);
        int herr = 0;
        JB_CUDA(cudaMemcpyAsync(&herr, L.d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
        JB_CUDA(cudaStreamSynchronize(s));
        if (herr) return fail(JB_E_NOMEM, "stash table overflow (hash positions are badly skewed)");
      } else {
        JB_CUDA(cudaMemcpyPeer(d->stash, d->device, db->stash, db->device, sn * sizeof(Slot)));
        JB_CUDA(cudaMalloc(&d->stash_list, fin.n_stash * sizeof(Slot)));
        JB_CUDA(cudaMemcpyPeer(d->stash_list, d->device, db->stash_list, db->device, fin.n_stash * sizeof(Slot)));
      }
    }
    JB_CUDA(cudaSetDevice(device));
  }
  if (compact && fin.n_stash == 0) { cudaFree(db->stash_list); db->stash_list = nullptr; }
  // wide: a probe that starts in part p ends at the first empty slot: there must be one inside p's copy of the next part's start
  // (a compact probe reads at most kMaxHops + 1 <= overlap sectors by construction)
  for (int p2 = 0; p2 + 1 < n_dev && !compact; ++p2) {
    if (only_part >= 0 && p2 != only_part) continue;
    const uint64_t ov_slots = overlap * 2;
    std::vector<Slot> ov((size_t)ov_slots);
    JB_CUDA(cudaSetDevice(devices[only_part >= 0 ? 0 : p2]));
    JB_CUDA(cudaMemcpy(ov.data(), parts[p2] + (part_first[p2 + 1] - part_first[p2]) * 32, (size_t)ov_slots * sizeof(Slot), cudaMemcpyDeviceToHost));
    bool empty = false;
    for (uint64_t q = 0; q + 2 < ov_slots && !empty; ++q) empty = ov[q].count == 0;
    if (!empty) return fail(JB_E_NOMEM, "a run of occupied slots crosses a whole partition overlap (" + std::to_string(ov_slots) + " slots); lower JB_TABLE_LOAD");
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double busiest = 0;
  for (double b : ring.busy) busiest = std::max(busiest, b);
  for (int i = 0; i < n_local; ++i) {
    dbs[i]->max_disp = fin.max_disp; dbs[i]->build_ms = ms; dbs[i]->load_seconds = secs; dbs[i]->read_seconds = busiest;
    dbs[i]->file_bytes = fsize; outs[i] = dbs[i];
  }
  L.keep = true;
  return JB_OK;
}

int open_any(const char* path, int n_dev, const int* devices, jb_db** outs, int only_part) {
  int rc = open_impl(path, n_dev, devices, outs, only_part, -1);
  if (rc == kStashOverflow) rc = open_impl(path, n_dev, devices, outs, only_part, kLayoutWide);
  if (rc == kStashOverflow) rc = JB_E_NOMEM;
  return rc;
}

}  // namespace

extern "C" int jb_db_open(const char* path, int device, jb_db** out) {
  if (!path || !out) return fail(JB_E_ARG, "jb_db_open: null argument");
  *out = nullptr;
  int rc = open_any(path, 1, &device, out, -1);
  if (rc == JB_OK) {
    rc = build_scan_index(*out);
    if (rc != JB_OK) { jb_db_close(*out); *out = nullptr; }
  }
  return rc;
}

// ---- one process per GPU: build the own part, swap CUDA IPC handles through the host, attach the others
struct PartHandle {             // the 64 + 32 bytes a rank publishes
  cudaIpcMemHandle_t mem;
  uint64_t home_units, n_units, overlap;
  int32_t layout, cnt_bits;
};
static_assert(sizeof(PartHandle) <= JB_PART_HANDLE_BYTES, "JB_PART_HANDLE_BYTES is too small");

extern "C" int jb_db_open_part(const char* path, int device, int rank, int world, jb_db** out) {
  if (!path || !out || world < 1 || world > kMaxParts || rank < 0 || rank >= world) return fail(JB_E_ARG, "jb_db_open_part: bad argument");
  *out = nullptr;
  return open_any(path, world, &device, out, rank);
}

extern "C" int jb_db_part_handle(const jb_db* db, char* handle) {
  if (!db || !handle) return fail(JB_E_ARG, "jb_db_part_handle: null argument");
  JB_CUDA(cudaSetDevice(db->device));
  PartHandle h{};
  JB_CUDA(cudaIpcGetMemHandle(&h.mem, db->mem));
  h.home_units = db->home_units; h.n_units = db->n_units; h.overlap = db->part_overlap;
  h.layout = db->geom.layout; h.cnt_bits = db->geom.cnt_bits;
  std::memset(handle, 0, JB_PART_HANDLE_BYTES);
  std::memcpy(handle, &h, sizeof h);
  return JB_OK;
}

extern "C" int jb_db_attach_parts(jb_db* db, const char* handles) {
  if (!db || !handles) return fail(JB_E_ARG, "jb_db_attach_parts: null argument");
  if (db->parts_ready) return JB_OK;
  JB_CUDA(cudaSetDevice(db->device));
  for (int p = 0; p < db->n_parts; ++p) {
    PartHandle h;
    std::memcpy(&h, handles + (size_t)p * JB_PART_HANDLE_BYTES, sizeof h);
    if (h.home_units != db->home_units || h.n_units != db->n_units || h.overlap != db->part_overlap || h.layout != db->geom.layout || h.cnt_bits != db->geom.cnt_bits)
      return fail(JB_E_ARG, "jb_db_attach_parts: rank " + std::to_string(p) + " built a table of another geometry (set JB_TABLE_LOAD / JB_TABLE_LAYOUT to the same value on every rank)");
    if (p == db->my_part) continue;
    void* ptr = nullptr;
    JB_CUDA(cudaIpcOpenMemHandle(&ptr, h.mem, cudaIpcMemLazyEnablePeerAccess));
    db->parts[p] = static_cast<uint8_t*>(ptr);
    db->ipc_mapped[p] = true;
  }
  db->parts_ready = true;
  return JB_OK;
}

extern "C" int jb_db_open_partitioned(const char* path, int n, const int* devices, jb_db** outs) {
  if (!path || !devices || !outs || n < 1) return fail(JB_E_ARG, "jb_db_open_partitioned: bad argument");
  for (int i = 0; i < n; ++i) outs[i] = nullptr;
  return open_any(path, n, devices, outs, -1);
}

int jb_check_ready(const jb_db* db) {
  if (db && !db->parts_ready) return jb::fail(JB_E_ARG, "partitioned table: attach the other parts first (jb_db_attach_parts)");
  return JB_OK;
}

void jb_release_cached_workspaces(int device);      // jb_polish.cu

extern "C" void jb_db_close(jb_db* db) {
  if (!db) return;
  cudaSetDevice(db->device);
  jb_release_cached_workspaces(db->device);
  for (int p = 0; p < kMaxParts; ++p) if (db->ipc_mapped[p] && db->parts[p]) cudaIpcCloseMemHandle(db->parts[p]);
  free_scan_index(db);
  if (db->mem) cudaFree(db->mem);
  if (db->stash) cudaFree(db->stash);
  if (db->stash_list) cudaFree(db->stash_list);
  if (db->lut) cudaFree(db->lut);
  if (db->stream) cudaStreamDestroy(db->stream);
  delete db;
}

extern "C" int jb_db_info(const jb_db* db, jb_db_info_t* o) {
  if (!db || !o) return fail(JB_E_ARG, "jb_db_info: null argument");
  std::memset(o, 0, sizeof *o);
  const bool compact = db->geom.layout == kLayoutCompact;
  const int spu = compact ? kCompactSlots : 2;
  o->k = db->hdr.k(); o->key_bits = db->hdr.key_bits; o->counter_len = db->hdr.counter_len;
  o->canonical = db->hdr.canonical; o->matrix_r = db->hdr.matrix_r; o->matrix_c = db->hdr.matrix_c;
  o->matrix_identity = db->hdr.matrix_identity; o->device = db->device; o->hash_size = db->hdr.size;
  o->n_records = db->n_records; o->table_slots = db->n_units * (uint64_t)spu; o->table_bytes = db->table_bytes();
  o->max_displacement = db->max_disp; o->load_seconds = db->load_seconds; o->build_kernel_ms = db->build_ms;
  o->layout = db->geom.layout; o->slot_bytes = compact ? 8 : 16;
  o->remainder_bits = compact ? db->hdr.key_bits - db->geom.r_bits : db->hdr.key_bits;
  o->tag_bits = db->geom.tag_bits; o->count_bits = db->geom.cnt_bits; o->n_parts = db->n_parts;
  o->home_units = db->home_units; o->n_stash = db->n_stash;
  o->load_factor = db->home_units ? (double)db->n_records / ((double)db->home_units * spu) : 0.0;
  o->bytes_per_record = db->n_records ? (double)db->table_bytes() / (double)db->n_records : 0.0;
  o->read_seconds = db->read_seconds; o->file_bytes = db->file_bytes;
  if (db->idx_bits) {
    o->index_buckets = 1ull << db->idx_bits;
    o->index_bytes = db->idx_n_slots * 16 + (o->index_buckets + 1) * (db->idx_wide_off ? 8 : 4);
    o->index_mmer = db->idx_m; o->index_seconds = db->idx_seconds;
  }
  return JB_OK;
}

namespace {

// Compact dump (test aid, host side): rebuild every key from (sector, tag, remainder) with the inverse of the matrix's
// low block, take escaped counts and displaced records from the stash list, and check that slot order is file order.
int dump_compact(const jb_db* db, uint64_t* key_lo, uint64_t* key_hi, uint32_t* count) {
  const TableGeom& g = db->geom;
  const JfHeader& h = db->hdr;
  const int r = g.r_bits;
  const std::vector<uint64_t> inv = low_block_inverse(h, r);
  const std::vector<uint64_t> lut = build_hash_lut(h);
  std::vector<unsigned long long> words((size_t)db->n_units * kCompactSlots);
  JB_CUDA(cudaMemcpy(words.data(), db->mem, words.size() * 8, cudaMemcpyDeviceToHost));
  std::vector<Slot> stash((size_t)db->n_stash);
  if (db->n_stash) JB_CUDA(cudaMemcpy(stash.data(), db->stash_list, stash.size() * sizeof(Slot), cudaMemcpyDeviceToHost));
  struct Rec { uint64_t posx; Key key; uint32_t count; };
  std::vector<Rec> recs;
  recs.reserve((size_t)db->n_records);
  std::map<std::pair<uint64_t, uint64_t>, uint32_t> stash_count;
  auto posx_of = [&](const Key& key) { return key_posx(g, hash_lut(lut.data(), g.key_bytes, key), key); };
  for (const Slot& sr : stash) stash_count[{sr.hi, sr.lo}] = sr.count;
  const uint32_t cmask = cval_mask(g);
  const int bits = 64 - g.left_shift;
  const uint64_t tag_mask = g.tag_bits >= 64 ? ~0ull : (1ull << g.tag_bits) - 1;
  const uint64_t rmask = r >= 64 ? ~0ull : (1ull << r) - 1;
  uint64_t n_tomb = 0;
  for (uint64_t sidx = 0; sidx < words.size(); ++sidx) {
    const unsigned long long w = words[sidx];
    if (w == 0) continue;
    const uint32_t cv = (uint32_t)w & cmask;
    if (cv == cmask - 1u) { ++n_tomb; continue; }                       // tombstone: the record itself is in the stash list
    const uint64_t pat = w >> g.cnt_bits, tag = pat & tag_mask, rem = g.tag_bits >= 64 ? 0 : pat >> g.tag_bits;
    // the posx in the window of home lines [line - kMaxHops, line] that ends in `tag`
    const uint64_t line = sidx / kCompactSlots, first_line = line > (uint64_t)kMaxHops ? line - kMaxHops : 0;
    const unsigned __int128 span = (unsigned __int128)1 << bits;
    uint64_t lo_px = (uint64_t)(((unsigned __int128)first_line * span + g.units - 1) / g.units);          // smallest posx whose home line >= first_line
    uint64_t px = (lo_px & ~tag_mask) | tag;
    if (px < lo_px) px += tag_mask + 1;
    if (home_unit_of_posx(g, px) > line) return fail(JB_E_FORMAT, "compact dump: slot " + std::to_string(sidx) + " holds a tag no home line in reach explains");
    const uint64_t pos = px >> g.ext_bits;
    Key key;                                                            // high bits from the remainder, low r bits solved
    if (r == 0) { key.lo = rem; key.hi = 0; }
    else if (r >= 64) { key.lo = 0; key.hi = rem << (r - 64); }
    else { key.lo = rem << r; key.hi = rem >> (64 - r); }
    const uint64_t v = (pos ^ hash_lut(lut.data(), g.key_bytes, key)) & rmask;
    uint64_t low = 0;
    for (int b = 0; b < r; ++b) low |= (uint64_t)(__builtin_popcountll(inv[(size_t)b] & v) & 1) << b;
    key.lo |= low;
    if (posx_of(key) != px) return fail(JB_E_FORMAT, "compact dump: slot " + std::to_string(sidx) + " does not invert to a key of its own position");
    uint32_t c = cv;
    if (cv == cmask) {
      auto it = stash_count.find({key.hi, key.lo});
      if (it == stash_count.end()) return fail(JB_E_FORMAT, "compact dump: escaped record missing from the stash");
      c = it->second;
    }
    if (!recs.empty() && (recs.back().posx > px || (recs.back().posx == px && !key_less(recs.back().key, key))))
      return fail(JB_E_FORMAT, "compact dump: slot order is not (position, key) order at slot " + std::to_string(sidx));
    recs.push_back({px, key, c});
  }
  // displaced records: merge by (posx, key)
  std::vector<Rec> far;
  for (const Slot& sr : stash) {
    const Key key = {sr.lo, sr.hi};
    const uint64_t px = posx_of(key);
    const uint64_t hl = home_unit_of_posx(g, px);
    // an escaped record sits in reach of its home line and was emitted above; a tombstoned one was not
    bool emitted = false;
    for (uint64_t line = hl; line <= hl + kMaxHops && !emitted && line < db->n_units; ++line)
      for (int q = 0; q < kCompactSlots && !emitted; ++q) {
        const unsigned long long w = words[(size_t)line * kCompactSlots + q];
        emitted = w != 0 && ((uint32_t)w & cmask) == cmask && (w >> g.cnt_bits) == (compact_pattern(g, px, key) >> g.cnt_bits);
      }
    if (!emitted) far.push_back({px, key, sr.count});
  }
  if (far.size() != n_tomb) return fail(JB_E_FORMAT, "compact dump: " + std::to_string(n_tomb) + " tombstones but " + std::to_string(far.size()) + " displaced records in the stash");
  if (!far.empty()) {
    recs.insert(recs.end(), far.begin(), far.end());
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.posx != b.posx ? a.posx < b.posx : key_less(a.key, b.key); });
  }
  if (recs.size() != db->n_records) return fail(JB_E_FORMAT, "table holds " + std::to_string(recs.size()) + " records, file had " + std::to_string(db->n_records) + " (duplicate keys or zero counts?)");
  for (size_t i = 0; i < recs.size(); ++i) { key_lo[i] = recs[i].key.lo; key_hi[i] = recs[i].key.hi; count[i] = recs[i].count; }
  return JB_OK;
}

}  // namespace

extern "C" int jb_db_dump(const jb_db* db, uint64_t* key_lo, uint64_t* key_hi, uint32_t* count, uint64_t cap, uint64_t* n_out) {
  if (!db || !n_out) return fail(JB_E_ARG, "jb_db_dump: null argument");
  if (db->n_parts > 1) return fail(JB_E_UNSUPPORTED, "jb_db_dump: not available on a partitioned table");
  JB_CUDA(cudaSetDevice(db->device));
  *n_out = db->n_records;
  if (cap < db->n_records) return fail(JB_E_ARG, "jb_db_dump: arrays too small");
  if (db->n_records == 0) return JB_OK;
  if (db->geom.layout == kLayoutCompact) return dump_compact(db, key_lo, key_hi, count);
  cudaStream_t s = db->stream;
  const uint64_t n_slots = db->n_units * 2;
  const Slot* slots = reinterpret_cast<const Slot*>(db->mem);
  const int nb = (int)((n_slots + kThreads - 1) / kThreads);
  unsigned long long* d_cur = nullptr;
  uint64_t *d_lo = nullptr, *d_hi = nullptr;
  uint32_t* d_c = nullptr;
  JB_CUDA(cudaMalloc(&d_cur, (size_t)nb * 8));
  JB_CUDA(cudaMalloc(&d_lo, db->n_records * 8)); JB_CUDA(cudaMalloc(&d_hi, db->n_records * 8)); JB_CUDA(cudaMalloc(&d_c, db->n_records * 4));
  k_dump<<<nb, kThreads, 0, s>>>(slots, n_slots, d_lo, d_hi, d_c, d_cur, 0);
  std::vector<unsigned long long> cur(nb);
  JB_CUDA(cudaMemcpyAsync(cur.data(), d_cur, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  unsigned long long run = 0;
  for (int b = 0; b < nb; ++b) { unsigned long long c = cur[b]; cur[b] = run; run += c; }
  if (run != db->n_records) { cudaFree(d_cur); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_c); return fail(JB_E_FORMAT, "table holds " + std::to_string(run) + " records, file had " + std::to_string(db->n_records) + " (duplicate keys or zero counts?)"); }
  JB_CUDA(cudaMemcpyAsync(d_cur, cur.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, s));
  k_dump<<<nb, kThreads, 0, s>>>(slots, n_slots, d_lo, d_hi, d_c, d_cur, 1);
  JB_CUDA(cudaMemcpyAsync(key_lo, d_lo, db->n_records * 8, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaMemcpyAsync(key_hi, d_hi, db->n_records * 8, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaMemcpyAsync(count, d_c, db->n_records * 4, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_cur); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_c);
  return JB_OK;
}

extern "C" int jb_db_histogram(const jb_db* db, uint64_t high, uint64_t* hist) {
  if (!db || !hist) return fail(JB_E_ARG, "jb_db_histogram: null argument");
  if (!db->parts_ready) return fail(JB_E_ARG, "partitioned table: attach the other parts first (jb_db_attach_parts)");
  JB_CUDA(cudaSetDevice(db->device));
  unsigned long long* d_h = nullptr;
  JB_CUDA(cudaMalloc(&d_h, (high + 2) * 8));
  JB_CUDA(cudaMemsetAsync(d_h, 0, (high + 2) * 8, db->stream));
  const bool compact = db->geom.layout == kLayoutCompact;
  const uint64_t spu = compact ? kCompactSlots : 2;
  const int np = db->n_parts > 1 ? db->n_parts : 1;
  for (int p = 0; p < np; ++p) {                          // every part's own range (peer reads for the foreign ones)
    const uint8_t* base = db->n_parts > 1 ? db->parts[p] : db->mem;
    const uint64_t slots = (db->n_parts > 1 ? db->part_first[p + 1] - db->part_first[p] : db->n_units) * spu;
    if (compact) k_histogram_compact<<<148 * 8, 256, 0, db->stream>>>(reinterpret_cast<const unsigned long long*>(base), slots, cval_mask(db->geom), high, d_h);
    else k_histogram<<<148 * 8, 256, 0, db->stream>>>(reinterpret_cast<const Slot*>(base), slots, high, d_h);
  }
  if (compact && db->n_stash) k_histogram<<<148, 256, 0, db->stream>>>(db->stash_list, db->n_stash, high, d_h);
  JB_CUDA(cudaMemcpyAsync(hist, d_h, (high + 2) * 8, cudaMemcpyDeviceToHost, db->stream));
  JB_CUDA(cudaStreamSynchronize(db->stream));
  cudaFree(d_h);
  return JB_OK;
}

extern "C" int jb_solid_threshold(const uint64_t* counts, const uint64_t* freqs, size_t n_rows) {
  // src/jellyfish.py:8-22: walk the rows; while the frequency does not increase remember int(count/2);
  // at the first increase print it if >= 2, else exit 1.  No increase at all => no output.
  bool first = true;
  uint64_t last = 0, thr = 0;
  for (size_t i = 0; i < n_rows; ++i) {
    if (first) { last = freqs[i]; first = false; continue; }
    if (last >= freqs[i]) { last = freqs[i]; thr = counts[i] / 2; }
    else return thr < 2 ? fail(JB_E_ARG, "local minimum of the k-mer histogram gives a threshold < 2: read data not suitable for polishing") : (int)thr;
  }
  return fail(JB_E_ARG, "k-mer histogram has no local minimum");
}

// jb_table.cu -- .jf loader: streams the "binary/sorted" records of a Jellyfish database into a
// GPU-resident open-addressing table (replaces QueryMerFile's ctor: JF!swig/mer_file.i:18-36,
// JF!include/jellyfish/binary_dumper.hpp:126-146, JF!include/jellyfish/mapped_file.hpp:30-50).
//
// The file is sorted by (pos, key), pos = M*key & (size-1) (JF!include/jellyfish/sorted_dumper.hpp:72-103,
// JF!include/jellyfish/mer_heap.hpp:26-30).  home(key) is monotone in that order, so linear-probing
// placement needs neither atomics nor re-hashing of occupied slots:
//     slot_i = max(home_i, slot_{i-1} + 1) = i + max_{j<=i}(home_j - j)
// i.e. one prefix-max scan over the record stream (K0a unpack+hash, K0b block carry, K0c place).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "jb_internal.cuh"

namespace jb {
namespace {

constexpr int kRecPerBlock = 1024;   // 256 threads x 4 records
constexpr int kThreads = 256;

struct BuildState {          // device-resident carry between chunks
  long long carry;           // max_j (home_j - j) so far
  unsigned long long prev_home;       // home of the last record of the previous chunk
  unsigned long long next_prev_home;  // written by k_place, promoted by the next k_block_carry
  unsigned long long max_disp;
  unsigned long long last_slot;
  int err_unsorted;
  int err_overflow;
  int err_zero;                       // a record whose 32-bit count is 0 cannot be told from an empty slot
};

// K0a: raw records -> (key, count, home); per-block max of (home - global index).
__global__ void __launch_bounds__(kThreads)
k_unpack(const uint8_t* __restrict__ raw, uint64_t n, uint64_t base_index, int rec_bytes, int key_bytes,
         int cnt_bytes, TableView t, uint64_t* __restrict__ key_lo, uint64_t* __restrict__ key_hicnt,
         uint64_t* __restrict__ home, long long* __restrict__ block_max) {
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
    uint64_t pos = hash_lut(t.lut, key_bytes, key);
    uint64_t h = home_slot(t.g, pos, key);
    uint64_t gi = first + r;
    key_lo[gi] = key.lo;
    key_hicnt[gi] = (key.hi & 0xFFFFFFFFull) | ((uint64_t)c << 32);
    home[gi] = h;
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
  if (lane == 0) st->prev_home = st->next_prev_home;
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
// the first `overlap` slots of part p is also written behind the end of part p-1).
struct PlaceTarget {
  Slot* slots;
  int n_parts;
  uint64_t overlap;
  Slot* part[kMaxParts];
  uint64_t part_first[kMaxParts + 1];
};
__device__ __forceinline__ void put_slot(const PlaceTarget& T, uint64_t slot, const Slot& v) {
  if (T.n_parts <= 1) { T.slots[slot] = v; return; }
  int p = 0;
#pragma unroll 1
  for (int q = 1; q < T.n_parts; ++q) p += slot >= T.part_first[q] ? 1 : 0;
  if (T.part[p]) T.part[p][slot - T.part_first[p]] = v;                     // (a rank of jb_db_open_part keeps its own part only)
  if (p > 0 && slot < T.part_first[p] + T.overlap && T.part[p - 1]) T.part[p - 1][slot - T.part_first[p - 1]] = v;
}

// K0c: inclusive prefix-max inside the block, then scatter the 16-byte slots.
__global__ void __launch_bounds__(kThreads)
k_place(uint64_t n, uint64_t base_index, const uint64_t* __restrict__ key_lo, const uint64_t* __restrict__ key_hicnt,
        const uint64_t* __restrict__ home, const long long* __restrict__ block_carry, PlaceTarget target,
        uint64_t n_slots, BuildState* st) {
  __shared__ long long s_warp[kThreads / 32];
  const uint64_t first = (uint64_t)blockIdx.x * kRecPerBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread owns 4 consecutive records so the scan is a plain blocked scan
  long long t[4];
  uint64_t h[4];
  long long run = LLONG_MIN;
  bool unsorted = false;
  for (int q = 0; q < 4; ++q) {
    uint64_t gi = first + (uint64_t)threadIdx.x * 4 + q;
    if (gi < n) {
      h[q] = home[gi];
      uint64_t prev = gi > 0 ? home[gi - 1] : (base_index ? st->prev_home : 0);
      unsorted |= h[q] < prev;
      t[q] = (long long)h[q] - (long long)(base_index + gi);
    } else { h[q] = 0; t[q] = LLONG_MIN; }
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
  for (int q = 0; q < 4; ++q) {
    uint64_t gi = first + (uint64_t)threadIdx.x * 4 + q;
    if (gi >= n) break;
    long long m = t[q] > pre ? t[q] : pre;
    uint64_t slot = (uint64_t)((long long)(base_index + gi) + m);
    if (slot + 2 >= n_slots) { overflow = true; continue; }
    uint64_t hc = key_hicnt[gi];
    Slot s; s.lo = key_lo[gi]; s.hi = (uint32_t)hc; s.count = (uint32_t)(hc >> 32);
    if (s.count == 0) st->err_zero = 1;
    put_slot(target, slot, s);
    disp = max(disp, (unsigned long long)(slot - h[q]));
    last = slot;
  }
  if (disp) atomicMax(&st->max_disp, disp);
  if (last) atomicMax(&st->last_slot, last);
  if (unsorted) st->err_unsorted = 1;
  if (overflow) st->err_overflow = 1;
  for (int q = 0; q < 4; ++q)
    if (first + (uint64_t)threadIdx.x * 4 + q == n - 1) st->next_prev_home = h[q];
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

__global__ void k_histogram(const Slot* __restrict__ slots, uint64_t n_slots, unsigned long long high,
                            unsigned long long* hist) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = slots[i].count;
    if (c) atomicAdd(&hist[c > high ? high + 1 : c], 1ull);
  }
}

int ceil_log2(uint64_t x) { int b = 0; while ((1ull << b) < x && b < 63) ++b; return b; }

}  // namespace
}  // namespace jb

using namespace jb;

namespace {

// One database, n >= 1 handles.  n == 1: the whole table on devices[0].  n > 1: the slots partitioned over the devices
// (SURVEY.md §8e fallback): the records are unpacked and placed by kernels on devices[0], whose stores go straight into
// the owning part (peer stores), so the table never has to fit one GPU; every handle then probes all parts.
// only_part >= 0 (jb_db_open_part, one process per GPU): n_dev parts in all, this process builds and keeps part only_part
// on devices[0]; the other parts are attached later through CUDA IPC handles.
int open_impl(const char* path, int n_dev, const int* devices, jb_db** outs, int only_part = -1) {
  auto t0 = std::chrono::steady_clock::now();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(JB_E_CUDA, "no CUDA device: libjasper_b200 has no CPU path");
  if (n_dev < 1 || n_dev > kMaxParts) return fail(JB_E_ARG, "jb_db_open: between 1 and " + std::to_string(kMaxParts) + " devices");
  const int n_local = only_part >= 0 ? 1 : n_dev;       // handles (and devices) of this process
  for (int i = 0; i < n_local; ++i) {
    if (devices[i] < 0 || devices[i] >= ndev) return fail(JB_E_ARG, "jb_db_open: bad device index");
    for (int j = 0; j < i; ++j) if (devices[j] == devices[i]) return fail(JB_E_ARG, "jb_db_open_partitioned: a device is listed twice");
  }
  const int device = devices[0];
  if (n_dev > 1 && only_part < 0) {
    for (int i = 0; i < n_dev; ++i)
      for (int j = 0; j < n_dev; ++j) {
        if (i == j) continue;
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

  int fd = ::open(path, O_RDONLY);
  if (fd < 0) return fail(JB_E_IO, std::string("Can't open file '") + path + "'");
  struct stat st;
  fstat(fd, &st);
  const uint64_t fsize = (uint64_t)st.st_size;
  std::vector<char> head(9);
  if (fsize < 9 || pread(fd, head.data(), 9, 0) != 9) { ::close(fd); return fail(JB_E_FORMAT, "file too short for a Jellyfish header"); }
  uint64_t hlen = 0;
  for (int i = 0; i < 9; ++i) hlen = hlen * 10 + (uint64_t)((head[i] >= '0' && head[i] <= '9') ? head[i] - '0' : 0);
  head.resize(9 + std::min<uint64_t>(hlen, fsize - 9));
  if ((uint64_t)pread(fd, head.data(), head.size(), 0) != head.size()) { ::close(fd); return fail(JB_E_IO, "short read of header"); }
  std::vector<jb_db*> dbs((size_t)n_local, nullptr);
  for (int i = 0; i < n_local; ++i) dbs[i] = new jb_db();
  auto drop = [&]() { for (jb_db* d : dbs) jb_db_close(d); };
  jb_db* db = dbs[0];
  std::string err = parse_jf_header(head.data(), head.size(), &db->hdr);
  if (!err.empty()) { ::close(fd); drop(); return fail(JB_E_FORMAT, err); }
  const JfHeader& h = db->hdr;
  if (h.k() > JB_MAX_K || h.k() < 1) { ::close(fd); drop(); return fail(JB_E_UNSUPPORTED, "k = " + std::to_string(h.k()) + " outside [1, " + std::to_string(JB_MAX_K) + "]"); }
  const uint64_t body = fsize - h.data_offset;
  const int rec = h.record_bytes();
  if (body % rec) { ::close(fd); drop(); return fail(JB_E_FORMAT, "Size of database (" + std::to_string(body) + ") must be a multiple of the length of a record (" + std::to_string(rec) + ")"); }
  const uint64_t n = body / rec;

  // ---- geometry -------------------------------------------------------------------------------
  double load = 0.5;
  if (const char* e = getenv("JB_TABLE_LOAD")) { double v = atof(e); if (v > 0.05 && v <= 0.95) load = v; }
  size_t free_b = ~(size_t)0, total_b = 0;
  for (int i = 0; i < n_local; ++i) {                     // the tightest device decides
    size_t f = 0;
    JB_CUDA(cudaSetDevice(devices[i]));
    JB_CUDA(cudaMemGetInfo(&f, &total_b));
    free_b = std::min(free_b, f);
  }
  JB_CUDA(cudaSetDevice(device));
  uint64_t cap = 0, tail = 0, n_slots = 0;
  for (int attempt = 0; attempt < 8; ++attempt) {
    uint64_t half = std::max<uint64_t>(256, (uint64_t)((double)n / load / 2.0) + 1);
    cap = 2 * half;
    tail = std::max<uint64_t>(65536, cap / 64);
    n_slots = cap + tail + 8;
    if ((n_slots / (uint64_t)n_dev + kPartOverlap + tail) * sizeof(Slot) + (512ull << 20) < free_b) break;
    if (load >= 0.9) { ::close(fd); drop(); return fail(JB_E_NOMEM, "k-mer table (" + std::to_string(n) + " records) does not fit in device memory"); }
    load = std::min(0.9, load + 0.1);
  }
  int r = ceil_log2(h.size);
  int want = ceil_log2(cap / 2) + 6 - r;
  int ext = std::max(0, std::min({want, h.key_bits, 64 - r}));
  TableGeom g{};
  g.half_cap = cap / 2; g.pos_mask = h.size - 1; g.k = h.k(); g.key_bytes = h.key_bytes();
  g.ext_bits = ext; g.left_shift = 64 - (r + ext); g.key_top_shift = h.key_bits - ext;

  // parts: equal shares of [0, cap), the last one also takes the overflow tail
  uint64_t part_first[kMaxParts + 1] = {};
  uint64_t overlap = n_dev > 1 ? kPartOverlap : 0;
  for (int p2 = 0; p2 <= n_dev; ++p2) part_first[p2] = p2 == n_dev ? n_slots : (cap / (uint64_t)n_dev * (uint64_t)p2) & ~1ull;
  for (int p2 = 0; p2 + 1 < n_dev; ++p2) overlap = std::min<uint64_t>(overlap, (part_first[p2 + 2] - part_first[p2 + 1]) & ~(uint64_t)1);
  std::vector<uint64_t> lut = build_hash_lut(h);
  Slot* parts[kMaxParts] = {};
  for (int li = 0; li < n_local; ++li) {
    const int i = only_part >= 0 ? only_part : li;        // which part this handle owns
    jb_db* d = dbs[li];
    if (li) d->hdr = db->hdr;
    d->n_records = n; d->device = devices[li]; d->geom = g; d->cap = cap; d->n_slots = n_slots;
    JB_CUDA(cudaSetDevice(devices[li]));
    JB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    const uint64_t mine = part_first[i + 1] - part_first[i] + (i + 1 < n_dev ? overlap : 0);
    if (cudaMalloc(&d->slots, mine * sizeof(Slot)) != cudaSuccess) { cudaGetLastError(); ::close(fd); drop(); return fail(JB_E_NOMEM, "cannot allocate the k-mer table part on device " + std::to_string(devices[li])); }
    JB_CUDA(cudaMemsetAsync(d->slots, 0, mine * sizeof(Slot), d->stream));
    JB_CUDA(cudaMalloc(&d->lut, lut.size() * 8));
    JB_CUDA(cudaMemcpyAsync(d->lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, d->stream));
    JB_CUDA(cudaStreamSynchronize(d->stream));
    parts[i] = d->slots;
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
  PlaceTarget target{};
  target.slots = db->slots; target.n_parts = n_dev; target.overlap = overlap;
  for (int q = 0; q < n_dev; ++q) target.part[q] = parts[q];
  for (int q = 0; q <= n_dev; ++q) target.part_first[q] = part_first[q];

  // ---- stream the records ---------------------------------------------------------------------
  const uint64_t chunk_rec = std::min<uint64_t>(std::max<uint64_t>(n, 1), 4ull << 20);
  const uint64_t chunk_bytes = chunk_rec * rec;
  uint8_t* pinned[2] = {nullptr, nullptr};
  uint8_t* d_raw[2] = {nullptr, nullptr};
  cudaEvent_t copied[2], ev0, ev1;
  uint64_t *d_lo = nullptr, *d_hc = nullptr, *d_home = nullptr;
  long long* d_bmax = nullptr;
  BuildState* d_st = nullptr;
  const int max_blocks = (int)((chunk_rec + kRecPerBlock - 1) / kRecPerBlock);
  for (int b = 0; b < 2; ++b) {
    JB_CUDA(cudaMallocHost(&pinned[b], chunk_bytes + 16));
    JB_CUDA(cudaMalloc(&d_raw[b], chunk_bytes + 16));
    JB_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
  }
  JB_CUDA(cudaEventCreate(&ev0)); JB_CUDA(cudaEventCreate(&ev1));
  JB_CUDA(cudaMalloc(&d_lo, chunk_rec * 8)); JB_CUDA(cudaMalloc(&d_hc, chunk_rec * 8)); JB_CUDA(cudaMalloc(&d_home, chunk_rec * 8));
  JB_CUDA(cudaMalloc(&d_bmax, (size_t)max_blocks * 8));
  JB_CUDA(cudaMalloc(&d_st, sizeof(BuildState)));
  BuildState init{};
  init.carry = LLONG_MIN;
  JB_CUDA(cudaMemcpyAsync(d_st, &init, sizeof init, cudaMemcpyHostToDevice, s));
  TableView view = db->view();
  JB_CUDA(cudaEventRecord(ev0, s));
  uint64_t done = 0;
  int buf = 0;
  bool io_error = false;
  while (done < n) {
    uint64_t m = std::min(chunk_rec, n - done);
    JB_CUDA(cudaEventSynchronize(copied[buf]));          // previous H2D out of this pinned buffer has finished
    uint64_t want_b = m * rec, got = 0;
    while (got < want_b) {
      ssize_t rd = pread(fd, pinned[buf] + got, want_b - got, (off_t)(h.data_offset + done * rec + got));
      if (rd <= 0) { io_error = true; break; }
      got += (uint64_t)rd;
    }
    if (io_error) break;
    JB_CUDA(cudaMemcpyAsync(d_raw[buf], pinned[buf], want_b, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaEventRecord(copied[buf], s));
    int nb = (int)((m + kRecPerBlock - 1) / kRecPerBlock);
    k_unpack<<<nb, kThreads, (size_t)kRecPerBlock * rec + 16, s>>>(d_raw[buf], m, done, rec, h.key_bytes(), h.counter_len, view,
                                                                    d_lo, d_hc, d_home, d_bmax);
    k_block_carry<<<1, 32, 0, s>>>(d_bmax, nb, d_st);
    k_place<<<nb, kThreads, 0, s>>>(m, done, d_lo, d_hc, d_home, d_bmax, target, n_slots, d_st);
    done += m;
    buf ^= 1;
  }
  JB_CUDA(cudaEventRecord(ev1, s));
  BuildState fin{};
  JB_CUDA(cudaMemcpyAsync(&fin, d_st, sizeof fin, cudaMemcpyDeviceToHost, s));
  cudaError_t ce = cudaStreamSynchronize(s);
  ::close(fd);
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0, ev1);
  for (int b = 0; b < 2; ++b) { cudaFreeHost(pinned[b]); cudaFree(d_raw[b]); cudaEventDestroy(copied[b]); }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  cudaFree(d_lo); cudaFree(d_hc); cudaFree(d_home); cudaFree(d_bmax); cudaFree(d_st);
  if (ce != cudaSuccess) { drop(); return fail(JB_E_CUDA, std::string("table build: ") + cudaGetErrorString(ce)); }
  if (io_error) { drop(); return fail(JB_E_IO, "short read while streaming records"); }
  if (fin.err_unsorted) { drop(); return fail(JB_E_FORMAT, "records are not sorted by (hash position, key): not a jellyfish binary/sorted file"); }
  if (fin.err_zero) { drop(); return fail(JB_E_UNSUPPORTED, "database holds a record whose count is 0 modulo 2^32"); }
  if (fin.err_overflow) { drop(); return fail(JB_E_NOMEM, "table overflow tail exhausted (hash positions are badly skewed); lower JB_TABLE_LOAD"); }
  // a probe that starts in part p ends at the first empty slot: there must be one inside p's copy of the next part's start
  for (int p2 = 0; p2 + 1 < n_dev; ++p2) {
    if (only_part >= 0 && p2 != only_part) continue;
    std::vector<Slot> ov((size_t)overlap);
    JB_CUDA(cudaSetDevice(devices[only_part >= 0 ? 0 : p2]));
    JB_CUDA(cudaMemcpy(ov.data(), parts[p2] + (part_first[p2 + 1] - part_first[p2]), (size_t)overlap * sizeof(Slot), cudaMemcpyDeviceToHost));
    bool empty = false;
    for (uint64_t q = 0; q + 2 < overlap && !empty; ++q) empty = ov[q].count == 0;
    if (!empty) { drop(); return fail(JB_E_NOMEM, "a run of occupied slots crosses a whole partition overlap (" + std::to_string(overlap) + " slots); lower JB_TABLE_LOAD"); }
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int i = 0; i < n_local; ++i) { dbs[i]->max_disp = fin.max_disp; dbs[i]->build_ms = ms; dbs[i]->load_seconds = secs; outs[i] = dbs[i]; }
  return JB_OK;
}

}  // namespace

extern "C" int jb_db_open(const char* path, int device, jb_db** out) {
  if (!path || !out) return fail(JB_E_ARG, "jb_db_open: null argument");
  *out = nullptr;
  return open_impl(path, 1, &device, out);
}

// ---- one process per GPU: build the own part, swap CUDA IPC handles through the host, attach the others
struct PartHandle {             // the 64 + 24 bytes a rank publishes
  cudaIpcMemHandle_t mem;
  uint64_t cap, n_slots, overlap;
};
static_assert(sizeof(PartHandle) <= JB_PART_HANDLE_BYTES, "JB_PART_HANDLE_BYTES is too small");

extern "C" int jb_db_open_part(const char* path, int device, int rank, int world, jb_db** out) {
  if (!path || !out || world < 1 || world > kMaxParts || rank < 0 || rank >= world) return fail(JB_E_ARG, "jb_db_open_part: bad argument");
  *out = nullptr;
  return open_impl(path, world, &device, out, rank);
}

extern "C" int jb_db_part_handle(const jb_db* db, char* handle) {
  if (!db || !handle) return fail(JB_E_ARG, "jb_db_part_handle: null argument");
  JB_CUDA(cudaSetDevice(db->device));
  PartHandle h{};
  JB_CUDA(cudaIpcGetMemHandle(&h.mem, db->slots));
  h.cap = db->cap; h.n_slots = db->n_slots; h.overlap = db->part_overlap;
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
    if (h.cap != db->cap || h.n_slots != db->n_slots || h.overlap != db->part_overlap)
      return fail(JB_E_ARG, "jb_db_attach_parts: rank " + std::to_string(p) + " built a table of another geometry (set JB_TABLE_LOAD to the same value on every rank)");
    if (p == db->my_part) continue;
    void* ptr = nullptr;
    JB_CUDA(cudaIpcOpenMemHandle(&ptr, h.mem, cudaIpcMemLazyEnablePeerAccess));
    db->parts[p] = static_cast<Slot*>(ptr);
    db->ipc_mapped[p] = true;
  }
  db->parts_ready = true;
  return JB_OK;
}

extern "C" int jb_db_open_partitioned(const char* path, int n, const int* devices, jb_db** outs) {
  if (!path || !devices || !outs || n < 1) return fail(JB_E_ARG, "jb_db_open_partitioned: bad argument");
  for (int i = 0; i < n; ++i) outs[i] = nullptr;
  return open_impl(path, n, devices, outs);
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
  if (db->slots) cudaFree(db->slots);
  if (db->lut) cudaFree(db->lut);
  if (db->stream) cudaStreamDestroy(db->stream);
  delete db;
}

extern "C" int jb_db_info(const jb_db* db, jb_db_info_t* o) {
  if (!db || !o) return fail(JB_E_ARG, "jb_db_info: null argument");
  std::memset(o, 0, sizeof *o);
  o->k = db->hdr.k(); o->key_bits = db->hdr.key_bits; o->counter_len = db->hdr.counter_len;
  o->canonical = db->hdr.canonical; o->matrix_r = db->hdr.matrix_r; o->matrix_c = db->hdr.matrix_c;
  o->matrix_identity = db->hdr.matrix_identity; o->device = db->device; o->hash_size = db->hdr.size;
  o->n_records = db->n_records; o->table_slots = db->n_slots; o->table_bytes = db->n_slots * sizeof(Slot);
  o->max_displacement = db->max_disp; o->load_seconds = db->load_seconds; o->build_kernel_ms = db->build_ms;
  return JB_OK;
}

extern "C" int jb_db_dump(const jb_db* db, uint64_t* key_lo, uint64_t* key_hi, uint32_t* count, uint64_t cap, uint64_t* n_out) {
  if (!db || !n_out) return fail(JB_E_ARG, "jb_db_dump: null argument");
  if (db->n_parts > 1) return fail(JB_E_UNSUPPORTED, "jb_db_dump: not available on a partitioned table");
  JB_CUDA(cudaSetDevice(db->device));
  *n_out = db->n_records;
  if (cap < db->n_records) return fail(JB_E_ARG, "jb_db_dump: arrays too small");
  if (db->n_records == 0) return JB_OK;
  cudaStream_t s = db->stream;
  const int nb = (int)((db->n_slots + kThreads - 1) / kThreads);
  unsigned long long* d_cur = nullptr;
  uint64_t *d_lo = nullptr, *d_hi = nullptr;
  uint32_t* d_c = nullptr;
  JB_CUDA(cudaMalloc(&d_cur, (size_t)nb * 8));
  JB_CUDA(cudaMalloc(&d_lo, db->n_records * 8)); JB_CUDA(cudaMalloc(&d_hi, db->n_records * 8)); JB_CUDA(cudaMalloc(&d_c, db->n_records * 4));
  k_dump<<<nb, kThreads, 0, s>>>(db->slots, db->n_slots, d_lo, d_hi, d_c, d_cur, 0);
  std::vector<unsigned long long> cur(nb);
  JB_CUDA(cudaMemcpyAsync(cur.data(), d_cur, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaStreamSynchronize(s));
  unsigned long long run = 0;
  for (int b = 0; b < nb; ++b) { unsigned long long c = cur[b]; cur[b] = run; run += c; }
  if (run != db->n_records) { cudaFree(d_cur); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_c); return fail(JB_E_FORMAT, "table holds " + std::to_string(run) + " records, file had " + std::to_string(db->n_records) + " (duplicate keys or zero counts?)"); }
  JB_CUDA(cudaMemcpyAsync(d_cur, cur.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, s));
  k_dump<<<nb, kThreads, 0, s>>>(db->slots, db->n_slots, d_lo, d_hi, d_c, d_cur, 1);
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
  if (db->n_parts > 1) {
    for (int p = 0; p < db->n_parts; ++p)                 // every part's own range (peer reads for the foreign ones)
      k_histogram<<<148 * 8, 256, 0, db->stream>>>(db->parts[p], db->part_first[p + 1] - db->part_first[p], high, d_h);
  } else
    k_histogram<<<148 * 8, 256, 0, db->stream>>>(db->slots, db->n_slots, high, d_h);
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

// jb_polish.cu -- the per-batch loop of src/jasper.py main() (lines 12-32) on the GPU:
//   for ite in 0..passes:  k_scan (event bitmask)  ->  k_walk (one warp per contig)  ->  k_finalize
// Contigs stay resident in HBM between passes in two flat ping-pong buffers (per-contig regions
// with slack for insertions); per pass only the contig lengths, tallies and fix rows come back.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>

#include "jb_rows.hpp"
#include "jb_scan.cuh"

namespace jb {

std::vector<ScanTile> make_tiles(const std::vector<ScanUnit>& units);
int scan_grid(uint32_t n_tiles);

struct DevDB {
  TableView t;
  __device__ __forceinline__ uint32_t count(const Key& canonical) const { return table_count(t, canonical); }
};

constexpr int kWalkWarps = 4;     // walkers per CTA
constexpr uint32_t kTrieCap = 1u << 20;     // de-Bruijn trie nodes per walker
constexpr uint32_t kPatchCap = 1u << 18;    // longest recoverable patch

struct WalkCfg { int k, step; uint32_t solid; int fix; };

__global__ void __launch_bounds__(kWalkWarps * 32)
k_walk(TableView t, UnitIO* units, const uint32_t* order, uint32_t n_units, WalkCfg cfg, BridgeScratch* pool,
       RowSink sink, unsigned int* next_unit) {
  __shared__ WarpScratch scratch[kWalkWarps];
  const int warp = threadIdx.x >> 5;
  const int walker_id = blockIdx.x * kWalkWarps + warp;
  DevDB db{t};
  while (true) {
    unsigned int slot = 0;
    if ((threadIdx.x & 31) == 0) slot = atomicAdd(next_unit, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= n_units) break;
    UnitIO& u = units[order[slot]];
    Walker<DevDB> w(db, cfg.k, cfg.step, cfg.solid, cfg.fix != 0, u, scratch[warp], pool ? pool + walker_id : nullptr, sink);
    w.run();
    __syncwarp();
  }
}

// out[fin_olen:out_len) = in[fin_ipos:in_len)  for every contig, whole-GPU streaming copy
__global__ void k_finalize(const UnitIO* units, uint32_t n_units) {
  for (uint32_t ui = blockIdx.y; ui < n_units; ui += gridDim.y) {
    const UnitIO u = units[ui];
    if (u.status) continue;
    const int64_t n = u.in_len - u.fin_ipos;
    const uint8_t* src = u.in + u.fin_ipos;
    uint8_t* dst = u.out + u.fin_olen;
    // align the destination to 16 bytes, then 16-byte stores fed by byte-funnelled loads when the
    // relative alignment differs (indels shift the two streams against each other)
    int64_t head = (int64_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = tid; q < head; q += nth) dst[q] = src[q];
    const int64_t nvec = (n - head) / 16;
    const uint8_t* s2 = src + head;
    uint8_t* d2 = dst + head;
    if ((((uintptr_t)s2) & 15) == 0) {
      for (int64_t v = tid; v < nvec; v += nth) reinterpret_cast<uint4*>(d2)[v] = reinterpret_cast<const uint4*>(s2)[v];
    } else {
      const int sh = (int)(((uintptr_t)s2) & 3);
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 - sh);
      for (int64_t v = tid; v < nvec; v += nth) {
        uint32_t w[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) w[q] = (q < 4 || sh) ? sw[v * 4 + q] : 0u;
        uint4 o;
        o.x = __funnelshift_r(w[0], w[1], sh * 8); o.y = __funnelshift_r(w[1], w[2], sh * 8);
        o.z = __funnelshift_r(w[2], w[3], sh * 8); o.w = __funnelshift_r(w[3], w[4], sh * 8);
        reinterpret_cast<uint4*>(d2)[v] = o;
      }
    }
    for (int64_t q = head + nvec * 16 + tid; q < n; q += nth) dst[q] = src[q];
  }
}

__global__ void k_init_pool(BridgeScratch* pool, int n, uint32_t* trie, uint32_t trie_cap, uint8_t* patch, uint32_t patch_cap) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  pool[w].trie = trie + (size_t)w * trie_cap; pool[w].trie_cap = trie_cap;
  pool[w].patch = patch + (size_t)w * patch_cap; pool[w].patch_cap = patch_cap;
}

}  // namespace jb

using namespace jb;

struct jb_batch {
  const jb_db* db = nullptr;
  size_t n = 0;
  std::vector<uint64_t> off, cap;      // per-contig region in the flat buffers (bytes; multiples of 32)
  std::vector<int64_t> len;            // current lengths
  uint64_t flat_bytes = 0;
  uint8_t* buf[2] = {nullptr, nullptr};
  int cur = 0;                         // buffer holding the current sequences
  uint32_t* events = nullptr;          // flat_bytes / 32 words
  UnitIO* d_units = nullptr;
  ScanUnit* d_scan_units = nullptr;
  ScanTile* d_tiles = nullptr; size_t tiles_cap = 0;
  uint32_t* d_order = nullptr;
  unsigned int* d_next = nullptr;
  DevRow* d_rows = nullptr; unsigned long long rows_cap = 0;
  uint8_t* d_aux = nullptr; unsigned long long aux_cap = 0;
  unsigned long long* d_used = nullptr;   // [0] rows, [1] aux, [2] probe steps
  BridgeScratch* d_pool = nullptr; uint32_t* d_trie = nullptr; uint8_t* d_patch = nullptr; int n_walkers = 0;
  uint8_t* pinned_in = nullptr;
  std::vector<PassOut> passes;
  int status = JB_OK;
  bool polished = false;
  ~jb_batch() {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(events); cudaFree(d_units); cudaFree(d_scan_units); cudaFree(d_tiles);
    cudaFree(d_order); cudaFree(d_next); cudaFree(d_rows); cudaFree(d_aux); cudaFree(d_used); cudaFree(d_pool);
    cudaFree(d_trie); cudaFree(d_patch);
    if (pinned_in) cudaFreeHost(pinned_in);
  }
};

struct jb_result {
  int status = JB_OK;
  std::vector<PassOut> passes;
  std::vector<std::string> seqs;
};

extern "C" int jb_batch_upload(const jb_db* db, const jb_seq_t* seqs, size_t n, jb_batch** out) {
  if (!db || (!seqs && n) || !out) return fail(JB_E_ARG, "jb_batch_upload: null argument");
  *out = nullptr;
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  auto b = new jb_batch();
  b->db = db; b->n = n;
  b->off.resize(n); b->cap.resize(n); b->len.resize(n);
  uint64_t total = 0, total_len = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t L = seqs[i].len;
    uint64_t c = (L + L / 8 + 4096 + 31) & ~31ull;
    b->off[i] = total; b->cap[i] = c; b->len[i] = (int64_t)L;
    total += c; total_len += L;
  }
  b->flat_bytes = total;
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1));
  JB_CUDA(cudaMalloc(&b->buf[0], total + 64)); JB_CUDA(cudaMalloc(&b->buf[1], total + 64));
  JB_CUDA(cudaMalloc(&b->events, (total / 32 + 64) * 4));
  JB_CUDA(cudaMalloc(&b->d_units, std::max<size_t>(1, n) * sizeof(UnitIO)));
  JB_CUDA(cudaMalloc(&b->d_scan_units, std::max<size_t>(1, n) * sizeof(ScanUnit)));
  JB_CUDA(cudaMalloc(&b->d_order, std::max<size_t>(1, n) * 4));
  JB_CUDA(cudaMalloc(&b->d_next, 4));
  b->rows_cap = total_len / 16 + 4096;
  b->aux_cap = total_len / 4 + (1u << 20);
  JB_CUDA(cudaMalloc(&b->d_rows, b->rows_cap * sizeof(DevRow)));
  JB_CUDA(cudaMalloc(&b->d_aux, b->aux_cap));
  JB_CUDA(cudaMalloc(&b->d_used, 32));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, db->device);
  b->n_walkers = (int)std::min<size_t>(std::max<size_t>(1, n), (size_t)sms * kWalkWarps);
  b->n_walkers = (b->n_walkers + kWalkWarps - 1) / kWalkWarps * kWalkWarps;
  JB_CUDA(cudaMalloc(&b->d_pool, (size_t)b->n_walkers * sizeof(BridgeScratch)));
  JB_CUDA(cudaMalloc(&b->d_trie, (size_t)b->n_walkers * kTrieCap * 4));
  JB_CUDA(cudaMalloc(&b->d_patch, (size_t)b->n_walkers * kPatchCap));
  k_init_pool<<<(b->n_walkers + 127) / 128, 128, 0, s>>>(b->d_pool, b->n_walkers, b->d_trie, kTrieCap, b->d_patch, kPatchCap);
  JB_CUDA(cudaGetLastError());
  // stage through pinned memory so the H2D copy runs at full PCIe rate
  JB_CUDA(cudaMallocHost(&b->pinned_in, std::max<uint64_t>(total_len, 1)));
  uint64_t p = 0;
  for (size_t i = 0; i < n; ++i) { if (seqs[i].len) std::memcpy(b->pinned_in + p, seqs[i].seq, seqs[i].len); p += seqs[i].len; }
  JB_CUDA(cudaEventRecord(e0, s));
  p = 0;
  for (size_t i = 0; i < n; ++i) {
    if (seqs[i].len) JB_CUDA(cudaMemcpyAsync(b->buf[0] + b->off[i], b->pinned_in + p, seqs[i].len, cudaMemcpyHostToDevice, s));
    p += seqs[i].len;
  }
  JB_CUDA(cudaEventRecord(e1, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  note_timing("h2d", ms);
  note_counter("h2d_bytes", (int64_t)total_len);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out = b;
  return JB_OK;
}

extern "C" void jb_batch_free(jb_batch* b) { delete b; }

extern "C" int jb_batch_polish(jb_batch* b, int k, uint32_t solid_thre, int passes, int fix) {
  if (!b) return fail(JB_E_ARG, "jb_batch_polish: null batch");
  const jb_db* db = b->db;
  if (k != db->hdr.k())
    return fail(JB_E_KMISMATCH, "--ksize " + std::to_string(k) + " differs from the database's k = " + std::to_string(db->hdr.k()) +
                                    " (the reference silently mis-reads k-mers in this case)");
  if (k < 8) return fail(JB_E_UNSUPPORTED, "polishing needs k >= 8");
  if (passes < 0) return fail(JB_E_ARG, "passes must be >= 0");
  if (b->polished) return fail(JB_E_ARG, "batch already polished");
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  const size_t n = b->n;
  cudaEvent_t ev_begin, ev_end, ev_a, ev_b, ev_c;
  JB_CUDA(cudaEventCreate(&ev_begin)); JB_CUDA(cudaEventCreate(&ev_end));
  JB_CUDA(cudaEventCreate(&ev_a)); JB_CUDA(cudaEventCreate(&ev_b)); JB_CUDA(cudaEventCreate(&ev_c));
  double scan_ms = 0, walk_ms = 0;
  int64_t scan_lookups = 0, walk_lookups = 0, launches = 0, n_events = 0, probe_steps = 0;
  std::vector<UnitIO> units(n);
  std::vector<ScanUnit> sunits(n);
  std::vector<uint32_t> order(n);
  std::vector<DevRow> rows;
  std::vector<uint8_t> aux;
  const WalkCfg cfg{k, sampling_step(k), solid_thre, 0};
  const TableView view = db->view();
  b->passes.clear();
  JB_CUDA(cudaEventRecord(ev_begin, s));
  for (int ite = 0; ite <= passes && n > 0; ++ite) {
    const bool do_fix = fix && ite != passes;
    PassOut po;
    uint8_t* in = b->buf[b->cur];
    uint8_t* outb = b->buf[b->cur ^ 1];
    for (size_t i = 0; i < n; ++i) {
      po.total += b->len[i] - k + 1;                                    // may be negative (src/jasper.py:51)
      sunits[i] = ScanUnit{b->off[i], (uint64_t)b->len[i], b->off[i]};
      UnitIO& u = units[i];
      std::memset(&u, 0, sizeof u);
      u.in = in + b->off[i]; u.in_len = b->len[i]; u.out = outb + b->off[i]; u.out_cap = (int64_t)b->cap[i];
      u.ev = b->events + b->off[i] / 32; u.unit = (uint32_t)i;
    }
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return b->len[x] > b->len[y]; });
    std::vector<ScanTile> tiles = make_tiles(sunits);
    if (tiles.size() > b->tiles_cap) {
      cudaFree(b->d_tiles);
      b->tiles_cap = tiles.size() + tiles.size() / 4 + 64;
      JB_CUDA(cudaMalloc(&b->d_tiles, b->tiles_cap * sizeof(ScanTile)));
    }
    JB_CUDA(cudaMemcpyAsync(b->d_tiles, tiles.data(), tiles.size() * sizeof(ScanTile), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_scan_units, sunits.data(), n * sizeof(ScanUnit), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_units, units.data(), n * sizeof(UnitIO), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_order, order.data(), n * 4, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemsetAsync(b->d_next, 0, 4, s));
    JB_CUDA(cudaMemsetAsync(b->d_used, 0, 32, s));
    ScanArgs a{};
    a.bases = in; a.units = b->d_scan_units; a.tiles = b->d_tiles; a.n_tiles = (uint32_t)tiles.size();
    a.events = b->events; a.solid_thre = solid_thre; a.probe_steps = b->d_used + 2;
    JB_CUDA(cudaEventRecord(ev_a, s));
    k_scan<false, true><<<scan_grid(a.n_tiles), kScanThreads, 0, s>>>(view, a);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaEventRecord(ev_b, s));
    WalkCfg c2 = cfg; c2.fix = do_fix ? 1 : 0;
    RowSink sink{b->d_rows, b->rows_cap, b->d_used, b->d_aux, b->aux_cap, b->d_used + 1};
    const int blocks = b->n_walkers / kWalkWarps;
    k_walk<<<blocks, kWalkWarps * 32, 0, s>>>(view, b->d_units, b->d_order, (uint32_t)n, c2, b->d_pool, sink, b->d_next);
    JB_CUDA(cudaGetLastError());
    launches += 2;
    if (do_fix) {
      dim3 g(148, (unsigned)std::min<size_t>(n, 64));
      k_finalize<<<g, 256, 0, s>>>(b->d_units, (uint32_t)n);
      JB_CUDA(cudaGetLastError());
      ++launches;
    }
    JB_CUDA(cudaEventRecord(ev_c, s));
    unsigned long long used[4] = {0, 0, 0, 0};
    JB_CUDA(cudaMemcpyAsync(units.data(), b->d_units, n * sizeof(UnitIO), cudaMemcpyDeviceToHost, s));
    JB_CUDA(cudaMemcpyAsync(used, b->d_used, 32, cudaMemcpyDeviceToHost, s));
    JB_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_a, ev_b); scan_ms += ms;
    cudaEventElapsedTime(&ms, ev_b, ev_c); walk_ms += ms;
    probe_steps += (int64_t)used[2];
    for (size_t i = 0; i < n; ++i) {
      const UnitIO& u = units[i];
      if (u.status) {
        b->status = u.status;
        return fail(u.status, "contig " + std::to_string(i) + ": a device work pool overflowed during polishing "
                              "(output slack, fix-row arena, or de-Bruijn path pool)");
      }
      po.wrong += u.wrong;
      scan_lookups += b->len[i] + 1 + (b->len[i] > 0 ? k : 0);
      walk_lookups += (int64_t)u.lookups;
      n_events += (int64_t)u.events;
      if (do_fix) b->len[i] = u.out_len;
    }
    if (used[0] > b->rows_cap || used[1] > b->aux_cap) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "fix-row arena overflow"); }
    if (do_fix) {
      rows.resize(used[0]); aux.resize(used[1]);
      if (used[0]) JB_CUDA(cudaMemcpyAsync(rows.data(), b->d_rows, used[0] * sizeof(DevRow), cudaMemcpyDeviceToHost, s));
      if (used[1]) JB_CUDA(cudaMemcpyAsync(aux.data(), b->d_aux, used[1], cudaMemcpyDeviceToHost, s));
      JB_CUDA(cudaStreamSynchronize(s));
      if (!rows_from_log(rows, aux, &po.rows)) b->status = JB_E_REFERENCE_CRASH;   // exit 1 in the reference
      b->cur ^= 1;
    }
    b->passes.push_back(std::move(po));
    if (b->status == JB_E_REFERENCE_CRASH) break;
  }
  JB_CUDA(cudaEventRecord(ev_end, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  note_timing("polish", ms); note_timing("polish_scan", scan_ms); note_timing("polish_walk", walk_ms);
  note_counter("scan_lookups", scan_lookups); note_counter("walk_lookups", walk_lookups);
  note_counter("kernel_launches", launches); note_counter("events", n_events); note_counter("probe_steps", probe_steps);
  cudaEventDestroy(ev_begin); cudaEventDestroy(ev_end); cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); cudaEventDestroy(ev_c);
  b->polished = true;
  if (n == 0) for (int ite = 0; ite <= passes; ++ite) b->passes.push_back(PassOut{});
  return JB_OK;
}

extern "C" int jb_batch_fetch(jb_batch* b, jb_result** out) {
  if (!b || !out) return fail(JB_E_ARG, "jb_batch_fetch: null argument");
  JB_CUDA(cudaSetDevice(b->db->device));
  cudaStream_t s = b->db->stream;
  auto r = new jb_result();
  r->status = b->status;
  r->passes = b->passes;
  r->seqs.resize(b->n);
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1));
  JB_CUDA(cudaEventRecord(e0, s));
  int64_t bytes = 0;
  for (size_t i = 0; i < b->n; ++i) {
    r->seqs[i].resize((size_t)b->len[i]);
    if (b->len[i]) JB_CUDA(cudaMemcpyAsync(&r->seqs[i][0], b->buf[b->cur] + b->off[i], (size_t)b->len[i], cudaMemcpyDeviceToHost, s));
    bytes += b->len[i];
  }
  JB_CUDA(cudaEventRecord(e1, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  note_timing("d2h", ms); note_counter("d2h_bytes", bytes);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out = r;
  return JB_OK;
}

extern "C" int jb_polish(const jb_db* db, const jb_seq_t* seqs, size_t n_seqs, int k, uint32_t solid_thre, int passes,
                         int fix, jb_result** out) {
  if (!out) return fail(JB_E_ARG, "jb_polish: null out");
  *out = nullptr;
  jb_batch* b = nullptr;
  int rc = jb_batch_upload(db, seqs, n_seqs, &b);
  if (rc) return rc;
  rc = jb_batch_polish(b, k, solid_thre, passes, fix);
  if (!rc) rc = jb_batch_fetch(b, out);
  jb_batch_free(b);
  return rc;
}

extern "C" int jb_result_status(const jb_result* r) { return r ? r->status : JB_E_ARG; }
extern "C" int jb_result_iterations(const jb_result* r) { return r ? (int)r->passes.size() : JB_E_ARG; }
extern "C" int jb_result_tally(const jb_result* r, int ite, int64_t* wrong, int64_t* total) {
  if (!r || ite < 0 || ite >= (int)r->passes.size()) return fail(JB_E_ARG, "jb_result_tally: bad iteration");
  if (wrong) *wrong = r->passes[ite].wrong;
  if (total) *total = r->passes[ite].total;
  return JB_OK;
}
extern "C" int64_t jb_result_num_rows(const jb_result* r, int ite) {
  if (!r || ite < 0 || ite >= (int)r->passes.size()) return JB_E_ARG;
  return (int64_t)r->passes[ite].rows.size();
}
extern "C" int jb_result_rows(const jb_result* r, int ite, jb_fix_row_t* rows, size_t cap) {
  if (!r || ite < 0 || ite >= (int)r->passes.size() || (!rows && cap)) return fail(JB_E_ARG, "jb_result_rows: bad argument");
  const auto& v = r->passes[ite].rows;
  for (size_t i = 0; i < v.size() && i < cap; ++i) {
    std::memset(&rows[i], 0, sizeof rows[i]);
    rows[i].seq_index = v[i].seq_index; rows[i].coord = v[i].coord;
    std::strncpy(rows[i].mutation, v[i].mutation.c_str(), sizeof rows[i].mutation - 1);
    std::strncpy(rows[i].original, v[i].original.c_str(), sizeof rows[i].original - 1);
  }
  return JB_OK;
}
extern "C" int64_t jb_result_num_seqs(const jb_result* r) { return r ? (int64_t)r->seqs.size() : JB_E_ARG; }
extern "C" int jb_result_seq(const jb_result* r, size_t i, const char** data, uint64_t* len) {
  if (!r || i >= r->seqs.size() || !data || !len) return fail(JB_E_ARG, "jb_result_seq: bad argument");
  *data = r->seqs[i].data(); *len = r->seqs[i].size();
  return JB_OK;
}
extern "C" void jb_result_free(jb_result* r) { delete r; }

extern "C" int jb_db_replicate(jb_db* src, int n, const int* devices, jb_db** outs) {
  if (!src || (n && (!devices || !outs))) return fail(JB_E_ARG, "jb_db_replicate: null argument");
  for (int i = 0; i < n; ++i) {
    auto d = new jb_db(*src);
    d->device = devices[i]; d->slots = nullptr; d->lut = nullptr; d->stream = nullptr;
    JB_CUDA(cudaSetDevice(devices[i]));
    JB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    JB_CUDA(cudaMalloc(&d->slots, src->n_slots * sizeof(Slot)));
    JB_CUDA(cudaMalloc(&d->lut, (size_t)src->hdr.key_bytes() * 256 * 8));
    // device-to-device over NVLink; one disk read serves every GPU of the box
    JB_CUDA(cudaMemcpyPeerAsync(d->slots, devices[i], src->slots, src->device, src->n_slots * sizeof(Slot), d->stream));
    JB_CUDA(cudaMemcpyPeerAsync(d->lut, devices[i], src->lut, src->device, (size_t)src->hdr.key_bytes() * 256 * 8, d->stream));
    JB_CUDA(cudaStreamSynchronize(d->stream));
    outs[i] = d;
  }
  return JB_OK;
}

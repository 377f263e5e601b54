// jb_polish.cu -- the per-batch loop of src/jasper.py main() (lines 12-32) on the GPU.  Per pass:
//   k_scan          every window's count -> event / weak bitmasks (+ 1-per-1024 event summary)
//   k_find_anchors  first phase-independent restart point in every seg_block positions (jb_segments.cuh)
//   k_walk          one warp per SEGMENT runs the reference's sequential state machine
//   k_concat        stitches the segments' new bases into the other ping-pong buffer
// Contigs stay resident in HBM between passes in flat ping-pong buffers (per-contig regions with
// slack for insertions); per pass only segment summaries, tallies and fix rows come back.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>

#include "jb_plan.hpp"
#include "jb_scan.cuh"

namespace jb {

std::vector<ScanTile> make_tiles(const std::vector<ScanUnit>& units);
int scan_grid(uint32_t n_tiles);

struct DevDB {
  TableView t;
  __device__ __forceinline__ uint32_t count(const Key& canonical) const { return table_count(t, canonical); }
};

constexpr int kWalkWarps = 4;               // walkers per CTA
constexpr int kWalkCtasPerSm = 5;
constexpr uint32_t kTrieCap = 1u << 20;     // de-Bruijn trie nodes per bridging slot
constexpr uint32_t kPatchCap = 1u << 18;    // longest recoverable patch
constexpr uint32_t kPoolSlots = 256;        // bridging scratch slots shared by all walkers

struct WalkCfg { int k, step; uint32_t solid; int fix; };

__global__ void __launch_bounds__(kWalkWarps * 32)
k_walk(TableView t, UnitIO* segs, const uint32_t* order, uint32_t n_segs, WalkCfg cfg, BridgePool pool, RowSink sink,
       JobSink jobs, unsigned int* next_seg) {
  __shared__ WarpScratch scratch[kWalkWarps];
  const int warp = threadIdx.x >> 5;
  DevDB db{t};
  while (true) {
    unsigned int slot = 0;
    if ((threadIdx.x & 31) == 0) slot = atomicAdd(next_seg, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= n_segs) break;
    UnitIO& u = segs[order[slot]];
    Walker<DevDB> w(db, cfg.k, cfg.step, cfg.solid, cfg.fix != 0, u, scratch[warp], pool, sink, jobs);
    w.run();
    __syncwarp();
  }
}

// dst[0:n) = src[0:n), the CTAs of one grid row share the job; 16-byte stores, byte-funnelled loads when
// the two streams are misaligned against each other (indels shift them)
__device__ __forceinline__ void copy_span(const uint8_t* src, uint8_t* dst, int64_t n) {
  if (n <= 0) return;
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

// every copy job of an accepted segment: outb[dst_off[slot] + rel ...) = src[0:len); superseded slots have dst_off < 0
__global__ void k_concat(const CopyJob* jobs, const unsigned long long* n_jobs, const int64_t* dst_off, uint8_t* outb) {
  const unsigned long long n = *n_jobs;
  for (unsigned long long j = blockIdx.y; j < n; j += gridDim.y) {
    const CopyJob job = jobs[j];
    const int64_t off = dst_off[job.slot];
    if (off < 0) continue;
    copy_span(job.src, outb + off + job.rel, job.len);
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
  uint8_t* stage = nullptr;            // where segment walkers materialise edited bases
  int cur = 0;                         // buffer holding the current sequences
  uint32_t* events = nullptr;          // flat_bytes / 32 words
  uint32_t* weak = nullptr;
  uint32_t* events_l1 = nullptr; size_t l1_words = 0;
  UnitIO* d_segs = nullptr; size_t segs_cap = 0;
  uint32_t* d_order = nullptr;
  int64_t* d_dst_off = nullptr;
  ScanUnit* d_scan_units = nullptr;
  SegUnit* d_seg_units = nullptr;
  ScanTile* d_tiles = nullptr; size_t tiles_cap = 0;
  SegBlock* d_blocks = nullptr; long long* d_anchors = nullptr; size_t blocks_cap = 0;
  unsigned int* d_next = nullptr;
  DevRow* d_rows = nullptr; unsigned long long rows_cap = 0;
  uint8_t* d_aux = nullptr; unsigned long long aux_cap = 0;
  unsigned long long* d_used = nullptr;   // [0] rows, [1] aux, [2] probe steps, [3] copy jobs
  CopyJob* d_jobs = nullptr; unsigned long long jobs_cap = 0;
  BridgeScratch* d_pool = nullptr; uint32_t* d_trie = nullptr; uint8_t* d_patch = nullptr; uint32_t* d_busy = nullptr;
  uint8_t* pinned_in = nullptr;
  int sms = 148;
  std::vector<PassOut> passes;
  int status = JB_OK;
  bool polished = false;
  ~jb_batch() {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(stage); cudaFree(events); cudaFree(weak); cudaFree(events_l1);
    cudaFree(d_segs); cudaFree(d_order); cudaFree(d_dst_off); cudaFree(d_scan_units); cudaFree(d_seg_units); cudaFree(d_tiles);
    cudaFree(d_blocks); cudaFree(d_anchors); cudaFree(d_next); cudaFree(d_rows); cudaFree(d_aux); cudaFree(d_used); cudaFree(d_jobs);
    cudaFree(d_pool); cudaFree(d_trie); cudaFree(d_patch); cudaFree(d_busy);
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
  JB_CUDA(cudaMalloc(&b->buf[0], total + 64)); JB_CUDA(cudaMalloc(&b->buf[1], total + 64)); JB_CUDA(cudaMalloc(&b->stage, total + 64));
  JB_CUDA(cudaMalloc(&b->events, (total / 32 + 64) * 4));
  JB_CUDA(cudaMalloc(&b->weak, (total / 32 + 64) * 4));
  b->l1_words = total / 32768 + 64;
  JB_CUDA(cudaMalloc(&b->events_l1, b->l1_words * 4));
  JB_CUDA(cudaMalloc(&b->d_scan_units, std::max<size_t>(1, n) * sizeof(ScanUnit)));
  JB_CUDA(cudaMalloc(&b->d_seg_units, std::max<size_t>(1, n) * sizeof(SegUnit)));
  JB_CUDA(cudaMalloc(&b->d_next, 4));
  b->rows_cap = total_len / 16 + 4096;
  b->aux_cap = total_len / 4 + (1u << 20);
  JB_CUDA(cudaMalloc(&b->d_rows, b->rows_cap * sizeof(DevRow)));
  JB_CUDA(cudaMalloc(&b->d_aux, b->aux_cap));
  JB_CUDA(cudaMalloc(&b->d_used, 32));
  b->jobs_cap = total_len / 8 + 4 * n + 65536;
  JB_CUDA(cudaMalloc(&b->d_jobs, b->jobs_cap * sizeof(CopyJob)));
  cudaDeviceGetAttribute(&b->sms, cudaDevAttrMultiProcessorCount, db->device);
  JB_CUDA(cudaMalloc(&b->d_pool, (size_t)kPoolSlots * sizeof(BridgeScratch)));
  JB_CUDA(cudaMalloc(&b->d_trie, (size_t)kPoolSlots * kTrieCap * 4));
  JB_CUDA(cudaMalloc(&b->d_patch, (size_t)kPoolSlots * kPatchCap));
  JB_CUDA(cudaMalloc(&b->d_busy, (size_t)kPoolSlots * 4));
  JB_CUDA(cudaMemsetAsync(b->d_busy, 0, (size_t)kPoolSlots * 4, s));
  k_init_pool<<<(kPoolSlots + 127) / 128, 128, 0, s>>>(b->d_pool, (int)kPoolSlots, b->d_trie, kTrieCap, b->d_patch, kPatchCap);
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

namespace {

int64_t pick_seg_block(int64_t total_len, int sms) {
  // (multiples of 32: k_find_anchors walks whole bitmask words)
  if (const char* e = getenv("JB_SEG_BLOCK")) { long v = atol(e); if (v >= 64) return (v + 31) / 32 * 32; if (v == 0) return 0; }
  // enough segments to fill the resident walkers a few times over, but not so short that seams dominate
  const int64_t resident = (int64_t)sms * kWalkCtasPerSm * kWalkWarps;
  int64_t blk = total_len / (4 * resident);
  blk = std::max<int64_t>(1024, std::min<int64_t>(8192, blk));
  return (blk + 31) / 32 * 32;
}

UnitIO make_seg(const jb_batch* b, const SegPlan& p, uint8_t* in, uint32_t gen, uint32_t slot, int keep_margin) {
  UnitIO u;
  std::memset(&u, 0, sizeof u);
  const size_t i = p.unit;
  u.in = in + b->off[i]; u.in_len = b->len[i];
  u.out = b->stage + b->off[i] + p.stage_off; u.out_cap = p.stage_cap;
  u.ev = b->events + b->off[i] / 32; u.ev_l1 = b->events_l1; u.ev_bit_base = b->off[i];
  u.base = p.base; u.next_anchor = p.next_anchor;
  u.unit = p.unit; u.seg = p.seg; u.gen = gen; u.slot = slot;
  u.materialize_all = gen ? 1u : 0u;               // re-runs take the conservative path
  u.keep_margin = keep_margin;
  return u;
}

}  // namespace

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
  cudaEvent_t ev_begin, ev_end, ev_a, ev_b, ev_c, ev_d, ev_e;
  JB_CUDA(cudaEventCreate(&ev_e));
  JB_CUDA(cudaEventCreate(&ev_begin)); JB_CUDA(cudaEventCreate(&ev_end));
  JB_CUDA(cudaEventCreate(&ev_a)); JB_CUDA(cudaEventCreate(&ev_b)); JB_CUDA(cudaEventCreate(&ev_c)); JB_CUDA(cudaEventCreate(&ev_d));
  double scan_ms = 0, walk_ms = 0, plan_ms = 0;
  int64_t scan_lookups = 0, walk_lookups = 0, launches = 0, n_events = 0, probe_steps = 0, n_segments = 0, n_reruns = 0;
  int64_t max_seg_events = 0, max_seg_lookups = 0, max_seg_span = 0;
  double kwalk_ms = 0;
  std::vector<ScanUnit> sunits(n);
  std::vector<SegUnit> gunits(n);
  std::vector<int64_t> caps(n);
  std::vector<DevRow> rows;
  std::vector<uint8_t> aux;
  std::vector<long long> anchors;
  const WalkCfg cfg{k, sampling_step(k), solid_thre, 0};
  const TableView view = db->view();
  const BridgePool pool{b->d_pool, kPoolSlots, b->d_busy};
  int64_t total_len = 0;
  for (size_t i = 0; i < n; ++i) { caps[i] = (int64_t)b->cap[i]; total_len += b->len[i]; }
  const int64_t seg_block = pick_seg_block(total_len, b->sms);
  int64_t margin = anchor_margin(k);
  if (const char* e = getenv("JB_ANCHOR_MARGIN")) margin = std::max<long>(0, atol(e));   // tests: provoke seam re-runs
  int keep_margin = default_keep_margin(k);
  if (const char* e = getenv("JB_KEEP_MARGIN")) keep_margin = (int)std::max<long>(0, atol(e));   // tests: provoke far-back re-runs
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
      gunits[i] = SegUnit{b->off[i], (uint64_t)b->len[i]};
      scan_lookups += b->len[i] + 1 + (b->len[i] > 0 ? k : 0);
    }
    // ---- scan: counts -> event / weak bits
    std::vector<ScanTile> tiles = make_tiles(sunits);
    if (tiles.size() > b->tiles_cap) {
      cudaFree(b->d_tiles);
      b->tiles_cap = tiles.size() + tiles.size() / 4 + 64;
      JB_CUDA(cudaMalloc(&b->d_tiles, b->tiles_cap * sizeof(ScanTile)));
    }
    std::vector<BlockRef> blocks;
    if (seg_block > 0) blocks = make_blocks(b->len, seg_block);
    if (blocks.size() > b->blocks_cap) {
      cudaFree(b->d_blocks); cudaFree(b->d_anchors);
      b->blocks_cap = blocks.size() + blocks.size() / 4 + 64;
      JB_CUDA(cudaMalloc(&b->d_blocks, b->blocks_cap * sizeof(SegBlock)));
      JB_CUDA(cudaMalloc(&b->d_anchors, b->blocks_cap * sizeof(long long)));
    }
    static_assert(sizeof(BlockRef) == sizeof(SegBlock), "host and device block descriptors must match");
    JB_CUDA(cudaMemcpyAsync(b->d_tiles, tiles.data(), tiles.size() * sizeof(ScanTile), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_scan_units, sunits.data(), n * sizeof(ScanUnit), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_seg_units, gunits.data(), n * sizeof(SegUnit), cudaMemcpyHostToDevice, s));
    if (!blocks.empty()) JB_CUDA(cudaMemcpyAsync(b->d_blocks, blocks.data(), blocks.size() * sizeof(SegBlock), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemsetAsync(b->d_next, 0, 4, s));
    JB_CUDA(cudaMemsetAsync(b->d_used, 0, 32, s));
    JB_CUDA(cudaMemsetAsync(b->events_l1, 0, b->l1_words * 4, s));
    ScanArgs a{};
    a.bases = in; a.units = b->d_scan_units; a.tiles = b->d_tiles; a.n_tiles = (uint32_t)tiles.size();
    a.events = b->events; a.weak = b->weak; a.events_l1 = b->events_l1; a.solid_thre = solid_thre; a.probe_steps = b->d_used + 2;
    JB_CUDA(cudaEventRecord(ev_a, s));
    k_scan<false, true><<<scan_grid(a.n_tiles), kScanThreads, 0, s>>>(view, a);
    JB_CUDA(cudaGetLastError());
    ++launches;
    JB_CUDA(cudaEventRecord(ev_b, s));
    // ---- anchors -> segments
    anchors.assign(blocks.size(), -1);
    if (!blocks.empty()) {
      const uint32_t nb = (uint32_t)blocks.size();
      k_find_anchors<<<(nb + 7) / 8, 256, 0, s>>>(b->events, b->weak, b->d_seg_units, b->d_blocks, nb, k, (int)seg_block, (int)margin, b->d_anchors);
      JB_CUDA(cudaGetLastError());
      ++launches;
      JB_CUDA(cudaMemcpyAsync(anchors.data(), b->d_anchors, nb * sizeof(long long), cudaMemcpyDeviceToHost, s));
      JB_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<SegPlan> plan = plan_segments(b->len, caps, blocks, anchors.data());
    std::vector<UnitIO> segs(plan.size());
    for (size_t q = 0; q < plan.size(); ++q) segs[q] = make_seg(b, plan[q], in, 0, (uint32_t)q, keep_margin);
    const size_t need = segs.size() + n;                  // room for whole-contig re-runs behind the plan
    if (need > b->segs_cap) {
      cudaFree(b->d_segs); cudaFree(b->d_order); cudaFree(b->d_dst_off);
      b->segs_cap = need + need / 4 + 64;
      JB_CUDA(cudaMalloc(&b->d_segs, b->segs_cap * sizeof(UnitIO)));
      JB_CUDA(cudaMalloc(&b->d_order, b->segs_cap * 4));
      JB_CUDA(cudaMalloc(&b->d_dst_off, b->segs_cap * 8));
    }
    std::vector<uint32_t> order(segs.size());
    std::iota(order.begin(), order.end(), 0u);
    auto seg_span = [&](uint32_t q) { return (segs[q].next_anchor >= 0 ? segs[q].next_anchor : segs[q].in_len) - segs[q].base; };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return seg_span(x) > seg_span(y); });
    JB_CUDA(cudaMemcpyAsync(b->d_segs, segs.data(), segs.size() * sizeof(UnitIO), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(b->d_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice, s));
    // ---- walk
    WalkCfg c2 = cfg; c2.fix = do_fix ? 1 : 0;
    RowSink sink{b->d_rows, b->rows_cap, b->d_used, b->d_aux, b->aux_cap, b->d_used + 1};
    JobSink jobs{b->d_jobs, b->jobs_cap, b->d_used + 3};
    auto walk_grid = [&](size_t cnt) { return (unsigned)std::min<size_t>((cnt + kWalkWarps - 1) / kWalkWarps, (size_t)b->sms * kWalkCtasPerSm); };
    JB_CUDA(cudaEventRecord(ev_c, s));
    k_walk<<<walk_grid(segs.size()), kWalkWarps * 32, 0, s>>>(view, b->d_segs, b->d_order, (uint32_t)segs.size(), c2, pool, sink, jobs, b->d_next);
    JB_CUDA(cudaGetLastError());
    ++launches;
    JB_CUDA(cudaEventRecord(ev_e, s));
    JB_CUDA(cudaMemcpyAsync(segs.data(), b->d_segs, segs.size() * sizeof(UnitIO), cudaMemcpyDeviceToHost, s));
    JB_CUDA(cudaStreamSynchronize(s));
    { float m2 = 0; cudaEventElapsedTime(&m2, ev_c, ev_e); kwalk_ms += m2; }
    n_segments += (int64_t)segs.size();
    // ---- contigs whose seams did not hold (or whose segment slack ran out) are walked again as one segment
    std::vector<uint8_t> redo(n, 0);
    size_t n_redo = 0, n_slots = segs.size();
    for (const UnitIO& u : segs) {
      if (!u.status) continue;
      const bool split = plan[&u - segs.data()].n_segs > 1;
      if (u.status == kFarBack || (split && (u.status == kSeamViolation || u.status == JB_E_SCRATCH))) { if (!redo[u.unit]) { redo[u.unit] = 1; ++n_redo; } continue; }
      b->status = u.status;
      return fail(JB_E_SCRATCH, "contig " + std::to_string(u.unit) + ": a device work pool overflowed during polishing "
                                "(output slack, fix-row arena, or de-Bruijn path pool; status " + std::to_string(u.status) + ")");
    }
    if (n_redo) {
      std::vector<UnitIO> again;
      for (size_t i = 0; i < n; ++i) if (redo[i]) again.push_back(make_seg(b, whole_contig((uint32_t)i, caps[i]), in, 1, (uint32_t)(segs.size() + again.size()), keep_margin));
      std::vector<uint32_t> ord2(again.size());
      for (size_t q = 0; q < again.size(); ++q) ord2[q] = (uint32_t)(segs.size() + q);
      JB_CUDA(cudaMemcpyAsync(b->d_segs + segs.size(), again.data(), again.size() * sizeof(UnitIO), cudaMemcpyHostToDevice, s));
      JB_CUDA(cudaMemcpyAsync(b->d_order, ord2.data(), ord2.size() * 4, cudaMemcpyHostToDevice, s));
      JB_CUDA(cudaMemsetAsync(b->d_next, 0, 4, s));
      k_walk<<<walk_grid(again.size()), kWalkWarps * 32, 0, s>>>(view, b->d_segs, b->d_order, (uint32_t)again.size(), c2, pool, sink, jobs, b->d_next);
      JB_CUDA(cudaGetLastError());
      ++launches;
      JB_CUDA(cudaMemcpyAsync(again.data(), b->d_segs + segs.size(), again.size() * sizeof(UnitIO), cudaMemcpyDeviceToHost, s));
      JB_CUDA(cudaStreamSynchronize(s));
      n_reruns += (int64_t)again.size();
      std::vector<UnitIO> merged;
      size_t ai = 0;
      for (size_t q = 0; q < segs.size();) {
        const uint32_t unit = segs[q].unit;
        if (redo[unit]) {
          const UnitIO& u = again[ai++];
          if (u.status) {
            b->status = u.status;
            return fail(JB_E_SCRATCH, "contig " + std::to_string(unit) + ": a device work pool overflowed during polishing (status " + std::to_string(u.status) + ")");
          }
          merged.push_back(u);
          while (q < segs.size() && segs[q].unit == unit) ++q;
        } else merged.push_back(segs[q++]);
      }
      n_slots = segs.size() + again.size();
      segs.swap(merged);
    }
    for (const UnitIO& u : segs) {
      po.wrong += u.wrong; walk_lookups += (int64_t)u.lookups; n_events += (int64_t)u.events;
      max_seg_events = std::max<int64_t>(max_seg_events, (int64_t)u.events);
      max_seg_lookups = std::max<int64_t>(max_seg_lookups, (int64_t)u.lookups);
      max_seg_span = std::max<int64_t>(max_seg_span, u.end_old - u.base);
    }
    unsigned long long used[4] = {0, 0, 0, 0};
    JB_CUDA(cudaMemcpyAsync(used, b->d_used, 32, cudaMemcpyDeviceToHost, s));
    JB_CUDA(cudaStreamSynchronize(s));
    if (do_fix) {
      // ---- stitch the segments into the other buffer
      std::vector<int64_t> new_lens(b->len);
      std::vector<SegPlace> place = place_segments(segs, &new_lens);
      std::vector<int64_t> dst_off(n_slots, -1);       // by launch slot; superseded attempts stay at -1
      std::vector<size_t> first_seg(n, 0);
      for (size_t q = 0; q < segs.size(); ++q) {
        dst_off[segs[q].slot] = (int64_t)b->off[segs[q].unit] + place[q].new_off;
        if (segs[q].seg == 0) first_seg[segs[q].unit] = q;
      }
      if (used[3] > b->jobs_cap) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "copy-job arena overflow"); }
      for (size_t i = 0; i < n; ++i)
        if (new_lens[i] > caps[i]) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "contig " + std::to_string(i) + " outgrew its output slack"); }
      JB_CUDA(cudaMemcpyAsync(b->d_dst_off, dst_off.data(), dst_off.size() * 8, cudaMemcpyHostToDevice, s));
      dim3 g(4, (unsigned)std::min<unsigned long long>(std::max<unsigned long long>(used[3], 1), 16384));
      k_concat<<<g, 256, 0, s>>>(b->d_jobs, b->d_used + 3, b->d_dst_off, outb);
      JB_CUDA(cudaGetLastError());
      ++launches;
      JB_CUDA(cudaEventRecord(ev_d, s));
      JB_CUDA(cudaStreamSynchronize(s));
      if (used[0] > b->rows_cap || used[1] > b->aux_cap) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "fix-row arena overflow"); }
      rows.resize(used[0]); aux.resize(used[1]);
      if (used[0]) JB_CUDA(cudaMemcpyAsync(rows.data(), b->d_rows, used[0] * sizeof(DevRow), cudaMemcpyDeviceToHost, s));
      if (used[1]) JB_CUDA(cudaMemcpyAsync(aux.data(), b->d_aux, used[1], cudaMemcpyDeviceToHost, s));
      JB_CUDA(cudaStreamSynchronize(s));
      settle_rows(&rows, segs, place, first_seg);
      if (!rows_from_log(rows, aux, &po.rows)) b->status = JB_E_REFERENCE_CRASH;   // exit 1 in the reference
      b->len = new_lens;
      b->cur ^= 1;
    } else {
      JB_CUDA(cudaEventRecord(ev_d, s));
      JB_CUDA(cudaStreamSynchronize(s));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_a, ev_b); scan_ms += ms;
    cudaEventElapsedTime(&ms, ev_b, ev_c); plan_ms += ms;
    cudaEventElapsedTime(&ms, ev_c, ev_d); walk_ms += ms;
    probe_steps += (int64_t)used[2];
    b->passes.push_back(std::move(po));
    if (b->status == JB_E_REFERENCE_CRASH) break;
  }
  JB_CUDA(cudaEventRecord(ev_end, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  note_timing("polish", ms); note_timing("polish_scan", scan_ms); note_timing("polish_walk", walk_ms); note_timing("polish_plan", plan_ms);
  note_counter("scan_lookups", scan_lookups); note_counter("walk_lookups", walk_lookups);
  note_counter("kernel_launches", launches); note_counter("events", n_events); note_counter("probe_steps", probe_steps);
  note_timing("polish_kwalk", kwalk_ms);
  note_counter("max_seg_events", max_seg_events); note_counter("max_seg_lookups", max_seg_lookups); note_counter("max_seg_span", max_seg_span);
  note_counter("segments", n_segments); note_counter("segment_reruns", n_reruns); note_counter("seg_block", seg_block);
  cudaEventDestroy(ev_begin); cudaEventDestroy(ev_end); cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); cudaEventDestroy(ev_c); cudaEventDestroy(ev_d); cudaEventDestroy(ev_e);
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

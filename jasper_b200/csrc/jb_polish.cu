// jb_polish.cu -- the per-batch loop of src/jasper.py main() (lines 12-32) on the GPU.  Per pass:
//   k_scan          every window's count (one random 128-byte line of the table per window)
//   k_events        counts -> event / weak bitmasks (+ 1-per-1024 event summary)
//   k_find_anchors  first phase-independent restart point in every seg_block positions (jb_segments.cuh)
//   k_walk          one warp per SEGMENT runs the reference's sequential state machine
//   k_concat        stitches the segments' new bases into the other ping-pong buffer
// Contigs stay resident in HBM between passes in flat ping-pong buffers (per-contig regions with
// slack for insertions); per pass only segment summaries, tallies and fix rows come back.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>

#include "jb_plan.hpp"
#include "jb_scan.cuh"

namespace jb {



// One copy of hash + probe in the walker's image: the walker is bound by instruction fetch (ncu: stall_no_instruction
// > 50 % with the hot loop spread over ~160 KB of SASS), so every call site branches here instead of inlining it.
__device__ __noinline__ uint32_t probe_count(const TableView* t, uint64_t lo, uint64_t hi) {
  const Key key = {lo, hi};
  return table_count(*t, key);
}
struct DevDB {
  TableView t;
  __device__ __forceinline__ uint32_t count(const Key& canonical) const { return probe_count(&t, canonical.lo, canonical.hi); }
};

constexpr int kWalkWarps = 4;               // walkers per CTA
#ifndef JB_WALK_CTAS
#define JB_WALK_CTAS 4
#endif
constexpr int kWalkCtasPerSm = JB_WALK_CTAS;
// Bridging scratch: about one slot per resident walker (592 CTAs x 4 warps on a B200), so that a warp with a multi-error
// stretch hardly ever waits for one.  With 512 slots of 1 M nodes k_verdicts2 spent 6 % of its samples asleep in
// acquire_bridge and 7.5 % of its instructions polling (profiles/r2aa_*); 2048 x 256 K nodes: both verdict kernels
// 1.60 -> 1.38 ms per step, no search of either bench workload outgrows a slot (profiles/r2ae_pool_ab.txt).  A search that
// does reports kTrieFull and its contig is re-walked with the grown pool (ensure_big_pool below).
#ifndef JB_TRIE_CAP_LOG2
#define JB_TRIE_CAP_LOG2 18
#endif
#ifndef JB_POOL_SLOTS
#define JB_POOL_SLOTS 2048
#endif
constexpr uint32_t kTrieCap = 1u << JB_TRIE_CAP_LOG2;   // de-Bruijn trie nodes per bridging slot
constexpr uint32_t kPatchCap = 1u << 18;    // longest recoverable patch
constexpr uint32_t kPoolSlots = JB_POOL_SLOTS;          // bridging scratch slots shared by all walkers

struct WalkCfg { int k, step; uint32_t solid; int fix; };

#ifndef JB_WALK_MINBLOCKS
#define JB_WALK_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kWalkWarps * 32, JB_WALK_MINBLOCKS)
k_walk(TableView t, UnitIO* segs, const uint32_t* order, uint32_t n_segs, WalkCfg cfg, BridgePool pool, RowSink sink,
       JobSink jobs, unsigned int* next_seg) {
  __shared__ WarpScratch scratch[kWalkWarps];
  __shared__ DevDB db;                     // shared, so the out-of-line scoring helpers take it by reference
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) db.t = t;
  __syncthreads();
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

// One warp per 1024-position tile of the weak bitmask: every weak run that ENDS inside the tile gets its verdict
// (jb_verdict.cuh), published under the run's end in vmap.  All warps run the same few KB of code (delimit, pack,
// score, probe), which is the point: this work used to sit inside the walkers' divergent, fetch-starved kernel.
struct VerdictArgs {
  const uint8_t* bases; const ScanUnit* units; const uint32_t* tile_first; uint32_t n_units, n_tiles;
  const uint32_t* counts; const uint8_t* okb; const uint32_t* weak;
  Verdict* verdicts; unsigned long long cap; unsigned long long* used; uint32_t* vmap;
  uint32_t* verdict_unit;       // contig of every verdict (k_verdicts2 needs it)
  Verdict2* verdicts2;
  uint8_t* dry_out;             // k_verdicts2: kDryOut bytes of edit scratch per resident warp
  int dry_post;                 // k_verdicts2 also evaluates the scan step behind the fix
  unsigned long long* next;     // k_verdicts2: work queue head (zeroed before the launch)
  int k, step; uint32_t solid;
};
constexpr int kVerdictWarps = 4;
constexpr int kDryOut = 1024;
#ifndef JB_VERDICT_MINBLOCKS
#define JB_VERDICT_MINBLOCKS 8       // (r2y: 64 registers, a few spilled words, twice the warps: 0.65 -> 0.50 ms on the clean workload; the walkers lose at 6 and 8)
#endif
__global__ void __launch_bounds__(kVerdictWarps * 32, JB_VERDICT_MINBLOCKS)
k_verdicts(TableView t, VerdictArgs a) {
  __shared__ WarpScratch scratch[kVerdictWarps];
  __shared__ DevDB db;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) db.t = t;
  __syncthreads();
  WarpScratch& s = scratch[warp];
  for (uint32_t ti = blockIdx.x * kVerdictWarps + warp; ti < a.n_tiles; ti += gridDim.x * kVerdictWarps) {
    const ScanTile tile = decode_tile(a.tile_first, a.n_units, ti);
    const ScanUnit u = a.units[tile.unit];
    const int64_t len = (int64_t)u.len;
    const uint32_t* wk = a.weak + (u.out_off >> 5);
    const int64_t wp = (int64_t)tile.start + 32 * lane;                  // first position of this lane's word
    uint32_t ends = 0;                                                   // positions whose bit is clear and whose predecessor's is set
    if (wp <= len) {
      const uint32_t cur = wk[wp >> 5];
      const uint32_t prev_bit = wp > 0 ? (wk[(wp >> 5) - 1] >> 31) : 0u;
      ends = ~cur & ((cur << 1) | prev_bit);
    }
    while (true) {
      const unsigned ball = __ballot_sync(0xffffffffu, ends != 0u);
      if (!ball) break;
      const int L = __ffs((int)ball) - 1;
      const uint32_t e = __shfl_sync(0xffffffffu, ends, L);
      const int64_t re = (int64_t)tile.start + 32 * L + (__ffs((int)e) - 1);
      if (lane == L) ends &= ends - 1;
      if (re > len || re < 1) continue;
      Verdict* slot = reinterpret_cast<Verdict*>(s.fx);                  // (16 bytes of scratch: the row text area is idle here)
      int64_t re2 = -1;
      const bool ok = make_verdict(db, s, a.bases + u.seq_off, a.counts + u.out_off, a.okb + u.out_off, len, a.k, a.step,
                                   (int64_t)a.solid, re - 1, slot, &re2, a.dry_post != 2);
      if (ok && re2 == re && lane == 0) {
        const unsigned long long id = atomicAdd(a.used, 1ull);
        if (id < a.cap) {
          a.verdicts[id] = *slot;
          a.verdict_unit[id] = tile.unit;
          uint32_t* cell = a.vmap + ((u.out_off + (uint64_t)re) >> 5);
          if (atomicCAS(cell, 0u, (uint32_t)id + 1u) != 0u) atomicExch(cell, kVmapAmbiguous);
        }
      }
      __syncwarp();
    }
  }
}

// Second stage: one warp per verdict that only carries bounds (kVDelimit) runs the
// walker's own dispatch on the pristine contig (Walker::dry_fix) and stores what it decided.  Same code as in k_walk,
// but one independent event per warp: the homopolymer / diploid repairs leave the walkers' sequential chains.
__global__ void __launch_bounds__(kWalkWarps * 32, JB_WALK_MINBLOCKS)
k_verdicts2(TableView t, VerdictArgs a, BridgePool pool) {
  __shared__ WarpScratch scratch[kWalkWarps];
  __shared__ DevDB db;
  __shared__ UnitIO units[kWalkWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) db.t = t;
  __syncthreads();
  const unsigned long long n = min(*a.used, a.cap);
  const RowSink no_rows{nullptr, 0, nullptr, nullptr, 0, nullptr};
  const JobSink no_jobs{nullptr, 0, nullptr};
  while (true) {                                       // events cost anything from a few to hundreds of microseconds: dynamic queue
    unsigned long long id = 0;
    if (lane == 0) id = atomicAdd(a.next, 1ull);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id >= n) break;
    const Verdict v = a.verdicts[id];
    Verdict2* out = a.verdicts2 + id;
    const ScanUnit su = a.units[a.verdict_unit[id]];
    const int nbad = (int)v.d_ga - (int)v.d_s0;
    (void)nbad;
    if (v.kind != kVDelimit || su.len >= 0xFFFFFFFFull) { if (lane == 0) out->valid = 0; continue; }
    UnitIO& u = units[warp];
    __syncwarp();
    if (lane == 0) {
      u = UnitIO{};
      u.in = a.bases + su.seq_off; u.in_len = (int64_t)su.len;
      u.cnt = a.counts + su.out_off; u.okb = a.okb + su.out_off;
      u.next_anchor = -1; u.keep_margin = default_keep_margin(a.k);
      // a segment that starts shortly in front of the run: the edit and the scan step behind it fit kDryOut bytes of scratch
      const int64_t s0 = (int64_t)v.re - (int64_t)v.d_jq + (int64_t)v.d_s0;
      u.base = s0 > 64 ? s0 - 64 : 0;
      u.out = a.dry_out + ((size_t)blockIdx.x * kWalkWarps + warp) * kDryOut; u.out_cap = a.dry_post ? kDryOut : 0;
    }
    __syncwarp();
    Walker<DevDB> w(db, a.k, a.step, a.solid, true, u, scratch[warp], pool, no_rows, no_jobs);
    w.dry_fix(v, (int64_t)v.re, out);
    __syncwarp();
  }
}

// Small results go to pinned host memory either through the copy engine or -- while the polished bases occupy it
// (early fetch, side stream) -- by a kernel that stores straight into the mapped host buffer: a device-to-host copy on
// the main stream would queue behind the whole 64 MB transfer.
__global__ void k_store_host(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t nv = n / 16;
  for (size_t v = i; v < nv; v += nth) reinterpret_cast<uint4*>(dst)[v] = reinterpret_cast<const uint4*>(src)[v];
  for (size_t q = nv * 16 + i; q < n; q += nth) dst[q] = src[q];
}
// ... and the other way: a pass's small uploads are read straight out of the mapped pinned staging area (4-byte words)
__global__ void k_load_host(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n_words) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t v = i; v < n_words; v += nth) dst[v] = src[v];
}

// dst[0:n) = src[0:n) by threads tid of nth; 16-byte stores, byte-funnelled loads when the two streams are misaligned
// against each other (indels shift them)
__device__ __forceinline__ void copy_span(const uint8_t* src, uint8_t* dst, int64_t n, int64_t tid, int64_t nth) {
  if (n <= 0) return;
  int64_t head = (int64_t)((16 - ((uintptr_t)dst & 15)) & 15);
  if (head > n) head = n;
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

// every copy job of an accepted segment: outb[dst_off[slot] + rel ...) = src[0:len); superseded slots have dst_off < 0.
// Fixed 1-D grid, the job list shared out on the device (job_split): launched right behind k_place, no host in between.
__global__ void __launch_bounds__(256)
k_concat(const CopyJob* jobs, const unsigned long long* n_jobs, unsigned long long jobs_cap, const int64_t* dst_off, uint8_t* outb) {
  const unsigned long long n = min(*n_jobs, jobs_cap);
  const JobSplit sp = job_split(n, false, 128u);
  for (unsigned long long j = sp.j0; j < n; j += sp.jstep) {
    const CopyJob job = jobs[j];
    const int64_t off = dst_off[job.slot];
    if (off < 0) continue;
    copy_span(job.src, outb + off + job.rel, job.len, sp.tid, sp.nth);
  }
}

// ---- segment descriptors on the device: the host uploads 40 bytes of plan per segment, not a 208-byte descriptor ----------
struct SegPlanDev { uint32_t unit, seg; int64_t base, next_anchor, stage_off, stage_cap; };
struct SegConsts {
  const uint8_t* in; uint8_t* stage; const uint32_t* counts; const uint8_t* okb; const uint32_t* weak; const uint32_t* vmap;
  const Verdict* verdicts; const Verdict2* verdicts2; const uint32_t* events; const uint32_t* ev_l1; const uint32_t* ev_l2;
  const ScanUnit* units; int32_t keep_margin;
};
// (the same fields make_seg fills on the host for a re-walk)
__global__ void k_make_segs(const SegPlanDev* __restrict__ plan, uint32_t n, SegConsts c, UnitIO* __restrict__ out) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const SegPlanDev p = plan[q];
  const ScanUnit su = c.units[p.unit];
  const uint64_t off = su.seq_off;
  UnitIO u;
  u.status = 0; u.n_rows = 0; u.new_len = 0; u.end_old = 0; u.wrong = 0; u.lookups = 0; u.events = 0; u.n_fast = 0; u.n_post = 0; u.n_v2 = 0;
  u.pad2 = 0; u.cycles = 0;
  u.in = c.in + off; u.in_len = (int64_t)su.len;
  u.out = c.stage + off + p.stage_off; u.out_cap = p.stage_cap;
  u.cnt = c.counts + off; u.okb = c.okb + off;
  u.weak = c.weak + off / 32; u.vmap = c.vmap; u.verdicts = c.verdicts; u.verdicts2 = c.verdicts2;
  u.ev = c.events + off / 32; u.ev_l1 = c.ev_l1; u.ev_l2 = c.ev_l2; u.ev_bit_base = off;
  u.base = p.base; u.next_anchor = p.next_anchor;
  u.unit = p.unit; u.seg = p.seg; u.gen = 0; u.slot = q;
  u.materialize_all = 0;
  u.keep_margin = c.keep_margin;
  out[q] = u;
}

// ---- placement on the device: between k_walk and k_concat nothing waits for the host ------------------------------------
// what a walker reports about its segment, compact (the descriptors are 200+ bytes apiece)
struct SegOut {
  int64_t end_old, new_len, wrong;
  int32_t status; uint32_t n_rows, n_fast, n_post, n_v2, pad;
  unsigned long long lookups, events, cycles;
};
__global__ void k_seg_gather(const UnitIO* __restrict__ segs, uint32_t n, SegOut* __restrict__ out, unsigned long long* flag) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const UnitIO& u = segs[q];
  SegOut o;
  o.end_old = u.end_old; o.new_len = u.new_len; o.wrong = u.wrong; o.status = u.status; o.n_rows = u.n_rows; o.n_fast = u.n_fast;
  o.n_post = u.n_post; o.n_v2 = u.n_v2; o.pad = 0; o.lookups = u.lookups; o.events = u.events; o.cycles = u.cycles;
  out[q] = o;
  if (u.status) atomicOr(flag, 1ull);
}

// One CTA.  Segments are in (contig, segment) order, slot q = segment q: the offset of a segment's new bases inside its
// new contig is the running sum of the new lengths before it (jb_plan.hpp, place_segments); the contigs' new lengths fall
// out of the same sums.  If any segment failed, or a contig outgrew its region, nothing is placed (every dst_off < 0, the
// old lengths kept): k_concat and the recount behind this launch then do nothing and the host takes the slow path.
struct PlaceArgs {
  const SegOut* so; uint32_t n_segs;
  const uint32_t* seg_unit;      // per segment: its contig
  const uint32_t* unit_first;    // per contig: index of its first segment; [n_units] = n_segs
  const ScanUnit* units_old; const int64_t* caps; uint32_t n_units;
  int64_t* prefix;               // scratch, n_segs + 1
  int64_t* dst_off; uint32_t* slot_unit; ScanUnit* units_new;
  unsigned long long* flag;      // in: bit 0 = a segment failed; out: bit 1 = a contig outgrew its region
};
constexpr int kPlaceThreads = 1024;
__global__ void __launch_bounds__(kPlaceThreads)
k_place(PlaceArgs a) {
  __shared__ int64_t part[kPlaceThreads];
  __shared__ unsigned int bad;
  const uint32_t t = threadIdx.x, n = a.n_segs;
  const uint32_t per = (n + kPlaceThreads - 1) / kPlaceThreads;
  const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
  int64_t sum = 0;
  for (uint32_t q = lo; q < hi; ++q) sum += a.so[q].new_len;
  part[t] = sum;
  if (t == 0) bad = (*a.flag & 1ull) ? 1u : 0u;
  __syncthreads();
  for (int d = 1; d < kPlaceThreads; d <<= 1) {         // inclusive scan of the per-thread sums
    const int64_t add = t >= (uint32_t)d ? part[t - d] : 0;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  int64_t run = part[t] - sum;
  for (uint32_t q = lo; q < hi; ++q) { a.prefix[q] = run; run += a.so[q].new_len; }
  if (t == kPlaceThreads - 1) a.prefix[n] = part[t];
  __syncthreads();
  for (uint32_t u = t; u < a.n_units; u += kPlaceThreads)
    if (a.prefix[a.unit_first[u + 1]] - a.prefix[a.unit_first[u]] > a.caps[u]) atomicOr(&bad, 2u);
  __syncthreads();
  const bool ok = bad == 0;
  for (uint32_t q = t; q < n; q += kPlaceThreads) {
    const uint32_t u = a.seg_unit[q];
    a.dst_off[q] = ok ? (int64_t)a.units_old[u].seq_off + a.prefix[q] - a.prefix[a.unit_first[u]] : -1;
    a.slot_unit[q] = u;
  }
  for (uint32_t u = t; u < a.n_units; u += kPlaceThreads) {
    ScanUnit nu = a.units_old[u];
    if (ok) nu.len = (uint64_t)(a.prefix[a.unit_first[u + 1]] - a.prefix[a.unit_first[u]]);
    a.units_new[u] = nu;
  }
  if (t == 0 && (bad & 2u)) atomicOr(a.flag, 2ull);
}

__global__ void k_init_pool(BridgeScratch* pool, int n, uint32_t* trie, uint32_t trie_cap, uint8_t* patch, uint32_t patch_cap) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  pool[w].trie = trie + (size_t)w * trie_cap; pool[w].trie_cap = trie_cap;
  pool[w].patch = patch + (size_t)w * patch_cap; pool[w].patch_cap = patch_cap;
}

}  // namespace jb

using namespace jb;

// ---------------------------------------------------------------------------------------------
// Work buffers.  Device allocations and pinned host staging are expensive to create (a pinned
// 64 MB buffer alone costs tens of milliseconds), so a finished batch parks its workspace in a
// process-wide cache and the next batch on the same device picks it up (grow-only).
namespace {

template <class T>
struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  int ensure(size_t n) {                         // capacity in elements, grow-only, contents not preserved
    if (n <= cap) return JB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return JB_E_NOMEM; }
    cap = want;
    return JB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {
  T* p = nullptr; size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return JB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    if (cudaMallocHost(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return JB_E_NOMEM; }
    cap = want;
    return JB_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Workspace {
  int device = -1;
  DevBuf<uint8_t> buf[2], stage;
  DevBuf<uint32_t> counts[2], events, weak, events_l1, events_l2, order, tile_first, block_first, slot_unit, vmap;
  DevBuf<Verdict> verdicts;
  DevBuf<Verdict2> verdicts2;
  DevBuf<uint32_t> verdict_unit;
  DevBuf<uint8_t> dry_out;
  DevBuf<uint8_t> okb[2];                   // counts / validity of the current and of the next pass's sequence
  DevBuf<ScanUnit> scan_units_new;
  DevBuf<UnitIO> segs;
  DevBuf<int64_t> dst_off;
  DevBuf<long long> anchors;
  DevBuf<ScanUnit> scan_units;
  DevBuf<SegUnit> seg_units;
  DevBuf<DevRow> rows;
  DevBuf<uint8_t> aux;
  DevBuf<CopyJob> jobs;
  DevBuf<unsigned long long> used;          // [0] rows, [1] aux, [2] probe steps, [3] copy jobs
  DevBuf<unsigned int> next;
  DevBuf<BridgeScratch> pool; DevBuf<uint32_t> trie; DevBuf<uint8_t> patch; DevBuf<uint32_t> busy;
  bool pool_ready = false;
  // grown bridging pool: a few slots with a much larger trie, allocated when a search outgrew a standard slot (kTrieFull)
  DevBuf<BridgeScratch> pool_big; DevBuf<uint32_t> trie_big; DevBuf<uint8_t> patch_big; DevBuf<uint32_t> busy_big;
  uint32_t big_nodes = 0;
  void release_big_pool() { pool_big.release(); trie_big.release(); patch_big.release(); busy_big.release(); big_nodes = 0; }
  PinBuf<uint8_t> h_in;                     // staged input bases
  PinBuf<UnitIO> h_segs;
  PinBuf<long long> h_anchors;
  PinBuf<DevRow> h_rows[2];                 // double-buffered: a host thread reads one log while the next lands
  PinBuf<uint8_t> h_aux[2];
  PinBuf<unsigned long long> h_used;
  // placement on the device (k_seg_gather, k_place) and the pinned staging of everything a pass uploads: the host never
  // hands the driver pageable memory inside the pass loop (such a copy would wait for the stream first)
  DevBuf<SegOut> seg_out; PinBuf<SegOut> h_seg_out;
  DevBuf<SegPlanDev> plan_d; PinBuf<SegPlanDev> h_plan;
  DevBuf<uint32_t> seg_unit, unit_first; DevBuf<int64_t> caps_d, prefix;
  PinBuf<uint8_t> h_geom; PinBuf<uint32_t> h_order, h_seg_unit; PinBuf<int64_t> h_dst_off; PinBuf<unsigned long long> h_probes;
  cudaStream_t s2 = nullptr;                // side stream: fix rows travel while the next pass scans
  cudaEvent_t ev_walk = nullptr, ev_rows = nullptr, ev_concat = nullptr, ev_group = nullptr;
  DevBuf<uint32_t> tile_first_groups;
  ~Workspace() {
    if (device < 0) return;
    cudaSetDevice(device);
    buf[0].release(); buf[1].release(); stage.release(); counts[0].release(); counts[1].release(); okb[0].release(); okb[1].release(); slot_unit.release(); scan_units_new.release(); events.release(); weak.release(); events_l1.release(); events_l2.release(); vmap.release(); verdicts.release(); verdicts2.release(); verdict_unit.release(); dry_out.release(); order.release();
    tile_first.release(); block_first.release(); segs.release(); dst_off.release(); anchors.release(); scan_units.release();
    seg_units.release(); rows.release(); aux.release(); jobs.release(); used.release(); next.release(); pool.release();
    trie.release(); patch.release(); busy.release(); release_big_pool();
    seg_out.release(); h_seg_out.release(); plan_d.release(); h_plan.release(); seg_unit.release(); unit_first.release(); caps_d.release(); prefix.release();
    h_geom.release(); h_order.release(); h_seg_unit.release(); h_dst_off.release(); h_probes.release();
    h_in.release(); h_segs.release(); h_anchors.release(); h_rows[0].release(); h_rows[1].release(); h_aux[0].release(); h_aux[1].release(); h_used.release();
    if (s2) cudaStreamDestroy(s2);
    if (ev_walk) cudaEventDestroy(ev_walk);
    if (ev_rows) cudaEventDestroy(ev_rows);
    if (ev_concat) cudaEventDestroy(ev_concat);
    if (ev_group) cudaEventDestroy(ev_group);
    tile_first_groups.release();
  }
};

std::mutex g_ws_mutex;
std::vector<Workspace*> g_ws_cache;          // at most one parked workspace per device

Workspace* take_workspace(int device) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (size_t i = 0; i < g_ws_cache.size(); ++i)
    if (g_ws_cache[i]->device == device) { Workspace* w = g_ws_cache[i]; g_ws_cache.erase(g_ws_cache.begin() + i); return w; }
  auto w = new Workspace();
  w->device = device;
  return w;
}
void park_workspace(Workspace* w) {
  if (!w) return;
  {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    bool have = false;
    for (Workspace* c : g_ws_cache) have |= c->device == w->device;
    if (!have) { g_ws_cache.push_back(w); return; }
  }
  delete w;
}

// pinned buffers that carry polished sequences back; results own one until jb_result_free
std::vector<std::pair<uint8_t*, size_t>> g_out_pool;
uint8_t* take_out_buffer(size_t n, size_t* cap) {
  {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    for (size_t i = 0; i < g_out_pool.size(); ++i)
      if (g_out_pool[i].second >= n) { uint8_t* p = g_out_pool[i].first; *cap = g_out_pool[i].second; g_out_pool.erase(g_out_pool.begin() + i); return p; }
    for (auto& e : g_out_pool) cudaFreeHost(e.first);     // too small: drop them
    g_out_pool.clear();
  }
  uint8_t* p = nullptr;
  size_t want = n + n / 8 + 4096;
  if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  *cap = want;
  return p;
}
void park_out_buffer(uint8_t* p, size_t cap) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  if (g_out_pool.size() < 2) g_out_pool.push_back({p, cap}); else cudaFreeHost(p);
}

}  // namespace

void jb_release_cached_workspaces(int device) {           // jb_db_close: nothing may outlive the context's owner
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (size_t i = 0; i < g_ws_cache.size();)
    if (g_ws_cache[i]->device == device) { delete g_ws_cache[i]; g_ws_cache.erase(g_ws_cache.begin() + i); } else ++i;
}

struct jb_batch {
  bool use_verdicts = true, use_verdicts2 = true;
  const jb_db* db = nullptr;
  Workspace* ws = nullptr;
  size_t n = 0;
  std::vector<uint64_t> off, cap;      // per-contig region in the flat buffers (bytes; multiples of 32)
  std::vector<int64_t> len;            // current lengths
  uint64_t flat_bytes = 0;
  int cur = 0;                         // buffer holding the current sequences
  int sms = 148;
  std::vector<PassOut> passes;
  int status = JB_OK;
  bool polished = false;
  // the final bases are known once the last FIXING pass has stitched them: their copy to the host starts there, on the
  // side stream, beside the count-only pass (jb_batch_fetch picks it up)
  bool prescanned = false;             // jb_polish: the first pass's k_scan already ran behind the upload
  bool fetch_follows = false;          // jb_polish: the caller wants the bases on the host right after polishing
  int64_t prescan_lookups = 0;
  cudaEvent_t ev_tmp[2] = {nullptr, nullptr};
  uint8_t* early = nullptr; size_t early_cap = 0;
  cudaEvent_t ev_early = nullptr;
  double early_ms_hint = 0;
  ~jb_batch() {
    if (ev_early) { cudaEventSynchronize(ev_early); cudaEventDestroy(ev_early); }
    for (cudaEvent_t e : ev_tmp) if (e) { cudaEventSynchronize(e); cudaEventDestroy(e); }
    if (early) park_out_buffer(early, early_cap);
    for (PassOut& po : passes) RowStore::give(std::move(po.rows));
    park_workspace(ws);
  }
};

struct jb_result {
  int status = JB_OK;
  std::vector<PassOut> passes;
  uint8_t* data = nullptr; size_t data_cap = 0;      // pinned: all sequences back to back
  std::vector<uint64_t> seq_off, seq_len;
  ~jb_result() {
    park_out_buffer(data, data_cap);
    for (PassOut& po : passes) RowStore::give(std::move(po.rows));
  }
};

#define JB_TRY(expr) do { int rc_ = (expr); if (rc_ != JB_OK) return fail(rc_, "out of memory for polishing work buffers"); } while (0)

namespace {
int upload_impl(const jb_db* db, const jb_seq_t* seqs, size_t n, jb_batch** out, bool prescan);
}
// device bytes upload_impl allocates per base: three base buffers, two count and two validity arrays (each over the padded
// regions, x 1.125), the bit masks, verdict arrays, fix rows, copy jobs
constexpr uint64_t kDeviceBytesPerBase = 26;
extern "C" int jb_polish_capacity(const jb_db* db, uint64_t* max_bases) {
  if (!db || !max_bases) return fail(JB_E_ARG, "jb_polish_capacity: null argument");
  JB_CUDA(cudaSetDevice(db->device));
  size_t free_b = 0, total_b = 0;
  JB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t pools = (uint64_t)kPoolSlots * ((uint64_t)kTrieCap * 4 + kPatchCap + sizeof(BridgeScratch)) + (1ull << 30);     // bridging scratch + slack
  *max_bases = free_b > pools ? (free_b - pools) / kDeviceBytesPerBase : 0;
  return JB_OK;
}

extern "C" int jb_batch_upload(const jb_db* db, const jb_seq_t* seqs, size_t n, jb_batch** out) {
  return upload_impl(db, seqs, n, out, false);
}
namespace {
// prescan (jb_polish): the copies go to the side stream in groups of contigs and the first pass's k_scan of a group starts
// as soon as the group has landed, so the PCIe transfer of the next group runs beside it
int upload_impl(const jb_db* db, const jb_seq_t* seqs, size_t n, jb_batch** out, bool prescan) {
  if (!db || (!seqs && n) || !out) return fail(JB_E_ARG, "jb_batch_upload: null argument");
  *out = nullptr;
  if (int rc_ready = jb_check_ready(db)) return rc_ready;
  JB_CUDA(cudaSetDevice(db->device));
  cudaStream_t s = db->stream;
  std::unique_ptr<jb_batch> b(new jb_batch());
  b->db = db; b->n = n;
  b->ws = take_workspace(db->device);
  Workspace& w = *b->ws;
  b->off.resize(n); b->cap.resize(n); b->len.resize(n);
  uint64_t total = 0, total_len = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t L = seqs[i].len;
    uint64_t c = (L + L / 8 + 4096 + 31) & ~31ull;
    b->off[i] = total; b->cap[i] = c; b->len[i] = (int64_t)L;
    total += c; total_len += L;
  }
  b->flat_bytes = total;
  cudaDeviceGetAttribute(&b->sms, cudaDevAttrMultiProcessorCount, db->device);
  if (!w.s2) {
    JB_CUDA(cudaStreamCreateWithFlags(&w.s2, cudaStreamNonBlocking));
    JB_CUDA(cudaEventCreateWithFlags(&w.ev_walk, cudaEventDisableTiming));
    JB_CUDA(cudaEventCreateWithFlags(&w.ev_rows, cudaEventDisableTiming));
    JB_CUDA(cudaEventCreateWithFlags(&w.ev_concat, cudaEventDisableTiming));
  }
  JB_TRY(w.buf[0].ensure(total + 64)); JB_TRY(w.buf[1].ensure(total + 64)); JB_TRY(w.stage.ensure(total + 64));
  for (int q = 0; q < 2; ++q) { JB_TRY(w.counts[q].ensure(total + 64)); JB_TRY(w.okb[q].ensure(total + 64)); }
  JB_TRY(w.scan_units_new.ensure(n + 1));
  JB_TRY(w.events.ensure(total / 32 + 64)); JB_TRY(w.weak.ensure(total / 32 + 64)); JB_TRY(w.events_l1.ensure(total / 32768 + 64)); JB_TRY(w.events_l2.ensure(total / (32768 * 32) + 64));
  JB_TRY(w.vmap.ensure(total / 32 + 64)); JB_TRY(w.verdicts.ensure(total / 64 + 4096));
  JB_TRY(w.verdicts2.ensure(total / 64 + 4096)); JB_TRY(w.verdict_unit.ensure(total / 64 + 4096));
  JB_TRY(w.scan_units.ensure(n + 1)); JB_TRY(w.seg_units.ensure(n + 1)); JB_TRY(w.tile_first.ensure(n + 2)); JB_TRY(w.block_first.ensure(n + 2));
  JB_TRY(w.next.ensure(4)); JB_TRY(w.used.ensure(16)); JB_TRY(w.h_used.ensure(16));
  JB_TRY(w.unit_first.ensure(n + 2)); JB_TRY(w.caps_d.ensure(n + 1));
  JB_TRY(w.h_geom.ensure((n + 2) * (4 + 4 + 4 + sizeof(ScanUnit) + sizeof(SegUnit) + 8) + 64));
  JB_TRY(w.rows.ensure(total_len / 32 + 4096)); JB_TRY(w.aux.ensure(total_len / 4 + (1u << 20)));
  for (int q = 0; q < 2; ++q) { JB_TRY(w.h_rows[q].ensure(total_len / 32 + 4096)); JB_TRY(w.h_aux[q].ensure(total_len / 4 + (1u << 20))); }
  JB_TRY(w.jobs.ensure(total_len / 16 + 4 * n + 65536));
  if (!w.pool_ready) {
    JB_TRY(w.pool.ensure(kPoolSlots)); JB_TRY(w.trie.ensure((size_t)kPoolSlots * kTrieCap)); JB_TRY(w.patch.ensure((size_t)kPoolSlots * kPatchCap));
    JB_TRY(w.busy.ensure(kPoolSlots));
    JB_CUDA(cudaMemsetAsync(w.busy.p, 0, (size_t)kPoolSlots * 4, s));
    k_init_pool<<<(kPoolSlots + 127) / 128, 128, 0, s>>>(w.pool.p, (int)kPoolSlots, w.trie.p, kTrieCap, w.patch.p, kPatchCap);
    JB_CUDA(cudaGetLastError());
    w.pool_ready = true;
  }
  JB_TRY(w.h_in.ensure(std::max<uint64_t>(total_len, 1)));
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1));
  JB_CUDA(cudaEventRecord(e0, s));
  // stage through pinned memory in slices: a few host threads copy slices while the DMA engine drains the
  // ones already staged (one thread saturates at ~9 GB/s, the link takes 3x that)
  // A contig that already lies in page-locked memory (jb_host_alloc, cudaHostAlloc, cudaHostRegister) goes to the device
  // straight from where it is: no staging copy, the link runs at its own speed (~50 GB/s instead of the ~20 GB/s of memcpy + DMA).
  struct Slice { const char* src; uint64_t dev_off, pin_off, len; size_t unit; bool last; bool direct; };
  std::vector<Slice> slices;
  constexpr uint64_t kSlice = 4ull << 20;
  uint64_t p = 0;
  size_t n_staged = 0;
  for (size_t i = 0; i < n; ++i) {
    bool direct = false;
    if (seqs[i].len) {
      cudaPointerAttributes pa{};
      if (cudaPointerGetAttributes(&pa, seqs[i].seq) == cudaSuccess) direct = pa.type == cudaMemoryTypeHost;
      else cudaGetLastError();
    }
    for (uint64_t at = 0; at < seqs[i].len; at += kSlice) {
      const uint64_t m = std::min<uint64_t>(kSlice, seqs[i].len - at);
      slices.push_back(Slice{seqs[i].seq + at, b->off[i] + at, p, m, i, at + m == seqs[i].len, direct});
      if (!direct) { p += m; ++n_staged; }
    }
  }
  note_counter("h2d_direct_bytes", (int64_t)(total_len - p));
  const int n_thr = (int)std::max<size_t>(1, std::min<size_t>(4, n_staged / 2));
  std::vector<std::atomic<int>> ready(slices.size());
  for (auto& r : ready) r.store(0);
  std::vector<std::thread> copiers;
  uint8_t* pin = w.h_in.p;
  // staged slices are dealt to the copier threads round robin by their rank among the staged ones
  std::vector<int> owner(slices.size(), -1);
  { size_t q = 0; for (size_t j = 0; j < slices.size(); ++j) if (!slices[j].direct) owner[j] = (int)(q++ % (size_t)n_thr); }
  for (int t = 1; t < n_thr; ++t)
    copiers.emplace_back([&, t]() {
      for (size_t j = 0; j < slices.size(); ++j) {
        if (owner[j] != t) continue;
        std::memcpy(pin + slices[j].pin_off, slices[j].src, slices[j].len);
        ready[j].store(1, std::memory_order_release);
      }
    });
  struct Joiner { std::vector<std::thread>* t; ~Joiner() { for (auto& x : *t) if (x.joinable()) x.join(); } } joiner{&copiers};
  // groups of whole contigs for the pipelined first scan
  prescan = prescan && total_len > 0 && getenv("JB_NO_PRESCAN") == nullptr;
  std::vector<size_t> group_end;                        // contig index one past each group
  std::vector<ScanUnit> sunits(n);
  std::vector<uint32_t> tf;                             // per group: its own tile prefix, back to back
  std::vector<size_t> tf_at;
  if (prescan) {
    constexpr uint64_t kGroupBytes = 12ull << 20;
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
      sunits[i] = ScanUnit{b->off[i], (uint64_t)b->len[i], b->off[i]};
      acc += seqs[i].len;
      if (acc >= kGroupBytes || i + 1 == n) { group_end.push_back(i + 1); acc = 0; }
    }
    size_t first = 0;
    for (size_t g = 0; g < group_end.size(); ++g) {
      tf_at.push_back(tf.size());
      tf.push_back(0);
      for (size_t i = first; i < group_end[g]; ++i) tf.push_back(tf.back() + tiles_of((uint64_t)b->len[i]));
      first = group_end[g];
    }
    JB_TRY(w.tile_first_groups.ensure(tf.size() + 1));
    JB_CUDA(cudaMemcpyAsync(w.scan_units.p, sunits.data(), n * sizeof(ScanUnit), cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemcpyAsync(w.tile_first_groups.p, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice, s));
    if (!w.ev_group) JB_CUDA(cudaEventCreateWithFlags(&w.ev_group, cudaEventDisableTiming));
    JB_CUDA(cudaEventRecord(w.ev_group, s));            // (the copies must not overtake earlier work on s that still reads buf[0])
    JB_CUDA(cudaStreamWaitEvent(w.s2, w.ev_group, 0));
  }
  cudaStream_t cs = prescan ? w.s2 : s;
  if (prescan) JB_CUDA(cudaEventRecord(e0, cs));
  const TableView view = db->view();
  cudaError_t copy_err = cudaSuccess;
  size_t g = 0, g_first = 0;
  for (size_t j = 0; j < slices.size(); ++j) {
    if (!slices[j].direct) {
      if (owner[j] == 0) std::memcpy(pin + slices[j].pin_off, slices[j].src, slices[j].len);
      else while (!ready[j].load(std::memory_order_acquire)) std::this_thread::yield();
    }
    if (copy_err == cudaSuccess)
      copy_err = cudaMemcpyAsync(w.buf[0].p + slices[j].dev_off, slices[j].direct ? (const uint8_t*)slices[j].src : pin + slices[j].pin_off, slices[j].len,
                                 cudaMemcpyHostToDevice, cs);
    if (!prescan) continue;
    if (!slices[j].last) continue;
    size_t done_units = slices[j].unit + 1;               // contigs fully enqueued (empty ones have no slices)
    while (done_units < n && seqs[done_units].len == 0) ++done_units;
    while (g < group_end.size() && group_end[g] <= done_units && copy_err == cudaSuccess) {
      JB_CUDA(cudaEventRecord(w.ev_group, cs));
      JB_CUDA(cudaStreamWaitEvent(s, w.ev_group, 0));
      ScanArgs a{};
      a.bases = w.buf[0].p; a.units = w.scan_units.p + g_first; a.tile_first = w.tile_first_groups.p + tf_at[g];
      a.n_units = (uint32_t)(group_end[g] - g_first); a.n_tiles = tf[tf_at[g] + (group_end[g] - g_first)];
      a.counts = w.counts[0].p; a.okb = w.okb[0].p; a.probe_steps = nullptr;
      if (int rc = launch_scan(view, a, s)) return rc;
      g_first = group_end[g];
      ++g;
    }
  }
  for (auto& t : copiers) t.join();
  JB_CUDA(copy_err);
  JB_CUDA(cudaEventRecord(e1, cs));
  if (prescan) {
    b->prescanned = true;
    b->prescan_lookups = 0;
    for (size_t i = 0; i < n; ++i) b->prescan_lookups += b->len[i] + 1;
    // no wait here: jb_batch_polish queues behind the scans on s; the pinned staging area belongs to the workspace
    note_timing("h2d", -1.0);
  } else {
    JB_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    note_timing("h2d", ms);
  }
  note_counter("h2d_bytes", (int64_t)total_len);
  if (!prescan) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  else { b->ev_tmp[0] = e0; b->ev_tmp[1] = e1; }
  *out = b.release();
  return JB_OK;
}
}  // namespace

extern "C" void jb_batch_free(jb_batch* b) { delete b; }

// page-locked host memory for callers that want their contigs to travel without a staging copy
extern "C" int jb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(JB_E_ARG, "jb_host_alloc: null argument");
  *out = nullptr;
  if (cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return fail(JB_E_NOMEM, "cannot page-lock " + std::to_string(bytes) + " bytes of host memory");
  }
  return JB_OK;
}
extern "C" void jb_host_free(void* p) { if (p) cudaFreeHost(p); }

namespace {

// dst: pinned host memory (mapped), src: device memory, both 16-byte aligned
int small_d2h(void* dst, const void* src, size_t n, bool by_kernel, cudaStream_t s) {
  if (n == 0) return JB_OK;
  if (!by_kernel) { JB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, s)); return JB_OK; }
  const unsigned blocks = (unsigned)std::min<size_t>(64, (n / 16 + 255) / 256 + 1);
  k_store_host<<<blocks, 256, 0, s>>>(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n);
  JB_CUDA(cudaGetLastError());
  return JB_OK;
}

// The grown bridging pool: kBigPoolSlots slots whose tries take up to half of the free device memory (at most 2^30 - 1 nodes
// each: a node id occupies 30 bits of its child's entry).  Released when the batch has been polished.
constexpr uint32_t kBigPoolSlots = 4;
int ensure_big_pool(Workspace& w, cudaStream_t s) {
  if (w.big_nodes) return JB_OK;
  size_t free_b = 0, total_b = 0;
  JB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  uint64_t nodes = std::min<uint64_t>((1ull << 30) - 1, free_b / 2 / ((uint64_t)kBigPoolSlots * 4));
  if (const char* e = getenv("JB_BIG_TRIE_NODES")) nodes = std::min<uint64_t>(nodes, (uint64_t)atoll(e));    // tests: exhaust the grown pool too
  if (nodes <= kTrieCap) return fail(JB_E_NOMEM, "no device memory left to grow the de-Bruijn path pool");
  if (w.pool_big.ensure(kBigPoolSlots) || w.trie_big.ensure((size_t)kBigPoolSlots * nodes) || w.patch_big.ensure((size_t)kBigPoolSlots * kPatchCap) ||
      w.busy_big.ensure(kBigPoolSlots)) {
    w.release_big_pool();
    return fail(JB_E_NOMEM, "out of device memory while growing the de-Bruijn path pool");
  }
  JB_CUDA(cudaMemsetAsync(w.busy_big.p, 0, (size_t)kBigPoolSlots * 4, s));
  k_init_pool<<<1, 128, 0, s>>>(w.pool_big.p, (int)kBigPoolSlots, w.trie_big.p, (uint32_t)nodes, w.patch_big.p, kPatchCap);
  JB_CUDA(cudaGetLastError());
  w.big_nodes = (uint32_t)nodes;
  add_counter("pool_growths", 1);
  return JB_OK;
}

// dst: device memory, src: pinned host memory (mapped), both 4-byte aligned, n a multiple of 4
int small_h2d(void* dst, const void* src, size_t n, bool by_kernel, cudaStream_t s) {
  if (n == 0) return JB_OK;
  if (!by_kernel || (n & 3)) { JB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s)); return JB_OK; }
  const unsigned blocks = (unsigned)std::min<size_t>(64, (n / 4 + 255) / 256);
  k_load_host<<<blocks, 256, 0, s>>>(static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(src), n / 4);
  JB_CUDA(cudaGetLastError());
  return JB_OK;
}

int64_t pick_seg_block(int64_t total_len, int sms) {
  // (multiples of 32: k_find_anchors walks whole bitmask words)
  if (const char* e = getenv("JB_SEG_BLOCK")) { long v = atol(e); if (v >= 64) return (v + 31) / 32 * 32; if (v == 0) return 0; }
  // enough segments to fill the resident walkers a few times over, but not so short that seams dominate
  const int64_t resident = (int64_t)sms * kWalkCtasPerSm * kWalkWarps;
  int64_t blk = total_len / (4 * resident);
  blk = std::max<int64_t>(1024, std::min<int64_t>(8192, blk));
  return (blk + 31) / 32 * 32;
}

UnitIO make_seg(const jb_batch* b, const SegPlan& p, uint8_t* in, uint32_t gen, uint32_t slot, int keep_margin, int cidx) {
  UnitIO u;                                        // (result fields are written by the walker)
  const size_t i = p.unit;
  const Workspace& w = *b->ws;
  u.status = 0; u.n_rows = 0; u.new_len = 0; u.end_old = 0; u.wrong = 0; u.lookups = 0; u.events = 0; u.n_fast = 0; u.n_post = 0; u.n_v2 = 0;
  u.in = in + b->off[i]; u.in_len = b->len[i];
  u.out = w.stage.p + b->off[i] + p.stage_off; u.out_cap = p.stage_cap;
  u.cnt = w.counts[cidx].p + b->off[i];
  u.okb = w.okb[cidx].p + b->off[i];
  u.weak = w.weak.p + b->off[i] / 32; u.vmap = b->use_verdicts ? w.vmap.p : nullptr; u.verdicts = w.verdicts.p;
  u.verdicts2 = b->use_verdicts2 ? w.verdicts2.p : nullptr;
  u.ev = w.events.p + b->off[i] / 32; u.ev_l1 = w.events_l1.p; u.ev_l2 = getenv("JB_NO_L2") ? nullptr : w.events_l2.p; u.ev_bit_base = b->off[i];
  u.base = p.base; u.next_anchor = p.next_anchor;
  u.unit = p.unit; u.seg = p.seg; u.gen = gen; u.slot = slot;
  u.materialize_all = gen ? 1u : 0u;               // re-runs take the conservative path
  u.keep_margin = keep_margin;
  return u;
}

// The fix log of one pass waits in pinned memory until its copy has landed (while the next pass scans);
// turning it into csv rows (sort, alignment of multi-error patches, text) then runs on a host thread
// beside the GPU work of the following passes.
struct PendingRows {
  bool live = false;
  int buf = 0;               // which pinned log
  size_t pass = 0;
  unsigned long long n_rows = 0, n_aux = 0;
  std::vector<UnitIO> segs;
  std::vector<SegPlace> place;
  std::vector<size_t> first_seg;
};

struct RowJob {
  std::thread th;
  std::vector<UnitIO> segs;
  std::vector<SegPlace> place;
  std::vector<size_t> first_seg;
  PassOut out;
  bool ok = true;
  bool nomem = false;        // the host-side alignment of a bridged patch ran out of memory
  size_t pass = 0;
  int buf = 0;
  std::atomic<bool> copied{true};
};

int finish_rows(jb_batch* b, PendingRows* pr, std::vector<std::unique_ptr<RowJob>>* jobs) {
  if (!pr->live) return JB_OK;
  Workspace& w = *b->ws;
  JB_CUDA(cudaEventSynchronize(w.ev_rows));
  std::unique_ptr<RowJob> j(new RowJob());
  j->segs.swap(pr->segs); j->place.swap(pr->place); j->first_seg.swap(pr->first_seg);
  j->pass = pr->pass;
  RowJob* raw = j.get();
  const DevRow* src_rows = w.h_rows[pr->buf].p; const uint8_t* src_aux = w.h_aux[pr->buf].p;
  raw->buf = pr->buf;
  const size_t n_rows = pr->n_rows, n_aux = pr->n_aux;
  raw->copied.store(false);
  (void)n_aux;
  const bool trace = getenv("JB_TRACE") != nullptr;
  j->th = std::thread([raw, src_rows, src_aux, n_rows, trace]() {
    // straight from the pinned log (the next pass's log is not copied into it before `copied` is set)
    const auto t0 = std::chrono::steady_clock::now();
    // (an exception must not leave this thread: std::terminate would take the host process down)
    try { raw->ok = rows_in_order(src_rows, n_rows, src_aux, raw->segs, raw->place, raw->first_seg, &raw->out); }
    catch (const std::exception&) { raw->ok = false; raw->nomem = true; }
    raw->copied.store(true, std::memory_order_release);
    if (trace) fprintf(stderr, "[jb trace] row thread: %zu log rows -> %zu csv rows in %.3f ms\n", n_rows, raw->out.rows.size(),
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  });
  jobs->push_back(std::move(j));
  pr->live = false;
  return JB_OK;
}

void join_rows(jb_batch* b, std::vector<std::unique_ptr<RowJob>>* jobs) {
  for (auto& j : *jobs) {
    if (j->th.joinable()) j->th.join();
    if (j->pass < b->passes.size()) { b->passes[j->pass].rows.swap(j->out.rows); b->passes[j->pass].spill.swap(j->out.spill); }
    if (!j->ok) b->status = j->nomem ? JB_E_NOMEM : JB_E_REFERENCE_CRASH;          // (the latter: exit 1 in the reference)
  }
  jobs->clear();
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
  Workspace& w = *b->ws;
  const size_t n = b->n;
  // events: destroyed on every way out; one set of timing events per pass (nothing is read before the last pass has run:
  // the host does not wait for the stream between passes)
  struct EventBag {
    std::vector<cudaEvent_t> all;
    ~EventBag() { for (cudaEvent_t e : all) cudaEventDestroy(e); }
    cudaError_t make(cudaEvent_t* e, unsigned flags = cudaEventDefault) {
      cudaError_t rc = cudaEventCreateWithFlags(e, flags);
      if (rc == cudaSuccess) all.push_back(*e);
      return rc;
    }
  } evs;
  struct PassEvents { cudaEvent_t a, b, c, d, e, r0, r1, v0, v1; bool recount = false, verdicts = false; };
  std::vector<PassEvents> pev((size_t)passes + 1);
  for (PassEvents& pe : pev)
    for (cudaEvent_t* e : {&pe.a, &pe.b, &pe.c, &pe.d, &pe.e, &pe.r0, &pe.r1, &pe.v0, &pe.v1}) JB_CUDA(evs.make(e));
  cudaEvent_t ev_begin, ev_end, ev_anchors, ev_res;
  JB_CUDA(evs.make(&ev_begin)); JB_CUDA(evs.make(&ev_end));
  JB_CUDA(evs.make(&ev_anchors, cudaEventDisableTiming)); JB_CUDA(evs.make(&ev_res, cudaEventDisableTiming));
  double scan_ms = 0, walk_ms = 0, plan_ms = 0, kwalk_ms = 0;
  int64_t scan_lookups = 0, walk_lookups = 0, launches = 0, n_events = 0, probe_steps = 0, n_segments = 0, n_reruns = 0;
  int64_t max_seg_events = 0, max_seg_lookups = 0, max_seg_span = 0, n_fast = 0, n_post = 0, n_v2 = 0, n_speculative = 0;
  int64_t slow_cycles[3] = {0, 0, 0}, slow_events[3] = {0, 0, 0}, slow_lookups[3] = {0, 0, 0}, slow_span[3] = {0, 0, 0}, sum_cycles[3] = {0, 0, 0};
  std::vector<int64_t> caps(n);
  const WalkCfg cfg{k, sampling_step(k), solid_thre, 0};
  const TableView view = db->view();
  const BridgePool pool{w.pool.p, kPoolSlots, w.busy.p};
  int64_t total_len = 0;
  for (size_t i = 0; i < n; ++i) { caps[i] = (int64_t)b->cap[i]; total_len += b->len[i]; }
  const int64_t seg_block = pick_seg_block(total_len, b->sms);
  int64_t margin = anchor_margin(k);
  if (const char* e = getenv("JB_ANCHOR_MARGIN")) margin = std::max<long>(0, atol(e));   // tests: provoke seam re-runs
  int keep_margin = default_keep_margin(k);
  if (const char* e = getenv("JB_KEEP_MARGIN")) keep_margin = (int)std::max<long>(0, atol(e));   // tests: provoke far-back re-runs
  int cidx = 0;                        // which counts / validity buffer describes the current sequence
  bool have_counts = false;
  int64_t full_scans = 0, full_scan_lookups = 0, recount_probes = 0;
  double recount_ms = 0;
  const bool full_rescan = getenv("JB_FULL_RESCAN") != nullptr;      // tests: scan every pass from scratch
  const bool verify_recount = getenv("JB_VERIFY_RECOUNT") != nullptr; // tests: check k_recount against a full scan
  // tests: JB_NO_SPECULATION makes the host place the segments and launch k_concat / the recount itself, as the slow path does
  const bool speculate = getenv("JB_NO_SPECULATION") == nullptr && !verify_recount;
  const bool do_fix_verdicts = getenv("JB_NO_VERDICTS") == nullptr;      // tests: walkers decide everything themselves
  double verdict_ms = 0;
  int64_t n_verdicts = 0;
  b->passes.clear();
  b->passes.reserve((size_t)passes + 1);
  const bool trace = getenv("JB_TRACE") != nullptr;
  const auto t_origin = std::chrono::steady_clock::now();
  auto mark = [&](const char* what, int ite) {
    if (trace) fprintf(stderr, "[jb trace] pass %d %-18s %9.3f ms\n", ite, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_origin).count());
  };
  PendingRows pending;
  std::vector<std::unique_ptr<RowJob>> row_jobs;
  struct Joiner { jb_batch* b; std::vector<std::unique_ptr<RowJob>>* j; ~Joiner() { join_rows(b, j); } } joiner{b, &row_jobs};   // also on error returns
  unsigned long long* used = w.h_used.p;
  JB_CUDA(cudaEventRecord(ev_begin, s));
  for (int ite = 0; ite <= passes && n > 0; ++ite) {
    const bool do_fix = fix && ite != passes;
    b->passes.push_back(PassOut{});
    uint8_t* in = w.buf[b->cur].p;
    uint8_t* outb = w.buf[b->cur ^ 1].p;
    PassEvents& pe = pev[(size_t)ite];
    // the pass's per-contig geometry, staged in pinned memory (the copies below run when the stream gets there; the area is
    // written again only after this pass's ev_res has been waited for)
    uint8_t* geom = w.h_geom.p;
    int64_t* caps_h = reinterpret_cast<int64_t*>(geom); geom += (n + 1) * 8;          // (the 8-byte aligned arrays first)
    ScanUnit* sunits = reinterpret_cast<ScanUnit*>(geom); geom += (n + 1) * sizeof(ScanUnit);
    SegUnit* gunits = reinterpret_cast<SegUnit*>(geom); geom += (n + 1) * sizeof(SegUnit);
    uint32_t* tile_first = reinterpret_cast<uint32_t*>(geom); geom += (n + 2) * 4;
    uint32_t* block_first = reinterpret_cast<uint32_t*>(geom); geom += (n + 2) * 4;
    uint32_t* unit_first = reinterpret_cast<uint32_t*>(geom);
    tile_first[0] = 0; block_first[0] = 0;
    int64_t pass_total = 0, pass_lookups = 0;
    for (size_t i = 0; i < n; ++i) {
      caps_h[i] = caps[i];
      pass_total += b->len[i] - k + 1;                                  // may be negative (src/jasper.py:51)
      sunits[i] = ScanUnit{b->off[i], (uint64_t)b->len[i], b->off[i]};
      gunits[i] = SegUnit{b->off[i], (uint64_t)b->len[i]};
      pass_lookups += b->len[i] + 1;
      tile_first[i + 1] = tile_first[i] + tiles_of((uint64_t)b->len[i]);
      block_first[i + 1] = block_first[i] + (uint32_t)(seg_block > 0 && b->len[i] > 0 ? (b->len[i] - 1) / seg_block : 0);
    }
    b->passes[ite].total = pass_total;
    // ---- scan: counts -> event / weak bits
    std::vector<BlockRef> blocks;
    if (seg_block > 0) blocks = make_blocks(b->len, seg_block);
    JB_TRY(w.anchors.ensure(blocks.size() + 1)); JB_TRY(w.h_anchors.ensure(blocks.size() + 1));
    // (while the polished bases travel on the side stream the copy engine is taken: a 64 MB transfer in front of these few
    //  bytes would hold the whole count pass back, so a kernel fetches them from the mapped staging area instead)
    const bool up_by_kernel = b->early != nullptr;
    if (int rc = small_h2d(w.tile_first.p, tile_first, (n + 1) * 4, up_by_kernel, s)) return rc;
    if (int rc = small_h2d(w.block_first.p, block_first, (n + 1) * 4, up_by_kernel, s)) return rc;
    if (int rc = small_h2d(w.scan_units.p, sunits, n * sizeof(ScanUnit), up_by_kernel, s)) return rc;
    if (int rc = small_h2d(w.seg_units.p, gunits, n * sizeof(SegUnit), up_by_kernel, s)) return rc;
    if (ite == 0) JB_CUDA(cudaMemcpyAsync(w.caps_d.p, caps_h, n * 8, cudaMemcpyHostToDevice, s));
    JB_CUDA(cudaMemsetAsync(w.next.p, 0, 4, s));
    JB_CUDA(cudaMemsetAsync(w.used.p, 0, 128, s));       // (the previous pass's counters are on the host already, or on their way)
    JB_CUDA(cudaMemsetAsync(w.events_l1.p, 0, (b->flat_bytes / 32768 + 64) * 4, s));
    JB_CUDA(cudaMemsetAsync(w.events_l2.p, 0, (b->flat_bytes / (32768 * 32) + 64) * 4, s));
    ScanArgs a{};
    a.bases = in; a.units = w.scan_units.p; a.tile_first = w.tile_first.p; a.n_units = (uint32_t)n; a.n_tiles = tile_first[n];
    a.counts = w.counts[cidx].p; a.okb = w.okb[cidx].p; a.probe_steps = w.used.p + 2;
    JB_CUDA(cudaEventRecord(pe.a, s));
    if (ite == 0 && b->prescanned && !have_counts) {      // jb_polish: scanned group by group behind the upload
      have_counts = true;
      ++full_scans; full_scan_lookups += b->prescan_lookups;
    }
    if (!have_counts) {                                   // later passes inherit counts from k_recount
      if (int rc = launch_scan(view, a, s)) return rc;
      ++launches; ++full_scans;
      full_scan_lookups += pass_lookups;
    }
    JB_CUDA(cudaEventRecord(pe.b, s));
    EventArgs ea{};
    ea.units = w.scan_units.p; ea.tile_first = w.tile_first.p; ea.n_units = (uint32_t)n; ea.n_tiles = tile_first[n];
    ea.counts = w.counts[cidx].p; ea.okb = w.okb[cidx].p; ea.events = w.events.p; ea.weak = w.weak.p; ea.events_l1 = w.events_l1.p; ea.events_l2 = w.events_l2.p;
    ea.solid_thre = solid_thre; ea.k = k;
    if (int rc = launch_events(ea, s)) return rc;
    ++launches;
    // ---- anchors -> segments
    long long* anchors = w.h_anchors.p;
    if (!blocks.empty()) {
      const uint32_t nb = (uint32_t)blocks.size();
      k_find_anchors<<<(nb + 7) / 8, 256, 0, s>>>(w.events.p, w.weak.p, w.seg_units.p, w.block_first.p, (uint32_t)n, nb, k, (int)seg_block,
                                                   (int)margin, anchors);        // (straight into mapped pinned memory: no copy to queue)
      JB_CUDA(cudaGetLastError());
      ++launches;
    }
    JB_CUDA(cudaEventRecord(ev_anchors, s));             // the host plans segments from here on, while the verdict kernels run
    // ---- verdicts: every isolated error decided in bulk (jb_verdict.cuh)
    b->use_verdicts = do_fix_verdicts;
    b->use_verdicts2 = false;
    if (b->use_verdicts) {
      JB_CUDA(cudaMemsetAsync(w.vmap.p, 0, (b->flat_bytes / 32 + 64) * 4, s));
      JB_CUDA(cudaMemsetAsync(w.used.p + 6, 0, 8, s));
      VerdictArgs va{};
      va.bases = in; va.units = w.scan_units.p; va.tile_first = w.tile_first.p; va.n_units = (uint32_t)n; va.n_tiles = tile_first[n];
      va.counts = w.counts[cidx].p; va.okb = w.okb[cidx].p; va.weak = w.weak.p;
      va.verdicts = w.verdicts.p; va.cap = std::min(w.verdicts.cap, std::min(w.verdicts2.cap, w.verdict_unit.cap)); va.used = w.used.p + 6; va.vmap = w.vmap.p;
      va.verdict_unit = w.verdict_unit.p; va.verdicts2 = w.verdicts2.p;
      JB_TRY(w.dry_out.ensure((size_t)b->sms * kWalkCtasPerSm * kWalkWarps * kDryOut));
      va.next = w.used.p + 7;
      va.dry_out = w.dry_out.p; va.dry_post = getenv("JB_NO_V2_POST") ? 0 : (getenv("JB_NO_POST2") ? 2 : 1);
      va.k = k; va.step = cfg.step; va.solid = solid_thre;
      const unsigned vgrid = (unsigned)std::min<uint64_t>((tile_first[n] + kVerdictWarps - 1) / kVerdictWarps, (uint64_t)b->sms * 8);
      JB_CUDA(cudaEventRecord(pe.v0, s));
      k_verdicts<<<std::max(1u, vgrid), kVerdictWarps * 32, 0, s>>>(view, va);
      JB_CUDA(cudaGetLastError());
      ++launches;
      b->use_verdicts2 = do_fix && getenv("JB_NO_VERDICTS2") == nullptr;   // (a count-only pass applies nothing)
      if (b->use_verdicts2) {
        k_verdicts2<<<(unsigned)b->sms * kWalkCtasPerSm, kWalkWarps * 32, 0, s>>>(view, va, pool);
        JB_CUDA(cudaGetLastError());
        ++launches;
      }
      JB_CUDA(cudaEventRecord(pe.v1, s));
      pe.verdicts = true;
    }
    // while the scan runs: turn the previous pass's fix log into csv rows
    mark("scan launched", ite);
    if (int rc = finish_rows(b, &pending, &row_jobs)) return rc;
    mark("rows handed off", ite);
    JB_CUDA(cudaEventSynchronize(ev_anchors));
    mark("scan+anchors done", ite);
    std::vector<SegPlan> plan = plan_segments(b->len, caps, blocks, blocks.empty() ? nullptr : anchors);
    const size_t n_first = plan.size();
    mark("  planned", ite);
    JB_TRY(w.segs.ensure(n_first + 2 * n)); JB_TRY(w.h_segs.ensure(n_first + 2 * n));     // room for two rounds of whole-contig re-runs behind the plan
    JB_TRY(w.order.ensure(n_first + 2 * n)); JB_TRY(w.dst_off.ensure(n_first + 2 * n)); JB_TRY(w.h_dst_off.ensure(n_first + 2 * n));
    JB_TRY(w.h_order.ensure(n_first + 2 * n)); JB_TRY(w.seg_unit.ensure(n_first + 1)); JB_TRY(w.h_seg_unit.ensure(n_first + 1));
    JB_TRY(w.seg_out.ensure(n_first + 1)); JB_TRY(w.h_seg_out.ensure(n_first + 1)); JB_TRY(w.prefix.ensure(n_first + 2));
    JB_TRY(w.slot_unit.ensure(n_first + 2 * n + 1)); JB_TRY(w.h_probes.ensure((size_t)passes + 2));
    JB_TRY(w.plan_d.ensure(n_first + 1)); JB_TRY(w.h_plan.ensure(n_first + 1));
    UnitIO* hs = w.h_segs.p;
    uint32_t* seg_unit = w.h_seg_unit.p;
    SegPlanDev* hp = w.h_plan.p;
    for (size_t q = 0; q < n_first; ++q) {
      // the device builds the descriptor (k_make_segs); the host keeps the handful of fields it reads again later
      const SegPlan& p = plan[q];
      hp[q] = SegPlanDev{p.unit, p.seg, p.base, p.next_anchor, p.stage_off, p.stage_cap};
      UnitIO& u = hs[q];
      u.unit = p.unit; u.seg = p.seg; u.gen = 0; u.slot = (uint32_t)q; u.base = p.base; u.next_anchor = p.next_anchor; u.in_len = b->len[p.unit];
      seg_unit[q] = p.unit;
      if (p.seg == 0) unit_first[p.unit] = (uint32_t)q;
    }
    unit_first[n] = (uint32_t)n_first;
    mark("  descriptors", ite);
    // queue order: the few long segments (stretches without an anchor) first, the rest as planned
    uint32_t* order = w.h_order.p;
    {
      auto seg_span = [&](size_t q) { return (hs[q].next_anchor >= 0 ? hs[q].next_anchor : hs[q].in_len) - hs[q].base; };
      const int64_t long_span = 4 * std::max<int64_t>(seg_block, 1024);
      std::vector<uint32_t> longs;
      for (size_t q = 0; q < n_first; ++q) if (seg_span(q) > long_span) longs.push_back((uint32_t)q);
      std::sort(longs.begin(), longs.end(), [&](uint32_t x, uint32_t y) { return seg_span(x) > seg_span(y); });
      size_t at = 0;
      for (uint32_t q : longs) order[at++] = q;
      for (size_t q = 0; q < n_first; ++q) if (seg_span(q) <= long_span) order[at++] = (uint32_t)q;
    }
    if (int rc = small_h2d(w.plan_d.p, hp, n_first * sizeof(SegPlanDev), up_by_kernel, s)) return rc;
    {
      SegConsts sc{};
      sc.in = in; sc.stage = w.stage.p; sc.counts = w.counts[cidx].p; sc.okb = w.okb[cidx].p; sc.weak = w.weak.p;
      sc.vmap = b->use_verdicts ? w.vmap.p : nullptr; sc.verdicts = w.verdicts.p; sc.verdicts2 = b->use_verdicts2 ? w.verdicts2.p : nullptr;
      sc.events = w.events.p; sc.ev_l1 = w.events_l1.p; sc.ev_l2 = getenv("JB_NO_L2") ? nullptr : w.events_l2.p;
      sc.units = w.scan_units.p; sc.keep_margin = keep_margin;
      k_make_segs<<<(unsigned)((n_first + 255) / 256), 256, 0, s>>>(w.plan_d.p, (uint32_t)n_first, sc, w.segs.p);
      JB_CUDA(cudaGetLastError());
      ++launches;
    }
    if (int rc = small_h2d(w.order.p, order, n_first * 4, up_by_kernel, s)) return rc;
    if (do_fix) {
      JB_CUDA(cudaMemcpyAsync(w.seg_unit.p, seg_unit, n_first * 4, cudaMemcpyHostToDevice, s));
      JB_CUDA(cudaMemcpyAsync(w.unit_first.p, unit_first, (n + 1) * 4, cudaMemcpyHostToDevice, s));
    }
    // ---- walk (the previous pass's rows have left the arena: the stream waited for ev_rows)
    WalkCfg c2 = cfg; c2.fix = do_fix ? 1 : 0;
    RowSink sink{w.rows.p, w.rows.cap, w.used.p, w.aux.p, w.aux.cap, w.used.p + 1};
    JobSink jobs{w.jobs.p, w.jobs.cap, w.used.p + 3};
    auto walk_grid = [&](size_t cnt) { return (unsigned)std::min<size_t>((cnt + kWalkWarps - 1) / kWalkWarps, (size_t)b->sms * kWalkCtasPerSm); };
    JB_CUDA(cudaEventRecord(pe.c, s));
    k_walk<<<walk_grid(n_first), kWalkWarps * 32, 0, s>>>(view, w.segs.p, w.order.p, (uint32_t)n_first, c2, pool, sink, jobs, w.next.p);
    JB_CUDA(cudaGetLastError());
    ++launches;
    JB_CUDA(cudaEventRecord(pe.e, s));
    JB_CUDA(cudaEventRecord(w.ev_walk, s));              // the fix log may leave on the side stream from here on
    // ---- the segments' results, compact; placement of their new bases on the device; then -- without waiting for the
    // host -- the stitching and the recount.  The host reads the results meanwhile; if a segment failed (k_place then placed
    // nothing and the kernels behind it found nothing to do) it re-walks and launches them again itself.
    k_seg_gather<<<(unsigned)((n_first + 255) / 256), 256, 0, s>>>(w.segs.p, (uint32_t)n_first, w.seg_out.p, w.used.p + 8);
    JB_CUDA(cudaGetLastError());
    ++launches;
    if (do_fix) {
      PlaceArgs pa{};
      pa.so = w.seg_out.p; pa.n_segs = (uint32_t)n_first; pa.seg_unit = w.seg_unit.p; pa.unit_first = w.unit_first.p;
      pa.units_old = w.scan_units.p; pa.caps = w.caps_d.p; pa.n_units = (uint32_t)n; pa.prefix = w.prefix.p;
      pa.dst_off = w.dst_off.p; pa.slot_unit = w.slot_unit.p; pa.units_new = w.scan_units_new.p; pa.flag = w.used.p + 8;
      k_place<<<1, kPlaceThreads, 0, s>>>(pa);
      JB_CUDA(cudaGetLastError());
      ++launches;
    }
    const bool engine_busy = b->early != nullptr;        // the polished bases are travelling on the side stream
    SegOut* so = w.h_seg_out.p;
    if (int rc = small_d2h(so, w.seg_out.p, n_first * sizeof(SegOut), engine_busy, s)) return rc;
    if (int rc = small_d2h(used, w.used.p, 128, engine_busy, s)) return rc;
    JB_CUDA(cudaEventRecord(ev_res, s));
    const bool keep_counts = !full_rescan && ite + 1 <= passes;
    // k_concat, then the counts of the new sequence: copied for untouched stretches, probed again around the edits
    auto stitch_and_recount = [&]() -> int {
      k_concat<<<(unsigned)b->sms * 8, 256, 0, s>>>(w.jobs.p, w.used.p + 3, (unsigned long long)w.jobs.cap, w.dst_off.p, outb);
      JB_CUDA(cudaGetLastError());
      ++launches;
      JB_CUDA(cudaEventRecord(w.ev_concat, s));
      if (!keep_counts) return JB_OK;
      RecountArgs ra{};
      ra.jobs = w.jobs.p; ra.n_jobs = w.used.p + 3; ra.jobs_cap = (unsigned long long)w.jobs.cap; ra.dst_off = w.dst_off.p; ra.slot_unit = w.slot_unit.p;
      ra.units_new = w.scan_units_new.p;
      ra.in_buf = in; ra.flat_bytes = b->flat_bytes; ra.new_bases = outb;
      ra.cnt_old = w.counts[cidx].p; ra.okb_old = w.okb[cidx].p; ra.cnt_new = w.counts[cidx ^ 1].p; ra.okb_new = w.okb[cidx ^ 1].p;
      ra.probes = w.used.p + 4;
      RecountTailArgs rt{w.scan_units.p, w.scan_units_new.p, (uint32_t)n, w.counts[cidx].p, w.counts[cidx ^ 1].p, w.okb[cidx ^ 1].p};
      JB_CUDA(cudaMemsetAsync(w.used.p + 4, 0, 8, s));
      JB_CUDA(cudaEventRecord(pe.r0, s));
      if (int rc = launch_recount(view, ra, rt, s)) return rc;
      JB_CUDA(cudaEventRecord(pe.r1, s));
      pe.recount = true;
      launches += 3;
      // (by kernel, never by the copy engine: behind the last fixing pass the polished bases start travelling on the side
      //  stream, and a device-to-host copy on this stream would queue behind them and hold the stream for the whole transfer)
      w.h_probes.p[ite] = 0;
      return small_d2h(w.h_probes.p + ite, w.used.p + 4, 8, true, s);
    };
    if (do_fix && speculate) { if (int rc = stitch_and_recount()) return rc; }
    mark("walk launched", ite);
    JB_CUDA(cudaEventSynchronize(ev_res));
    mark("walk done", ite);
    n_segments += (int64_t)n_first;
    mark("  results in", ite);
    // the fix log leaves for the host right away, on the side stream (its row thread then starts while this thread still
    // merges the results); a re-walk appends rows and sends it again
    const int hb = ite & 1;                              // the log of pass ite-2 must have been read out of this buffer
    auto send_fix_log = [&]() -> int {
      for (auto& rj : row_jobs) if (rj->buf == hb) while (!rj->copied.load(std::memory_order_acquire)) std::this_thread::yield();
      JB_CUDA(cudaStreamWaitEvent(w.s2, w.ev_walk, 0));
      if (used[0]) JB_CUDA(cudaMemcpyAsync(w.h_rows[hb].p, w.rows.p, used[0] * sizeof(DevRow), cudaMemcpyDeviceToHost, w.s2));
      if (used[1]) JB_CUDA(cudaMemcpyAsync(w.h_aux[hb].p, w.aux.p, used[1], cudaMemcpyDeviceToHost, w.s2));
      JB_CUDA(cudaEventRecord(w.ev_rows, w.s2));
      return JB_OK;
    };
    const bool log_fits = used[0] <= w.rows.cap && used[1] <= w.aux.cap;
    if (do_fix && log_fits) { if (int rc = send_fix_log()) return rc; }
    for (size_t q = 0; q < n_first; ++q) {               // the results join the descriptors the host built
      UnitIO& u = hs[q];
      const SegOut& o = so[q];
      u.end_old = o.end_old; u.new_len = o.new_len; u.wrong = o.wrong; u.status = o.status; u.n_rows = o.n_rows; u.n_fast = o.n_fast;
      u.n_post = o.n_post; u.n_v2 = o.n_v2; u.lookups = o.lookups; u.events = o.events; u.cycles = o.cycles;
    }
    // ---- contigs whose seams did not hold (or whose segment slack ran out) are walked again as one segment; a contig whose
    // de-Bruijn search outgrew a standard trie (kTrieFull) is walked again with the grown pool
    std::vector<uint8_t> redo(n, 0);
    std::vector<uint32_t> todo;
    bool need_big = false;
    size_t n_slots = n_first;
    for (size_t q = 0; q < n_first; ++q) {
      const UnitIO& u = hs[q];
      if (!u.status) continue;
      const bool split = plan[q].n_segs > 1;
      if (u.status == kTrieFull || u.status == kFarBack || (split && (u.status == kSeamViolation || u.status == JB_E_SCRATCH))) {
        need_big = need_big || u.status == kTrieFull;
        if (!redo[u.unit]) { redo[u.unit] = 1; todo.push_back(u.unit); }
        continue;
      }
      b->status = u.status;
      return fail(JB_E_SCRATCH, "contig " + std::to_string(u.unit) + ": a device work pool overflowed during polishing "
                                "(output slack, fix-row arena, or copy-job arena; status " + std::to_string(u.status) + ")");
    }
    std::vector<UnitIO> segs;                            // accepted segments in (contig, segment) order
    const bool rewalked = !todo.empty();
    if (do_fix && speculate && !rewalked) ++n_speculative;
    if (!todo.empty()) {
      std::sort(todo.begin(), todo.end());
      std::vector<UnitIO> redone(n);
      bool use_big = need_big;
      for (uint32_t round = 1; !todo.empty(); ++round) {
        BridgePool rp = pool;
        if (use_big) {
          if (int rc = ensure_big_pool(w, s)) return rc;
          rp = BridgePool{w.pool_big.p, kBigPoolSlots, w.busy_big.p};
        }
        const size_t na = todo.size();
        UnitIO* again = hs + n_slots;
        std::vector<uint32_t> ord2(na);
        for (size_t q = 0; q < na; ++q) {
          again[q] = make_seg(b, whole_contig(todo[q], caps[todo[q]]), in, round, (uint32_t)(n_slots + q), keep_margin, cidx);
          ord2[q] = (uint32_t)(n_slots + q);
        }
        JB_CUDA(cudaMemcpyAsync(w.segs.p + n_slots, again, na * sizeof(UnitIO), cudaMemcpyHostToDevice, s));
        JB_CUDA(cudaMemcpyAsync(w.order.p, ord2.data(), na * 4, cudaMemcpyHostToDevice, s));
        JB_CUDA(cudaMemsetAsync(w.next.p, 0, 4, s));
        k_walk<<<walk_grid(na), kWalkWarps * 32, 0, s>>>(view, w.segs.p, w.order.p, (uint32_t)na, c2, rp, sink, jobs, w.next.p);
        JB_CUDA(cudaGetLastError());
        ++launches;
        JB_CUDA(cudaMemcpyAsync(again, w.segs.p + n_slots, na * sizeof(UnitIO), cudaMemcpyDeviceToHost, s));
        JB_CUDA(cudaMemcpyAsync(used, w.used.p, 32, cudaMemcpyDeviceToHost, s));
        JB_CUDA(cudaStreamSynchronize(s));
        n_reruns += (int64_t)na;
        std::vector<uint32_t> next_todo;
        for (size_t q = 0; q < na; ++q) {
          const UnitIO& u = again[q];
          if (!u.status) { redone[u.unit] = u; continue; }
          if (u.status == kTrieFull && !use_big) { next_todo.push_back(u.unit); continue; }
          b->status = u.status;
          if (u.status == kTrieFull)
            return fail(JB_E_SCRATCH, "contig " + std::to_string(u.unit) + ": the de-Bruijn path search (src/jasper.py:527-583) needs more than " +
                                      std::to_string(w.big_nodes) + " path extensions for one stretch; the grown pool is exhausted");
          return fail(JB_E_SCRATCH, "contig " + std::to_string(u.unit) + ": a device work pool overflowed during polishing (status " + std::to_string(u.status) + ")");
        }
        n_slots += na;
        todo.swap(next_todo);
        use_big = true;
      }
      for (size_t q = 0; q < n_first;) {
        const uint32_t unit = hs[q].unit;
        if (redo[unit]) {
          segs.push_back(redone[unit]);
          while (q < n_first && hs[q].unit == unit) ++q;
        } else segs.push_back(hs[q++]);
      }
    } else {
      segs.assign(hs, hs + n_first);
    }
    for (const UnitIO& u : segs) {
      b->passes[ite].wrong += u.wrong; walk_lookups += (int64_t)u.lookups; n_events += (int64_t)u.events; n_fast += u.n_fast; n_post += u.n_post; n_v2 += u.n_v2;
      max_seg_events = std::max<int64_t>(max_seg_events, (int64_t)u.events);
      max_seg_lookups = std::max<int64_t>(max_seg_lookups, (int64_t)u.lookups);
      max_seg_span = std::max<int64_t>(max_seg_span, u.end_old - u.base);
      if ((int64_t)u.cycles > slow_cycles[ite < 3 ? ite : 2]) {
        slow_cycles[ite < 3 ? ite : 2] = (int64_t)u.cycles; slow_events[ite < 3 ? ite : 2] = (int64_t)u.events;
        slow_lookups[ite < 3 ? ite : 2] = (int64_t)u.lookups; slow_span[ite < 3 ? ite : 2] = u.end_old - u.base;
      }
      sum_cycles[ite < 3 ? ite : 2] += (int64_t)u.cycles;
    }
    probe_steps += (int64_t)used[2];
    if (b->use_verdicts) n_verdicts += (int64_t)used[6];
    mark("  results merged", ite);
    if (do_fix) {
      // ---- the segments' new bases are stitched into the other buffer (launched above, or here after a re-walk); the fix
      // log follows on the side stream
      if (used[0] > w.rows.cap || used[1] > w.aux.cap || used[3] > w.jobs.cap) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "fix-row / copy-job arena overflow"); }
      std::vector<int64_t> new_lens(b->len);
      pending.place = place_segments(segs, &new_lens);
      pending.first_seg.assign(n, 0);
      for (size_t q = 0; q < segs.size(); ++q)
        if (segs[q].seg == 0) pending.first_seg[segs[q].unit] = q;
      for (size_t i = 0; i < n; ++i)
        if (new_lens[i] > caps[i]) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "contig " + std::to_string(i) + " outgrew its output slack"); }
      if (rewalked || !speculate) {
        // the host's own placement goes to the device (by launch slot; superseded attempts stay at -1)
        int64_t* dst_off = w.h_dst_off.p;
        uint32_t* slot_unit = w.h_order.p;                // (the queue order has been consumed: its staging area is free)
        ScanUnit* units_new = sunits;                     // (likewise this pass's staged geometry)
        for (size_t q = 0; q < n_slots; ++q) { dst_off[q] = -1; slot_unit[q] = 0; }
        for (size_t q = 0; q < segs.size(); ++q) {
          dst_off[segs[q].slot] = (int64_t)b->off[segs[q].unit] + pending.place[q].new_off;
          slot_unit[segs[q].slot] = segs[q].unit;
        }
        for (size_t i = 0; i < n; ++i) units_new[i] = ScanUnit{b->off[i], (uint64_t)new_lens[i], b->off[i]};
        JB_CUDA(cudaMemcpyAsync(w.dst_off.p, dst_off, n_slots * 8, cudaMemcpyHostToDevice, s));
        JB_CUDA(cudaMemcpyAsync(w.slot_unit.p, slot_unit, n_slots * 4, cudaMemcpyHostToDevice, s));
        JB_CUDA(cudaMemcpyAsync(w.scan_units_new.p, units_new, n * sizeof(ScanUnit), cudaMemcpyHostToDevice, s));
        JB_CUDA(cudaEventRecord(w.ev_walk, s));           // (the re-walks wrote rows too: the log travels again)
        if (int rc = send_fix_log()) return rc;
        if (int rc = stitch_and_recount()) return rc;
      }
      if (ite == passes - 1 && !b->early && b->fetch_follows && getenv("JB_NO_EARLY_FETCH") == nullptr) {
        uint64_t total_new = 0;
        for (size_t i = 0; i < n; ++i) total_new += (uint64_t)new_lens[i];
        b->early = take_out_buffer(std::max<uint64_t>(total_new, 1), &b->early_cap);
        if (b->early) {
          if (!b->ev_early) JB_CUDA(cudaEventCreateWithFlags(&b->ev_early, cudaEventDisableTiming));
          JB_CUDA(cudaStreamWaitEvent(w.s2, w.ev_concat, 0));      // (recorded behind the k_concat that counts)
          uint64_t at = 0;
          for (size_t i = 0; i < n; ++i) {
            if (new_lens[i]) JB_CUDA(cudaMemcpyAsync(b->early + at, outb + b->off[i], (size_t)new_lens[i], cudaMemcpyDeviceToHost, w.s2));
            at += (uint64_t)new_lens[i];
          }
          JB_CUDA(cudaEventRecord(b->ev_early, w.s2));
        }
      }
      have_counts = false;
      if (keep_counts) {
        if (verify_recount) {
          // full scan of the new sequence into the buffer the old counts lived in, then compare
          ScanArgs va = a;
          va.bases = outb; va.units = w.scan_units_new.p; va.counts = w.counts[cidx].p; va.okb = w.okb[cidx].p; va.probe_steps = nullptr;
          std::vector<uint32_t> tf(n + 1, 0);
          for (size_t i = 0; i < n; ++i) tf[i + 1] = tf[i] + tiles_of((uint64_t)new_lens[i]);
          JB_CUDA(cudaMemcpyAsync(w.tile_first.p, tf.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
          va.n_tiles = tf[n];
          if (int rc = launch_scan(view, va, s)) return rc;
          unsigned long long bad = 0;
          if (int rc = compare_counts(w.scan_units_new.p, (uint32_t)n, w.counts[cidx].p, w.okb[cidx].p, w.counts[cidx ^ 1].p, w.okb[cidx ^ 1].p,
                                      w.used.p + 5, &bad, s)) return rc;
          if (bad) { b->status = JB_E_SCRATCH; return fail(JB_E_SCRATCH, "k_recount disagrees with a full scan at " + std::to_string(bad) + " positions (pass " + std::to_string(ite) + ")"); }
        }
        cidx ^= 1;
        have_counts = true;
      }
      JB_CUDA(cudaStreamWaitEvent(s, w.ev_rows, 0));     // the next walk reuses the row arena
      pending.live = true; pending.buf = hb; pending.pass = (size_t)ite; pending.n_rows = used[0]; pending.n_aux = used[1];
      pending.segs.swap(segs);
      b->len = new_lens;
      b->cur ^= 1;
    }
    JB_CUDA(cudaEventRecord(pe.d, s));
    mark("concat launched", ite);
    // this pass's fix log lands while k_concat / k_recount run: hand it to its row thread now, not at the next pass
    if (int rc = finish_rows(b, &pending, &row_jobs)) return rc;
    // (no wait for the stream here: the next pass is queued behind this one's kernels straight away)
    mark("pass done", ite);
  }
  if (int rc = finish_rows(b, &pending, &row_jobs)) return rc;
  join_rows(b, &row_jobs);
  mark("rows joined", passes);
  JB_CUDA(cudaEventRecord(ev_end, s));
  JB_CUDA(cudaStreamSynchronize(s));
  mark("stream drained", passes);
  float ms = 0;
  for (size_t q = 0; q < b->passes.size() && q < pev.size(); ++q) {
    const PassEvents& pe = pev[q];
    cudaEventElapsedTime(&ms, pe.a, pe.b); scan_ms += ms;
    cudaEventElapsedTime(&ms, pe.b, pe.c); plan_ms += ms;
    cudaEventElapsedTime(&ms, pe.c, pe.d); walk_ms += ms;
    cudaEventElapsedTime(&ms, pe.c, pe.e); kwalk_ms += ms;
    if (pe.recount) { cudaEventElapsedTime(&ms, pe.r0, pe.r1); recount_ms += ms; recount_probes += (int64_t)w.h_probes.p[q]; }
    if (pe.verdicts) { cudaEventElapsedTime(&ms, pe.v0, pe.v1); verdict_ms += ms; }
  }
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  note_timing("polish", ms); note_timing("polish_scan", scan_ms); note_timing("polish_walk", walk_ms); note_timing("polish_plan", plan_ms);
  note_timing("polish_kwalk", kwalk_ms);
  note_timing("polish_recount", recount_ms);
  note_counter("scan_lookups", full_scan_lookups + recount_probes); note_counter("walk_lookups", walk_lookups);
  note_counter("full_scans", full_scans); note_counter("full_scan_lookups", full_scan_lookups); note_counter("recount_probes", recount_probes);
  note_counter("kernel_launches", launches); note_counter("events", n_events); note_counter("probe_steps", probe_steps);
  note_counter("max_seg_events", max_seg_events); note_counter("max_seg_lookups", max_seg_lookups); note_counter("max_seg_span", max_seg_span);
  for (int q = 0; q < 3; ++q) {
    const std::string sfx = std::to_string(q);
    note_counter(("slow_cycles" + sfx).c_str(), slow_cycles[q]); note_counter(("slow_events" + sfx).c_str(), slow_events[q]);
    note_counter(("slow_lookups" + sfx).c_str(), slow_lookups[q]); note_counter(("slow_span" + sfx).c_str(), slow_span[q]);
    note_counter(("sum_cycles" + sfx).c_str(), sum_cycles[q]);
  }
  note_counter("verdicts", n_verdicts); note_timing("polish_verdicts", verdict_ms); note_counter("verdict2_events", n_v2); note_counter("fast_events", n_fast); note_counter("fast_steps", n_post); note_counter("segments", n_segments); note_counter("segment_reruns", n_reruns); note_counter("seg_block", seg_block);
  note_counter("speculative_passes", n_speculative);
  w.release_big_pool();
  b->polished = true;
  while ((int)b->passes.size() <= passes) b->passes.push_back(PassOut{});
  return JB_OK;
}

extern "C" int jb_batch_fetch(jb_batch* b, jb_result** out) {
  if (!b || !out) return fail(JB_E_ARG, "jb_batch_fetch: null argument");
  JB_CUDA(cudaSetDevice(b->db->device));
  cudaStream_t s = b->db->stream;
  std::unique_ptr<jb_result> r(new jb_result());
  r->status = b->status;
  if (b->fetch_follows) r->passes = std::move(b->passes);      // jb_polish frees the batch right behind this call: no copy of the rows
  else r->passes = b->passes;
  r->seq_off.resize(b->n); r->seq_len.resize(b->n);
  uint64_t total = 0;
  for (size_t i = 0; i < b->n; ++i) { r->seq_off[i] = total; r->seq_len[i] = (uint64_t)b->len[i]; total += (uint64_t)b->len[i]; }
  if (b->early && b->ev_early) {                        // the copy started behind the last fixing pass
    JB_CUDA(cudaEventSynchronize(b->ev_early));
    r->data = b->early; r->data_cap = b->early_cap;
    b->early = nullptr; b->early_cap = 0;
    note_timing("d2h", 0.0); note_counter("d2h_bytes", (int64_t)total);
    *out = r.release();
    return JB_OK;
  }
  r->data = take_out_buffer(std::max<uint64_t>(total, 1), &r->data_cap);
  if (!r->data) return fail(JB_E_NOMEM, "cannot allocate the pinned result buffer");
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0)); JB_CUDA(cudaEventCreate(&e1));
  JB_CUDA(cudaEventRecord(e0, s));
  for (size_t i = 0; i < b->n; ++i)
    if (b->len[i]) JB_CUDA(cudaMemcpyAsync(r->data + r->seq_off[i], b->ws->buf[b->cur].p + b->off[i], (size_t)b->len[i], cudaMemcpyDeviceToHost, s));
  JB_CUDA(cudaEventRecord(e1, s));
  JB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  note_timing("d2h", ms); note_counter("d2h_bytes", (int64_t)total);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *out = r.release();
  return JB_OK;
}

extern "C" int jb_polish(const jb_db* db, const jb_seq_t* seqs, size_t n_seqs, int k, uint32_t solid_thre, int passes,
                         int fix, jb_result** out) {
  if (!out) return fail(JB_E_ARG, "jb_polish: null out");
  *out = nullptr;
  jb_batch* b = nullptr;
  auto t0 = std::chrono::steady_clock::now();
  int rc = upload_impl(db, seqs, n_seqs, &b, true);
  if (rc) return rc;
  auto t1 = std::chrono::steady_clock::now();
  b->fetch_follows = true;
  rc = jb_batch_polish(b, k, solid_thre, passes, fix);
  auto t2 = std::chrono::steady_clock::now();
  if (!rc) rc = jb_batch_fetch(b, out);
  jb_batch_free(b);
  auto t3 = std::chrono::steady_clock::now();
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
  note_timing("wall_upload", ms(t0, t1)); note_timing("wall_polish", ms(t1, t2)); note_timing("wall_fetch", ms(t2, t3));
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
  const PassOut& po = r->passes[ite];
  const size_t m = std::min(cap, po.rows.size());
  for (size_t i = 0; i < m; ++i) po.expand(po.rows[i], &rows[i]);
  return JB_OK;
}
extern "C" int64_t jb_result_num_seqs(const jb_result* r) { return r ? (int64_t)r->seq_len.size() : JB_E_ARG; }
extern "C" int jb_result_seq(const jb_result* r, size_t i, const char** data, uint64_t* len) {
  if (!r || i >= r->seq_len.size() || !data || !len) return fail(JB_E_ARG, "jb_result_seq: bad argument");
  *data = (const char*)r->data + r->seq_off[i]; *len = r->seq_len[i];
  return JB_OK;
}
extern "C" void jb_result_free(jb_result* r) { delete r; }

extern "C" int jb_db_replicate(jb_db* src, int n, const int* devices, jb_db** outs) {
  if (!src || (n && (!devices || !outs))) return fail(JB_E_ARG, "jb_db_replicate: null argument");
  if (src->n_parts > 1) return fail(JB_E_UNSUPPORTED, "jb_db_replicate: the source is a partitioned table");
  for (int i = 0; i < n; ++i) outs[i] = nullptr;
  auto undo = [&]() { for (int i = 0; i < n; ++i) { jb_db_close(outs[i]); outs[i] = nullptr; } };
  const size_t lut_bytes = (size_t)src->hdr.key_bytes() * 256 * 8;
  const size_t stash_bytes = src->stash_units ? (size_t)((src->stash_units + std::max<uint64_t>(4096, src->stash_units / 16) + 4) * 2 * sizeof(Slot)) : 0;
  for (int i = 0; i < n; ++i) {
    auto d = new jb_db(*src);
    d->device = devices[i]; d->mem = nullptr; d->lut = nullptr; d->stream = nullptr; d->stash = nullptr; d->stash_list = nullptr;
    d->idx_slots = nullptr; d->idx_off = nullptr; d->idx_bits = 0;
    for (int q = 0; q < kMaxParts; ++q) { d->parts[q] = nullptr; d->ipc_mapped[q] = false; }
    outs[i] = d;
    cudaError_t e = cudaSetDevice(devices[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&d->mem, src->n_units * (uint64_t)src->unit_bytes());
    if (e == cudaSuccess) e = cudaMalloc(&d->lut, lut_bytes);
    if (e == cudaSuccess && stash_bytes) e = cudaMalloc(&d->stash, stash_bytes);
    if (e == cudaSuccess && src->n_stash) e = cudaMalloc(&d->stash_list, src->n_stash * sizeof(Slot));
    // device-to-device over NVLink; one disk read serves every GPU of the box
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(d->mem, devices[i], src->mem, src->device, src->n_units * (uint64_t)src->unit_bytes(), d->stream);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(d->lut, devices[i], src->lut, src->device, lut_bytes, d->stream);
    if (e == cudaSuccess && stash_bytes) e = cudaMemcpyPeerAsync(d->stash, devices[i], src->stash, src->device, stash_bytes, d->stream);
    if (e == cudaSuccess && src->n_stash) e = cudaMemcpyPeerAsync(d->stash_list, devices[i], src->stash_list, src->device, src->n_stash * sizeof(Slot), d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { undo(); cudaGetLastError(); return fail(JB_E_CUDA, std::string("jb_db_replicate: ") + cudaGetErrorString(e)); }
    d->parts[0] = d->mem;
    if (src->idx_bits) {                                  // the scan index travels too when it fits (else this replica scans through the hash table)
      const size_t off_bytes = ((size_t)1 << src->idx_bits) * (src->idx_wide_off ? 8 : 4) + (src->idx_wide_off ? 8 : 4);
      const size_t slot_bytes = (size_t)std::max<uint64_t>(src->idx_n_slots, 1) * 16;
      cudaError_t e2 = cudaMalloc(&d->idx_slots, slot_bytes);
      if (e2 == cudaSuccess) e2 = cudaMalloc(&d->idx_off, off_bytes);
      if (e2 == cudaSuccess) e2 = cudaMemcpyPeerAsync(d->idx_slots, devices[i], src->idx_slots, src->device, slot_bytes, d->stream);
      if (e2 == cudaSuccess) e2 = cudaMemcpyPeerAsync(d->idx_off, devices[i], src->idx_off, src->device, off_bytes, d->stream);
      if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(d->stream);
      if (e2 == cudaSuccess) d->idx_bits = src->idx_bits;
      else { cudaGetLastError(); free_scan_index(d); }
    }
  }
  return JB_OK;
}

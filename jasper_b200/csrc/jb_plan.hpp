// jb_plan.hpp -- host-side planning and assembly around the segment walkers (shared by the
// product's orchestration in jb_polish.cu and by the test-only host build of the walker).
//
// One pass over a batch:
//   scan (event / weak bits)  ->  first anchor of every seg_block positions  ->  plan_segments
//   ->  one walker per segment  ->  contigs with a seam violation are re-walked as one segment
//   ->  place_segments (prefix sums of the segments' new lengths)  ->  concat + fix rows.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <functional>
#include <new>
#include <stdexcept>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

#include "jb_rows.hpp"

namespace jb {

struct SegPlan {
  uint32_t unit, seg, n_segs;
  int64_t base, next_anchor;      // next_anchor = -1 for the last segment of the contig
  int64_t stage_off, stage_cap;   // where the segment materialises edited bases, relative to the contig's staging region
};

struct BlockRef { uint32_t unit; uint32_t pad; uint64_t start; };

// the blocks k_find_anchors inspects: every seg_block positions of every contig long enough to split
inline std::vector<BlockRef> make_blocks(const std::vector<int64_t>& lens, int64_t seg_block) {
  std::vector<BlockRef> b;
  for (size_t u = 0; u < lens.size(); ++u)
    for (int64_t s = seg_block; s < lens[u]; s += seg_block) b.push_back(BlockRef{(uint32_t)u, 0, (uint64_t)s});
  return b;
}

// anchors[i] belongs to blocks[i] (-1 = none).  caps[u] = bytes of contig u's staging region.
inline std::vector<SegPlan> plan_segments(const std::vector<int64_t>& lens, const std::vector<int64_t>& caps,
                                          const std::vector<BlockRef>& blocks, const long long* anchors) {
  constexpr int64_t kPadMax = 1024, kPadMin = 96;   // growth room of one segment before it needs the whole-contig fallback
  std::vector<SegPlan> plan;
  size_t bi = 0;
  for (size_t u = 0; u < lens.size(); ++u) {
    std::vector<int64_t> cuts{0};
    for (; bi < blocks.size() && blocks[bi].unit == u; ++bi)
      if (anchors && anchors[bi] > cuts.back()) cuts.push_back((int64_t)anchors[bi]);
    int64_t pad = std::min<int64_t>(kPadMax, (caps[u] - lens[u]) / (int64_t)cuts.size());
    while (pad < kPadMin && cuts.size() > 1) {       // not enough slack for that many segments: keep every other cut
      size_t w = 1;
      for (size_t q = 2; q < cuts.size(); q += 2) cuts[w++] = cuts[q];
      cuts.resize(w);
      pad = std::min<int64_t>(kPadMax, (caps[u] - lens[u]) / (int64_t)cuts.size());
    }
    const uint32_t n = (uint32_t)cuts.size();
    for (uint32_t s = 0; s < n; ++s) {
      SegPlan p;
      p.unit = (uint32_t)u; p.seg = s; p.n_segs = n;
      p.base = cuts[s]; p.next_anchor = s + 1 < n ? cuts[s + 1] : -1;
      const int64_t end = s + 1 < n ? cuts[s + 1] : lens[u];
      p.stage_off = cuts[s] + (int64_t)s * pad;
      p.stage_cap = n == 1 ? caps[u] : (end - cuts[s]) + pad;
      plan.push_back(p);
    }
  }
  return plan;
}

inline SegPlan whole_contig(uint32_t unit, int64_t cap) { return SegPlan{unit, 0, 1, 0, -1, 0, cap}; }

// After the walkers: per segment the offset of its new bases inside the new contig and the
// coordinate shift its rows need (segment-local new coordinate + shift = coordinate in the new contig).
struct SegPlace { int64_t new_off; int64_t row_shift; };

// segs: accepted segments in (unit, seg) order.
inline std::vector<SegPlace> place_segments(const std::vector<UnitIO>& segs, std::vector<int64_t>* new_lens) {
  std::vector<SegPlace> out(segs.size());
  int64_t run = 0;
  for (size_t i = 0; i < segs.size(); ++i) {
    if (segs[i].seg == 0) run = 0;
    out[i].new_off = run;
    out[i].row_shift = run - segs[i].base;          // bases before `base` moved by the earlier segments' net change
    run += segs[i].new_len;
    if (i + 1 == segs.size() || segs[i + 1].seg == 0) (*new_lens)[segs[i].unit] = run;
  }
  return out;
}

// Keep the rows of the accepted launch of every contig, move them into contig coordinates and put them in
// (contig, segment, emission) order WITHOUT sorting: a segment's rows carry distinct increasing sequence
// numbers below its n_rows, so each row has a fixed slot behind the rows of the segments before it.
// first_seg[u] = index of contig u's first accepted segment in segs / place.
inline void settle_rows(std::vector<DevRow>* rows, const std::vector<UnitIO>& segs, const std::vector<SegPlace>& place,
                        const std::vector<size_t>& first_seg) {
  std::vector<uint64_t> slot0(segs.size() + 1, 0);
  for (size_t q = 0; q < segs.size(); ++q) slot0[q + 1] = slot0[q] + segs[q].n_rows;
  std::vector<DevRow> placed((size_t)slot0[segs.size()]);
  std::vector<uint8_t> filled(placed.size(), 0);
  for (const DevRow& r0 : *rows) {
    const size_t s0 = first_seg[r0.unit];
    if (segs[s0].gen != r0.gen) continue;             // superseded attempt
    const size_t q = s0 + r0.seg;
    if (r0.seqno >= segs[q].n_rows) continue;         // cannot happen; stay in bounds regardless
    const size_t at = (size_t)slot0[q] + r0.seqno;
    placed[at] = r0;
    placed[at].coord += place[q].row_shift;
    filled[at] = 1;
  }
  size_t w = 0;
  for (size_t i = 0; i < placed.size(); ++i) if (filled[i]) placed[w++] = placed[i];
  placed.resize(w);
  rows->swap(placed);
}

// The same without copying or sorting the log: walk the rows in (contig, segment, emission) order through
// an index and emit csv rows straight from the (pinned) log.
inline bool rows_in_order(const DevRow* rows, size_t n_rows, const uint8_t* aux, const std::vector<UnitIO>& segs,
                          const std::vector<SegPlace>& place, const std::vector<size_t>& first_seg, PassOut* out) {
  std::vector<uint64_t> slot0(segs.size() + 1, 0);
  for (size_t q = 0; q < segs.size(); ++q) slot0[q + 1] = slot0[q] + segs[q].n_rows;
  std::vector<uint32_t> at((size_t)slot0[segs.size()], 0xFFFFFFFFu);
  const size_t n_slots = at.size();
  // Three phases on a few threads spawned once (the 63 k rows of a 64 Mb pass cost ~8 ms on one, and this runs beside the
  // GPU work of the following passes): index the log by (segment, emission order), convert slot ranges, splice in order.
  // (the alignments of bridged patches dominate: ~10 ms of one core per 64 Mb pass; with four threads this finished after the GPU)
  int n_thr = n_slots > 8192 ? (int)std::min<unsigned>(8, std::max<unsigned>(2, std::thread::hardware_concurrency() / 2)) : 1;
  if (const char* e = getenv("JB_ROW_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 16) n_thr = v; }   // tests
  // the few per-segment facts the index phase needs, packed (the descriptors are 250 bytes apiece)
  std::vector<uint32_t> seg_gen(segs.size()), seg_rows(segs.size());
  for (size_t q = 0; q < segs.size(); ++q) { seg_gen[q] = segs[q].gen; seg_rows[q] = segs[q].n_rows; }
  // phases 2 and 3 work on chunks of the slot range handed out dynamically: the alignments of bridged patches are what costs, and
  // they are not spread evenly (r2q: with one static range per thread the first thread was done at 0.64 ms, the last at 1.26)
  const int n_chunks = n_thr == 1 ? 1 : n_thr * 8;
  std::vector<PassOut> part((size_t)n_chunks);
  std::vector<char> part_ok((size_t)n_chunks, 1);
  std::vector<size_t> row_at((size_t)n_chunks + 1, out->rows.size());
  std::vector<uint32_t> spill_at((size_t)n_chunks + 1, (uint32_t)out->spill.size());
  std::atomic<int> arrived{0}, next_convert{0}, next_splice{0};
  auto barrier = [&](int phase) {                      // all workers reach phase `phase`; the threads are spawned once
    arrived.fetch_add(1, std::memory_order_acq_rel);
    while (arrived.load(std::memory_order_acquire) < phase * n_thr) std::this_thread::yield();
  };
  const bool trace = getenv("JB_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  const size_t rows_before = out->rows.size();
  auto work = [&](int t) {
    if (t == 0 && n_thr > 1) {
      // worker 0 makes room for the result while the others index: a fresh 2 MB vector costs ~0.4 ms of page faults, which used to
      // sit between the conversion and the splice with every other worker waiting (r2r); the size is an estimate, corrected below
      if (rows_before == 0) out->rows = RowStore::take(n_slots + n_slots / 8 + 1024);      // (recycled pages when there are any)
      out->rows.resize(rows_before + n_slots + n_slots / 8 + 1024);
    } else {                                            // 1. index: every live row has its own slot, the scatter needs no locks
      const size_t nt = n_thr > 1 ? (size_t)n_thr - 1 : 1, me = n_thr > 1 ? (size_t)t - 1 : 0;
      const size_t lo = n_rows * me / nt, hi = n_rows * (me + 1) / nt;
      for (size_t i = lo; i < hi; ++i) {
        const DevRow& r = rows[i];
        const size_t s0 = first_seg[r.unit];
        if (seg_gen[s0] != r.gen) continue;            // superseded attempt
        const size_t q = s0 + r.seg;
        if (r.seqno >= seg_rows[q]) continue;
        at[(size_t)slot0[q] + r.seqno] = (uint32_t)i;
      }
    }
    barrier(1);
    if (trace && t == 0) fprintf(stderr, "[jb trace]   rows: indexed at %.3f ms (%d threads)\n", since(), n_thr);
    for (int c = next_convert.fetch_add(1); c < n_chunks; c = next_convert.fetch_add(1)) {   // 2. convert chunks of slots
      const size_t lo = n_slots * (size_t)c / (size_t)n_chunks, hi = n_slots * (size_t)(c + 1) / (size_t)n_chunks;
      size_t q = (size_t)(std::upper_bound(slot0.begin(), slot0.end(), (uint64_t)lo) - slot0.begin());
      q = q ? q - 1 : 0;
      part[(size_t)c].rows = RowStore::take((hi - lo) + (hi - lo) / 4 + 16);
      part[(size_t)c].rows.reserve((hi - lo) + 16);
      for (size_t sl = lo; sl < hi; ++sl) {
        while (sl >= slot0[q + 1]) ++q;
        if (at[sl] == 0xFFFFFFFFu) continue;
        const DevRow& r = rows[at[sl]];
        // (nothing may throw past this point: the other workers wait at the barrier, and an exception that leaves a thread
        //  ends the process -- the alignment matrices of a very long bridged patch are the realistic std::bad_alloc)
        try { if (!row_to_csv(r, r.coord + place[q].row_shift, aux, &part[(size_t)c])) part_ok[(size_t)c] = 0; }
        catch (const std::exception&) { part_ok[(size_t)c] = 2; }
      }
    }
    if (trace) fprintf(stderr, "[jb trace]   rows: thread %d converted at %.3f ms\n", t, since());
    barrier(2);
    if (t == 0) {                                       // (sizes are known now: worker 0 lays the output out)
      for (int u = 0; u < n_chunks; ++u) {
        row_at[(size_t)u + 1] = row_at[(size_t)u] + part[(size_t)u].rows.size();
        spill_at[(size_t)u + 1] = spill_at[(size_t)u] + (uint32_t)part[(size_t)u].spill.size();
      }
      if (out->rows.size() < row_at[(size_t)n_chunks]) out->rows.resize(row_at[(size_t)n_chunks]);
      for (int u = 0; u < n_chunks; ++u) out->spill += part[(size_t)u].spill;
    }
    barrier(3);
    if (trace && t == 0) fprintf(stderr, "[jb trace]   rows: laid out at %.3f ms\n", since());
    for (int c = next_splice.fetch_add(1); c < n_chunks; c = next_splice.fetch_add(1)) {     // 3. splice in order
      HostRow* dst = out->rows.data() + row_at[(size_t)c];
      const std::vector<HostRow>& src = part[(size_t)c].rows;
      for (size_t i = 0; i < src.size(); ++i) {
        HostRow h = src[i];
        if (h.mut_len + h.orig_len > (int)sizeof h.text) h.spill += spill_at[(size_t)c];
        dst[i] = h;
      }
      if (n_thr > 1) RowStore::give(std::move(part[(size_t)c].rows));
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  out->rows.resize(row_at[(size_t)n_chunks]);           // (the estimate shrinks to what was spliced)
  bool ok = true;
  for (int c = 0; c < n_chunks; ++c) {
    if (part_ok[(size_t)c] == 2) throw std::bad_alloc();          // on the calling thread, where the caller catches it
    ok = ok && part_ok[(size_t)c];
  }
  if (trace) fprintf(stderr, "[jb trace]   rows: done at %.3f ms\n", since());
  return ok;
}

}  // namespace jb

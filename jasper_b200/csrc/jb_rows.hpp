// jb_rows.hpp -- host-side conversion of the walker's event log into fix.csv rows
// (src/jasper.py:214-222 and, for multi-error patches, :305-329).
#pragma once
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "jb_align.hpp"
#include "jb_walker.cuh"

namespace jb {

// One csv row, compact: almost every row carries 1-3 characters of text, and the rows of a 64 Mb pass are
// built beside the GPU work, so they must not cost 144 bytes (and a page fault per 28 rows) each.
struct HostRow {
  int64_t coord;
  uint32_t seq_index;
  uint32_t spill;            // offset of the text in PassOut::spill when it does not fit inline
  uint8_t mut_len, orig_len;
  char text[14];             // mutation then original
};
static_assert(sizeof(HostRow) == 32, "HostRow is meant to be 32 bytes");

// Row storage is recycled inside the process: a vector of 63 k rows is 2 MB of fresh pages every time it is allocated, and
// first-touch page faults cost ~2.5 us apiece on the boxes measured (r2s: 1.4 ms for one result vector) -- more than all the
// alignments of a pass on eight threads.  Finished results hand their vectors back (jb_result_free / jb_batch_free), the
// row threads take them from here.
struct RowStore {
  static std::mutex& lock() { static std::mutex m; return m; }
  static std::vector<std::vector<HostRow>>& pool() { static std::vector<std::vector<HostRow>> p; return p; }
  // an empty vector, with at least `want` rows of capacity when the pool has one
  static std::vector<HostRow> take(size_t want) {
    std::vector<HostRow> v;
    {
      std::lock_guard<std::mutex> g(lock());
      auto& p = pool();
      size_t best = p.size();
      for (size_t i = 0; i < p.size(); ++i)
        if (p[i].capacity() >= want && (best == p.size() || p[i].capacity() < p[best].capacity())) best = i;
      if (best < p.size()) { v.swap(p[best]); p.erase(p.begin() + (long)best); }
    }
    v.clear();
    return v;
  }
  static void give(std::vector<HostRow>&& v) {
    if (v.capacity() < 256) return;
    v.clear();
    std::lock_guard<std::mutex> g(lock());
    auto& p = pool();
    size_t bytes = v.capacity() * sizeof(HostRow);
    for (const auto& q : p) bytes += q.capacity() * sizeof(HostRow);
    if (p.size() < 256 && bytes <= (256u << 20)) p.push_back(std::move(v));     // (bounded: at most 256 MB parked)
  }
};

struct PassOut {
  int64_t wrong = 0, total = 0;
  std::vector<HostRow> rows;
  std::string spill;
  const char* mutation(const HostRow& r) const { return (r.mut_len + r.orig_len <= (int)sizeof r.text) ? r.text : spill.data() + r.spill; }
  const char* original(const HostRow& r) const { return mutation(r) + r.mut_len; }
  void add(uint32_t seq_index, int64_t coord, const char* mut, size_t mut_len, const char* orig, size_t orig_len) {
    HostRow h;
    if (mut_len > 63) mut_len = 63;
    if (orig_len > 63) orig_len = 63;
    h.coord = coord; h.seq_index = seq_index; h.spill = 0; h.mut_len = (uint8_t)mut_len; h.orig_len = (uint8_t)orig_len;
    if (mut_len + orig_len <= sizeof h.text) { std::memcpy(h.text, mut, mut_len); std::memcpy(h.text + mut_len, orig, orig_len); }
    else { h.spill = (uint32_t)spill.size(); spill.append(mut, mut_len); spill.append(orig, orig_len); }
    rows.push_back(h);
  }
  void expand(const HostRow& r, jb_fix_row_t* o) const {
    std::memset(o, 0, sizeof *o);                  // consumers read the char arrays as fixed-width fields
    o->seq_index = r.seq_index; o->coord = r.coord;
    std::memcpy(o->mutation, mutation(r), r.mut_len);
    std::memcpy(o->original, original(r), r.orig_len);
  }
};

// Appends the csv rows of `rows` (already in (contig, segment, emission) order) to out.  Returns false when
// the reference would have crashed (IndexError at src/jasper.py:221 / :309 -> exit 1).
// One logged row -> zero, one or two csv rows.
inline bool row_to_csv(const DevRow& r, int64_t coord, const uint8_t* aux, PassOut* out) {
  bool ok = true;
  {
    if (r.kind == kRowPlain || r.kind == kRowSpill) {
      const char* text = r.kind == kRowPlain ? r.text : (const char*)aux + r.aux_off;
      out->add(r.unit, coord, text, r.fb_len, text + r.fb_len, r.orig_len);
      return true;
    }
    uint32_t patch_len = 0, orig_len = 0;
    for (int q = 0; q < 4; ++q) { patch_len |= (uint32_t)aux[r.aux_off + q] << (8 * q); orig_len |= (uint32_t)aux[r.aux_off + 4 + q] << (8 * q); }
    std::string patch((const char*)aux + r.aux_off + 8, patch_len);
    std::string orig((const char*)aux + r.aux_off + 8 + patch_len, orig_len);
    std::string pa, oa;
    struct Diff { int64_t coord; std::string mutation, original; };
    std::vector<Diff> diffs;
    if (first_global_alignment(patch, orig, &pa, &oa)) {
      for (size_t c = 0; c < pa.size() && diffs.size() < 2; ++c) {     // only the first two are ever written (src/jasper.py:218-222)
        if (pa[c] == oa[c]) continue;
        Diff h{coord + (int64_t)c, "", ""};
        if (pa[c] == '-') { h.mutation = "-"; h.original = std::string("i") + oa[c]; }
        else if (oa[c] == '-') { h.mutation = std::string(1, pa[c]); h.original = "d-"; }
        else { h.mutation = std::string(1, pa[c]); h.original = std::string("s") + oa[c]; }
        diffs.push_back(h);
      }
    }
    if (diffs.empty()) {
      ok = false;
    } else if (diffs.size() == 1) {                // csv.writer prints the one-element lists' repr
      diffs[0].mutation = "['" + diffs[0].mutation + "']";
      diffs[0].original = "['" + diffs[0].original + "']";
    }
    for (const Diff& h : diffs) out->add(r.unit, h.coord, h.mutation.data(), h.mutation.size(), h.original.data(), h.original.size());
  }
  return ok;
}

inline bool rows_to_csv(const std::vector<DevRow>& rows, const std::vector<uint8_t>& aux, PassOut* out) {
  bool ok = true;
  out->rows.reserve(out->rows.size() + rows.size() + 16);
  for (const DevRow& r : rows) ok = row_to_csv(r, r.coord, aux.data(), out) && ok;
  return ok;
}

// Sorts `rows` by (contig, segment, emission order) first (callers without a segment table).
inline bool rows_from_log(std::vector<DevRow>& rows, const std::vector<uint8_t>& aux, PassOut* out) {
  std::sort(rows.begin(), rows.end(), [](const DevRow& x, const DevRow& y) {
    if (x.unit != y.unit) return x.unit < y.unit;
    if (x.seg != y.seg) return x.seg < y.seg;
    return x.seqno < y.seqno;
  });
  return rows_to_csv(rows, aux, out);
}

inline int sampling_step(int k) {                  // max(2, round(k/8)), Python's half-to-even round (src/jasper.py:20)
  double r = __builtin_nearbyint((double)k / 8.0);
  return std::max(2, (int)r);
}

}  // namespace jb

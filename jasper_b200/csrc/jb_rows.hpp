// jb_rows.hpp -- host-side conversion of the walker's event log into fix.csv rows
// (src/jasper.py:214-222 and, for multi-error patches, :305-329).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "jb_align.hpp"
#include "jb_walker.cuh"

namespace jb {

struct HostRow { uint32_t seq_index; int64_t coord; std::string mutation, original; };
struct PassOut { int64_t wrong = 0, total = 0; std::vector<HostRow> rows; };

// Appends the csv rows of `rows` (already in (contig, segment, emission) order) to out.  Returns false when
// the reference would have crashed (IndexError at src/jasper.py:221 / :309 -> exit 1).
inline bool rows_to_csv(const std::vector<DevRow>& rows, const std::vector<uint8_t>& aux, std::vector<HostRow>* out) {
  bool ok = true;
  out->reserve(out->size() + rows.size() + 16);
  for (const DevRow& r : rows) {
    if (r.kind == kRowPlain) {
      out->push_back(HostRow{r.unit, r.coord, std::string(r.fb, r.fb_len), std::string(r.orig, r.orig_len)});
      continue;
    }
    std::string patch((const char*)aux.data() + r.aux_off, r.aux_patch_len);
    std::string orig((const char*)aux.data() + r.aux_off + r.aux_patch_len, r.aux_orig_len);
    std::string pa, oa;
    std::vector<HostRow> diffs;
    if (first_global_alignment(patch, orig, &pa, &oa)) {
      for (size_t c = 0; c < pa.size(); ++c) {
        if (pa[c] == oa[c]) continue;
        HostRow h{r.unit, r.coord + (int64_t)c, "", ""};
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
      out->push_back(diffs[0]);
    } else {
      out->push_back(diffs[0]);
      out->push_back(diffs[1]);
    }
  }
  return ok;
}

// Sorts `rows` by (contig, segment, emission order) first (callers without a segment table).
inline bool rows_from_log(std::vector<DevRow>& rows, const std::vector<uint8_t>& aux, std::vector<HostRow>* out) {
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

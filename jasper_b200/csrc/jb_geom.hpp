// jb_geom.hpp -- host-side choice of the table geometry (jb_kmer.cuh "table geometry") and the GF(2) algebra behind the
// compact layout.  Shared by the loader (jb_table.cu) and by the CPU build of the table logic in tests/hostsim.
//
// Jellyfish keeps, in memory, only the high bits of a key next to the position it hashed to, because its matrix is
// drawn until the r x r block over the low r key bits is invertible (JF!lib/rectangular_binary_matrix.cc:160-204,
// :240-247; JF!include/jellyfish/large_hash_array.hpp:959).  The file keeps whole keys; we go back to the quotient.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "jb_header.hpp"
#include "jb_kmer.cuh"

namespace jb {

inline int ceil_log2_u64(uint64_t x) { int b = 0; while (b < 64 && (1ull << b) < x) ++b; return b; }

// column of the hash matrix that key bit b (0 = LSB) selects
inline uint64_t matrix_column(const JfHeader& h, int b) {
  if (h.matrix_identity) return b < 64 ? 1ull << b : 0ull;
  return h.columns[(size_t)(h.key_bits - 1 - b)];
}

// Rows of the inverse of the low r x r block: key bit b (< r) = parity(inv[b] & v) where v = pos ^ M * (key with its low
// r bits cleared).  Empty vector = the block is singular (no compact layout for this file).
inline std::vector<uint64_t> low_block_inverse(const JfHeader& h, int r) {
  if (r > 64 || r > h.key_bits) return {};
  if (r == 0) return std::vector<uint64_t>(1, 0);       // (never used: a marker for "invertible")
  const uint64_t rmask = r == 64 ? ~0ull : (1ull << r) - 1;
  // row i of A: bit b set iff column(b) has bit i; augmented with the identity
  std::vector<uint64_t> a((size_t)r, 0), inv((size_t)r, 0);
  for (int b = 0; b < r; ++b) {
    const uint64_t col = matrix_column(h, b) & rmask;
    for (int i = 0; i < r; ++i) if ((col >> i) & 1) a[(size_t)i] |= 1ull << b;
  }
  for (int i = 0; i < r; ++i) inv[(size_t)i] = 1ull << i;
  for (int b = 0; b < r; ++b) {                          // pivot on unknown b
    int p = -1;
    for (int i = b; i < r; ++i) if ((a[(size_t)i] >> b) & 1) { p = i; break; }
    if (p < 0) return {};
    std::swap(a[(size_t)b], a[(size_t)p]); std::swap(inv[(size_t)b], inv[(size_t)p]);
    for (int i = 0; i < r; ++i)
      if (i != b && ((a[(size_t)i] >> b) & 1)) { a[(size_t)i] ^= a[(size_t)b]; inv[(size_t)i] ^= inv[(size_t)b]; }
  }
  return inv;                                            // a is the identity now: x_b = parity(inv[b] & v)
}

struct GeomChoice {
  TableGeom g{};
  uint64_t n_units = 0;        // home range + tail + guard
  uint64_t tail_units = 0;
  double load = 0;
  std::string note;            // why not compact, if so
};

// compact: 1.05 sectors per present key at 0.5 (16 bytes per record), 1.14 at 0.65, 1.38 at 0.8; the loader raises the load
// until the table fits the device
constexpr double kDefaultLoadCompact = 0.5, kDefaultLoadWide = 0.5;

// layout_pref: -1 = compact when possible, else wide; kLayoutWide / kLayoutCompact = forced (compact may still fall back)
inline GeomChoice choose_geometry(const JfHeader& h, uint64_t n_records, double load, int layout_pref) {
  GeomChoice c;
  TableGeom& g = c.g;
  const int r = ceil_log2_u64(h.size);
  g.pos_mask = h.size - 1; g.k = h.k(); g.key_bytes = h.key_bytes(); g.r_bits = r;
  const uint64_t n = std::max<uint64_t>(n_records, 1);
  if (layout_pref != kLayoutWide) {
    const double ld = load > 0 ? load : kDefaultLoadCompact;
    const uint64_t lines = std::max<uint64_t>(64, (uint64_t)((double)n / (kCompactSlots * ld)) + 1);     // sectors
    // 4-8 posx values per sector: fine enough for an even spread, coarse enough for a short tag (every tag bit is a count bit less)
    const int ext = std::max(0, std::min({ceil_log2_u64(lines) + 2 - r, h.key_bits, 64 - r}));
    const int bits = r + ext;
    // posx values per sector; a probe may meet records of kMaxHops + 1 consecutive home sectors
    const long double per_line = (long double)(bits >= 64 ? 18446744073709551616.0L : (long double)(1ull << bits)) / (long double)lines;
    const uint64_t span = (uint64_t)(per_line * (kMaxHops + 1)) + 4;
    const int tag = std::min(bits, ceil_log2_u64(span + 1));
    const int rem = h.key_bits - r;
    const int cnt = std::min(64 - rem - tag, kMaxCvalBits + 1);
    if (r > h.key_bits || r > 64) c.note = "hash size exceeds the key space";
    else if (low_block_inverse(h, r).empty()) c.note = "the hash matrix's block over the low key bits is singular";
    else if (cnt < 9) c.note = "key remainder (" + std::to_string(rem) + " bits) + tag (" + std::to_string(tag) + ") leave no room for a count in 64 bits";
    else {
      g.layout = kLayoutCompact; g.units = lines; g.ext_bits = ext; g.left_shift = 64 - bits; g.key_top_shift = h.key_bits - ext;
      g.tag_bits = tag; g.cnt_bits = cnt;
      c.tail_units = std::max<uint64_t>(256, lines / 4096);
      c.n_units = lines + c.tail_units + 32;
      c.load = ld;
      return c;
    }
  }
  const double ld = load > 0 ? load : kDefaultLoadWide;
  const uint64_t half = std::max<uint64_t>(256, (uint64_t)((double)n / ld / 2.0) + 1);
  const int want = ceil_log2_u64(half) + 6 - r;
  const int ext = std::max(0, std::min({want, h.key_bits, 64 - r}));
  g.layout = kLayoutWide; g.units = half; g.ext_bits = ext; g.left_shift = 64 - (r + ext); g.key_top_shift = h.key_bits - ext;
  g.tag_bits = 0; g.cnt_bits = 32;
  c.tail_units = std::max<uint64_t>(32768, half / 64);
  c.n_units = half + c.tail_units + 4;
  c.load = ld;
  return c;
}

inline int unit_bytes_of(const TableGeom&) { return 32; }
inline int slots_per_unit(const TableGeom& g) { return g.layout == kLayoutCompact ? kCompactSlots : 2; }

}  // namespace jb

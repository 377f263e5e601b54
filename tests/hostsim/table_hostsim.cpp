// table_hostsim.cpp -- TEST INFRASTRUCTURE: the compact table's arithmetic (jb_kmer.cuh / jb_geom.hpp) run on the CPU.
// Places the records of a .jf sequentially by the loader's rule slot_i = max(home_i, slot_{i-1} + 1), then checks what
// the device code relies on: every record is found from its home line within kMaxHops (or is in the stash), absent
// keys are not found, no tag is ambiguous, and every slot inverts back to its key (the matrix's low block).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../jasper_b200/csrc/jb_geom.hpp"

using namespace jb;

namespace {
struct Rec { Key key; uint32_t count; uint64_t posx; };

// the host twin of compact_find (jb_internal.cuh): 0 = absent, else count; *need_stash when the stash must answer
uint32_t host_find(const TableGeom& g, const std::vector<uint64_t>& words, const Key& key, uint64_t posx, bool* need_stash, int* lines) {
  uint64_t line = home_unit_of_posx(g, posx);
  const uint64_t pat = compact_pattern(g, posx, key), lim = 1ull << g.cnt_bits;
  const uint32_t cmask = cval_mask(g);
  *need_stash = false;
  for (int hop = 0; hop <= kMaxHops; ++hop, ++line) {
    ++*lines;
    bool empty = false;
    for (int q = 0; q < kCompactSlots; ++q) {
      const uint64_t w = words[line * kCompactSlots + q], x = w ^ pat;
      if (w == 0) empty = true;
      const uint32_t cv = (uint32_t)x & cmask;
      if (x < lim && cv != cmask - 1u && w != 0) { if (cv == cmask) *need_stash = true; return cv; }
    }
    if (empty || !(words[line * kCompactSlots + kCompactSlots - 1] & spill_bit(g))) return 0;
  }
  *need_stash = true;
  return 0;
}
}  // namespace

// returns 0 when everything holds; fills stats = {layout, lines, tag_bits, cnt_bits, n_stash, max displacement in lines,
// lines read per present key x 1000, lines read per absent key x 1000, records}
extern "C" int hostsim_table_check(const char* path, double load, int layout_pref, int max_cval_bits, long long* stats, char* msg, int cap) {
  auto say = [&](const std::string& m) { snprintf(msg, (size_t)cap, "%s", m.c_str()); return 1; };
  std::ifstream f(path, std::ios::binary);
  if (!f) return say("cannot open");
  std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  JfHeader h;
  std::string err = parse_jf_header(all.data(), all.size(), &h);
  if (!err.empty()) return say(err);
  const int rec = h.record_bytes();
  const uint64_t n = (all.size() - h.data_offset) / rec;
  GeomChoice gc = choose_geometry(h, n, load, layout_pref);
  TableGeom g = gc.g;
  stats[0] = g.layout;
  if (g.layout != kLayoutCompact) { snprintf(msg, (size_t)cap, "%s", gc.note.c_str()); return 0; }
  if (max_cval_bits > 0 && g.cnt_bits > max_cval_bits + 1) g.cnt_bits = max_cval_bits + 1;      // force escapes
  const std::vector<uint64_t> lut = build_hash_lut(h);
  const Key mask = key_mask(h.k());
  std::vector<Rec> recs(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint8_t* p = (const uint8_t*)all.data() + h.data_offset + i * rec;
    Key key = {0, 0};
    for (int b = 0; b < h.key_bytes(); ++b) { uint64_t v = p[b]; if (b < 8) key.lo |= v << (8 * b); else key.hi |= v << (8 * (b - 8)); }
    key.lo &= mask.lo; key.hi &= mask.hi;
    uint32_t c = 0;
    for (int b = 0; b < h.counter_len && b < 4; ++b) c |= (uint32_t)p[h.key_bytes() + b] << (8 * b);
    recs[i] = {key, c, key_posx(g, hash_lut(lut.data(), g.key_bytes, key), key)};
    if (i && recs[i].posx < recs[i - 1].posx) return say("posx not monotone at record " + std::to_string(i));
  }
  std::vector<uint64_t> words((size_t)gc.n_units * kCompactSlots, 0);
  std::map<std::pair<uint64_t, uint64_t>, uint32_t> stash;
  const uint32_t cmask = cval_mask(g);
  int64_t prev = -1;
  uint64_t max_disp = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t home = home_unit_of_posx(g, recs[i].posx) * kCompactSlots;
    const uint64_t slot = (int64_t)home > prev ? home : (uint64_t)(prev + 1);
    prev = (int64_t)slot;
    if (slot + 2 >= words.size()) return say("overflow tail exhausted");
    const uint64_t line = slot / kCompactSlots, hl = home / kCompactSlots;
    max_disp = std::max(max_disp, line - hl);
    uint32_t cv = recs[i].count;
    if (line - hl > (uint64_t)kMaxHops) { cv = cmask - 1u; stash[{recs[i].key.hi, recs[i].key.lo}] = recs[i].count; }
    else if (recs[i].count > cmask - 2u) { cv = cmask; stash[{recs[i].key.hi, recs[i].key.lo}] = recs[i].count; }
    words[slot] |= compact_pattern(g, recs[i].posx, recs[i].key) | cv;
    if ((slot & (kCompactSlots - 1)) == 0 && hl < line) words[slot - 1] |= spill_bit(g);
  }
  // every record is found
  long long lines_hit = 0, lines_miss = 0;
  for (uint64_t i = 0; i < n; ++i) {
    bool ns = false; int ln = 0;
    uint32_t c = host_find(g, words, recs[i].key, recs[i].posx, &ns, &ln);
    lines_hit += ln;
    if (ns) { auto it = stash.find({recs[i].key.hi, recs[i].key.lo}); c = it == stash.end() ? 0 : it->second; }
    if (c != recs[i].count) return say("record " + std::to_string(i) + " found with count " + std::to_string(c) + ", file says " + std::to_string(recs[i].count));
  }
  // absent keys (a present key with one bit flipped, unless that is present too)
  std::map<std::pair<uint64_t, uint64_t>, uint32_t> present;
  for (const Rec& r : recs) present[{r.key.hi, r.key.lo}] = r.count;
  uint64_t n_abs = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Key key = recs[i].key;
    const int bit = (int)((i * 7919u) % (unsigned)h.key_bits);
    if (bit < 64) key.lo ^= 1ull << bit; else key.hi ^= 1ull << (bit - 64);
    if (present.count({key.hi, key.lo})) continue;
    bool ns = false; int ln = 0;
    const uint64_t px = key_posx(g, hash_lut(lut.data(), g.key_bytes, key), key);
    uint32_t c = host_find(g, words, key, px, &ns, &ln);
    lines_miss += ln; ++n_abs;
    if (ns) c = 0;                                  // (not in the stash: absent)
    if (c != 0) return say("absent key reported with count " + std::to_string(c));
  }
  // every slot inverts to its key (what jb_db_dump does for a compact table)
  const int r = g.r_bits;
  const std::vector<uint64_t> inv = low_block_inverse(h, r);
  if (inv.empty()) return say("low block singular but compact layout chosen");
  const int bits = 64 - g.left_shift;
  const uint64_t tag_mask = (1ull << g.tag_bits) - 1, rmask = r >= 64 ? ~0ull : (1ull << r) - 1;
  uint64_t k2 = 0;
  for (uint64_t sidx = 0; sidx < words.size(); ++sidx) {
    const uint64_t w = words[sidx];
    if (w == 0) continue;
    if (((uint32_t)w & cmask) == cmask - 1u) { ++k2; continue; }
    const uint64_t pat = w >> g.cnt_bits, tag = pat & tag_mask, rem = pat >> g.tag_bits;
    const uint64_t line = sidx / kCompactSlots, first_line = line > (uint64_t)kMaxHops ? line - kMaxHops : 0;
    const unsigned __int128 span = (unsigned __int128)1 << bits;
    const uint64_t lo_px = (uint64_t)(((unsigned __int128)first_line * span + g.units - 1) / g.units);
    uint64_t px = (lo_px & ~tag_mask) | tag;
    if (px < lo_px) px += tag_mask + 1;
    Key key;
    if (r == 0) { key.lo = rem; key.hi = 0; } else if (r >= 64) { key.lo = 0; key.hi = rem << (r - 64); } else { key.lo = rem << r; key.hi = rem >> (64 - r); }
    const uint64_t v = ((px >> g.ext_bits) ^ hash_lut(lut.data(), g.key_bytes, key)) & rmask;
    uint64_t low = 0;
    for (int b = 0; b < r; ++b) low |= (uint64_t)(__builtin_popcountll(inv[(size_t)b] & v) & 1) << b;
    key.lo |= low;
    // records were placed in order: the k2-th non-empty slot holds record k2
    if (!key_eq(key, recs[k2].key)) return say("slot " + std::to_string(sidx) + " does not invert to record " + std::to_string(k2));
    ++k2;
  }
  if (k2 != n) return say("slot count differs from the record count");
  stats[1] = (long long)g.units; stats[2] = g.tag_bits; stats[3] = g.cnt_bits; stats[4] = (long long)stash.size();
  stats[5] = (long long)max_disp; stats[6] = n ? lines_hit * 1000 / (long long)n : 0; stats[7] = n_abs ? lines_miss * 1000 / (long long)n_abs : 0;
  stats[8] = (long long)n;
  msg[0] = 0;
  return 0;
}

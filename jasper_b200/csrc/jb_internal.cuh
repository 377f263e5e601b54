// jb_internal.cuh -- structures shared by the translation units of libjasper_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/jasper_b200.h"
#include "jb_header.hpp"
#include "jb_kmer.cuh"
#include "jb_index.cuh"

namespace jb {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define JB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return jb::fail(JB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

void note_timing(const char* what, double ms);
void note_counter(const char* what, int64_t v);
void add_counter(const char* what, int64_t v);

// ---- resident table ---------------------------------------------------------------------------
// Layouts and slot arithmetic: jb_kmer.cuh ("table geometry").  A probe unit is one 32-byte sector: four 8-byte slots
// (compact) or two 16-byte slots (wide).
struct __align__(16) Slot {      // wide layout, and the stash of a compact table
  uint64_t lo;
  uint32_t hi;
  uint32_t count;
};

// A table is either whole (n_parts <= 1: every unit behind `base`) or PARTITIONED over the GPUs of one box (SURVEY.md §8e
// fallback for tables beyond one GPU's HBM): part p, resident on its own device, holds units [part_first[p],
// part_first[p+1] + overlap), i.e. its own range plus a copy of the start of the next part, so that a probe that
// starts in part p never leaves part p's memory.  A probe of a foreign part is a peer load over NVLink / NVSwitch.
constexpr int kMaxParts = 8;
constexpr uint64_t kPartOverlapWide = 32768;       // sectors; the loader verifies that every probe run entering it ends inside it
constexpr uint64_t kPartOverlapCompact = 32;       // sectors; a compact probe never reads beyond home + kMaxHops
struct TableView {
  const uint8_t* base;
  const uint64_t* lut;   // key_bytes * 256 entries
  TableGeom g;
  Key mask;
  int n_parts;
  const uint8_t* part[kMaxParts];
  uint64_t part_first[kMaxParts + 1];              // in units
  // compact only: records whose count exceeds the count field, or that were displaced beyond kMaxHops
  const Slot* stash;                               // wide slots, 2 per sector, linear probing; replicated on every device
  uint64_t stash_units;                            // 0 = no stash
  ScanIndex ix;                                    // the records once more, grouped by minimizer, for k_scan (jb_index.cuh); bucket_bits 0 = none
};

// the bytes of probe unit `unit`, wherever they live
__device__ __forceinline__ const uint8_t* unit_ptr(const TableView& t, uint64_t unit) {
  if (t.n_parts <= 1) return t.base + (unit << 5);
  int p = 0;
#pragma unroll 1
  for (int q = 1; q < t.n_parts; ++q) p += unit >= t.part_first[q] ? 1 : 0;      // (rare path: keep it small, no unrolling)
  return t.part[p] + ((unit - t.part_first[p]) << 5);
}

__device__ __forceinline__ void ldg256(const void* p, unsigned long long& w0, unsigned long long& w1, unsigned long long& w2, unsigned long long& w3) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p));
}

// Wide probe: starting at the home sector, scan slots until the key or an empty slot.
// 256-bit loads (LDG.E.256) fetch exactly one 32-byte sector per step.
__device__ __forceinline__ uint32_t wide_find(const ulonglong4* sec, const Key& key, int* steps = nullptr) {
  int n = 0;
  while (true) {
    unsigned long long w0, w1, w2, w3;
    ldg256(sec, w0, w1, w2, w3);
    ++n;
    uint32_t c0 = (uint32_t)(w1 >> 32), c1 = (uint32_t)(w3 >> 32);
    uint32_t r;
    bool done = true;
    if (c0 == 0) r = 0;
    else if (w0 == key.lo && (uint32_t)w1 == (uint32_t)key.hi) r = c0;
    else if (c1 == 0) r = 0;
    else if (w2 == key.lo && (uint32_t)w3 == (uint32_t)key.hi) r = c1;
    else { done = false; r = 0; }
    if (done) { if (steps) *steps = n; return r; }
    ++sec;
  }
}

// the stash of a compact table: a small wide table over the same posx
static __device__ __noinline__ uint32_t stash_find(const TableView& t, const Key& key, uint64_t posx) {
  if (t.stash_units == 0) return 0;
  const uint64_t sec = mulhi64(posx << t.g.left_shift, t.stash_units);
  return wide_find(reinterpret_cast<const ulonglong4*>(t.stash) + sec, key);
}

// Compact probe.  One step = one sector = four slots {pattern | count field}; x = slot ^ pattern is below 2^cnt_bits
// exactly when remainder and tag agree (then x is the count field).  Keys are unique, so at most one live slot matches; a
// tombstone (cval cmask - 1) never does.  The occupied slots of a sector are a prefix, and its last slot carries the
// "run continues" flag: a sector that is not full, or full without the flag, ends the probe.
struct CompactProbe {
  unsigned long long pat, lim;
  uint32_t cmask;
};
__device__ __forceinline__ CompactProbe compact_probe_of(const TableGeom& g, uint64_t posx, const Key& key) {
  CompactProbe p;
  p.pat = compact_pattern(g, posx, key); p.lim = 1ull << g.cnt_bits; p.cmask = cval_mask(g);
  return p;
}
// -> true when the probe ends in this sector; *cv = the matching slot's cval (0 = absent)
__device__ __forceinline__ bool compact_step(const CompactProbe& p, unsigned long long w0, unsigned long long w1, unsigned long long w2,
                                             unsigned long long w3, uint32_t* cv) {
  const unsigned long long x0 = w0 ^ p.pat, x1 = w1 ^ p.pat, x2 = w2 ^ p.pat, x3 = w3 ^ p.pat;
  // cval of every slot that agrees on remainder and tag (an EMPTY slot agrees with the all-zero pattern and yields 0).  Only
  // the last slot can carry the spill bit.  Live counts are below the tombstone value, the escape value is above it.
  const uint32_t m0 = x0 < p.lim ? (uint32_t)x0 : 0u, m1 = x1 < p.lim ? (uint32_t)x1 : 0u, m2 = x2 < p.lim ? (uint32_t)x2 : 0u;
  const uint32_t m3 = x3 < p.lim ? (uint32_t)x3 & p.cmask : 0u;
  uint32_t hit = max(max(m0, m1), max(m2, m3));
  if (hit == p.cmask - 1u) {       // a tombstone of another home in reach agrees: a live slot, if one agrees too, wins
    const uint32_t tomb = p.cmask - 1u;
    hit = max(max(m0 == tomb ? 0u : m0, m1 == tomb ? 0u : m1), max(m2 == tomb ? 0u : m2, m3 == tomb ? 0u : m3));
  }
  *cv = hit;
  return hit != 0 || w3 == 0 || !(w3 & (p.lim >> 1));
}

__device__ __forceinline__ uint32_t compact_find(const TableView& t, const Key& key, uint64_t posx, int* steps = nullptr) {
  const CompactProbe p = compact_probe_of(t.g, posx, key);
  const ulonglong4* sec = reinterpret_cast<const ulonglong4*>(unit_ptr(t, home_unit_of_posx(t.g, posx)));
  int n = 0;
#pragma unroll 1
  for (int hop = 0; hop <= kMaxHops; ++hop, ++sec) {
    unsigned long long w0, w1, w2, w3;
    ldg256(sec, w0, w1, w2, w3);
    ++n;
    uint32_t cv;
    if (compact_step(p, w0, w1, w2, w3, &cv)) {
      if (steps) *steps = n;
      return cv == p.cmask ? stash_find(t, key, posx) : cv;
    }
  }
  if (steps) *steps = n;
  return stash_find(t, key, posx);          // kMaxHops + 1 full sectors: the record, if any, is a tombstone
}

__device__ __forceinline__ uint32_t table_find(const TableView& t, const Key& key, uint64_t posx, int* steps = nullptr) {
  if (t.g.layout == kLayoutCompact) return compact_find(t, key, posx, steps);
  return wide_find(reinterpret_cast<const ulonglong4*>(unit_ptr(t, home_unit_of_posx(t.g, posx))), key, steps);
}

__device__ __forceinline__ uint32_t table_count(const TableView& t, const Key& canonical) {
  uint64_t pos = hash_lut(t.lut, t.g.key_bytes, canonical);
  return table_find(t, canonical, key_posx(t.g, pos, canonical));
}

}  // namespace jb

struct jb_db;
namespace jb {
int build_scan_index(jb_db* db);       // jb_index.cu: JB_OK also when no index is built (switched off, does not fit, partitioned table)
void free_scan_index(jb_db* db);
}
int jb_check_ready(const jb_db* db);   // JB_OK, or an error while a part-wise opened table still lacks its other parts

struct jb_db {
  jb::JfHeader hdr;
  int device = 0;
  uint8_t* mem = nullptr;      // this handle's units: the whole table, or its part (+ overlap) of a partitioned one
  uint64_t n_units = 0;        // of the WHOLE table: home range + overflow tail + guard
  uint64_t home_units = 0;     // units the home function maps to (= geom.units)
  uint64_t* lut = nullptr;     // device
  jb::TableGeom geom{};
  uint64_t n_records = 0;
  uint64_t max_disp = 0;       // longest displacement of a record from its home, in slots
  double load_seconds = 0, build_ms = 0;
  double read_seconds = 0;     // busiest reader thread's time inside read()
  uint64_t file_bytes = 0;
  cudaStream_t stream = nullptr;
  // compact layout: the stash (jb_kmer.cuh) as a wide mini-table, and the records in it as a list (dump / histogram)
  jb::Slot* stash = nullptr;
  uint64_t stash_units = 0;
  jb::Slot* stash_list = nullptr;
  uint64_t n_stash = 0;
  // partitioned tables (jb_db_open_partitioned): this handle owns part `my_part`; the others are peer memory
  int n_parts = 1, my_part = 0;
  bool parts_ready = true;             // false between jb_db_open_part and jb_db_attach_parts
  bool ipc_mapped[jb::kMaxParts] = {};
  uint64_t part_overlap = 0;           // units
  uint8_t* parts[jb::kMaxParts] = {};
  uint64_t part_first[jb::kMaxParts + 1] = {};     // units
  // the scan index (jb_index.cuh)
  void* idx_slots = nullptr; void* idx_off = nullptr;
  int idx_bits = 0, idx_m = 0;
  uint64_t idx_n_slots = 0;
  bool idx_wide_off = false;
  double idx_seconds = 0;
  int unit_bytes() const { return 32; }
  uint64_t my_units() const {          // units resident behind `mem`
    return n_parts > 1 ? part_first[my_part + 1] - part_first[my_part] + (my_part + 1 < n_parts ? part_overlap : 0) : n_units;
  }
  uint64_t table_bytes() const { return n_units * (uint64_t)unit_bytes() + stash_units * 32; }
  jb::TableView view() const {
    jb::TableView v;
    v.base = mem; v.lut = lut; v.g = geom; v.mask = jb::key_mask(hdr.k());
    v.n_parts = n_parts;
    for (int q = 0; q < jb::kMaxParts; ++q) v.part[q] = parts[q];
    for (int q = 0; q <= jb::kMaxParts; ++q) v.part_first[q] = part_first[q];
    v.stash = stash; v.stash_units = stash_units;
    v.ix.slots = reinterpret_cast<const uint4*>(idx_slots);
    v.ix.off32 = idx_wide_off ? nullptr : reinterpret_cast<const uint32_t*>(idx_off);
    v.ix.off64 = idx_wide_off ? reinterpret_cast<const uint64_t*>(idx_off) : nullptr;
    v.ix.bucket_bits = idx_bits; v.ix.m = idx_m;
    return v;
  }
};

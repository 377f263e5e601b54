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
// 16-byte slot {key.lo, key.hi : 32 | count : 32}; count == 0 <=> empty.  Two slots = one
// 32-byte sector, the unit a probe loads (32 algorithmic bytes per lookup).
struct __align__(16) Slot {
  uint64_t lo;
  uint32_t hi;
  uint32_t count;
};

// A table is either whole (n_parts <= 1: every slot in `slots`) or PARTITIONED over the GPUs of one box (SURVEY.md §8e
// fallback for tables beyond one GPU's HBM): part p, resident on its own device, holds slots [part_first[p],
// part_first[p+1] + kPartOverlap), i.e. its own range plus a copy of the start of the next part, so that a probe that
// starts in part p never leaves part p's memory.  A probe of a foreign part is a peer load over NVLink / NVSwitch.
constexpr int kMaxParts = 8;
constexpr uint64_t kPartOverlap = 65536;     // slots; the loader verifies that every probe run entering it ends inside it
struct TableView {
  const Slot* slots;
  const uint64_t* lut;   // key_bytes * 256 entries
  TableGeom g;
  Key mask;
  int n_parts;
  const Slot* part[kMaxParts];
  uint64_t part_first[kMaxParts + 1];
};

// the 32-byte sector of slot index `home` (even), wherever it lives
__device__ __forceinline__ const ulonglong4* home_sector(const TableView& t, uint64_t home) {
  if (t.n_parts <= 1) return reinterpret_cast<const ulonglong4*>(t.slots) + (home >> 1);
  int p = 0;
#pragma unroll 1
  for (int q = 1; q < t.n_parts; ++q) p += home >= t.part_first[q] ? 1 : 0;      // (rare path: keep it small, no unrolling)
  return reinterpret_cast<const ulonglong4*>(t.part[p]) + ((home - t.part_first[p]) >> 1);
}

// Probe: starting at the sector of the home slot, scan slots until the key or an empty slot.
// 256-bit loads (LDG.E.256) fetch exactly one 32-byte sector per step.
__device__ __forceinline__ uint32_t table_find(const TableView& t, const Key& key, uint64_t home, int* steps = nullptr) {
  const ulonglong4* sec = home_sector(t, home);
  int n = 0;
  while (true) {
    unsigned long long w0, w1, w2, w3;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(sec));
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

__device__ __forceinline__ uint32_t table_count(const TableView& t, const Key& canonical) {
  uint64_t pos = hash_lut(t.lut, t.g.key_bytes, canonical);
  return table_find(t, canonical, home_slot(t.g, pos, canonical));
}

}  // namespace jb

struct jb_db;
int jb_check_ready(const jb_db* db);   // JB_OK, or an error while a part-wise opened table still lacks its other parts

struct jb_db {
  jb::JfHeader hdr;
  int device = 0;
  jb::Slot* slots = nullptr;
  uint64_t n_slots = 0;        // cap + tail + guard
  uint64_t cap = 0;
  uint64_t* lut = nullptr;     // device
  jb::TableGeom geom{};
  uint64_t n_records = 0;
  uint64_t max_disp = 0;
  double load_seconds = 0, build_ms = 0;
  cudaStream_t stream = nullptr;
  // partitioned tables (jb_db_open_partitioned): this handle owns part `my_part`; the others are peer memory
  int n_parts = 1, my_part = 0;
  bool parts_ready = true;             // false between jb_db_open_part and jb_db_attach_parts
  bool ipc_mapped[jb::kMaxParts] = {};
  uint64_t part_overlap = 0;
  jb::Slot* parts[jb::kMaxParts] = {};
  uint64_t part_first[jb::kMaxParts + 1] = {};
  jb::TableView view() const {
    jb::TableView v;
    v.slots = slots; v.lut = lut; v.g = geom; v.mask = jb::key_mask(hdr.k());
    v.n_parts = n_parts;
    for (int q = 0; q < jb::kMaxParts; ++q) v.part[q] = parts[q];
    for (int q = 0; q <= jb::kMaxParts; ++q) v.part_first[q] = part_first[q];
    return v;
  }
};

// jb_scan.cuh -- K1/K2/K3: the bulk "count every window" kernel and the event bitmasks derived from it.
//
// k_scan: for every position p of every contig, counts[p] = C(seq[p:p+k]) exactly as
// qf[jf.MerDNA(seq[p:p+k]).get_canonical()] would return it (src/jasper.py:70-71; MerDNA padding
// semantics JF!include/jellyfish/mer_dna.hpp:525-542,642-644), for p in [0, len] (the walker also
// queries the empty tail window), plus one validity bit per position (window is k chars of ACGTacgt).
//
// k_events: from those counts, the bitmasks the walker hops over:
//   event bit p = 0 iff the reference's scan loop, when it visits p, takes the plain `i += k-1`
//                 branch (src/jasper.py:55-100): window all ACGTacgt, count >= solid_thre, and not
//                 (i > 0 and count < count(k bases before)/50).  Positions p < k are always flagged so
//                 the walker evaluates the `max(0,i-k)` / `max(k,i)` clamps of src/jasper.py:80 itself.
//   weak bit p  = window all ACGT and count < solid_thre (anchor search, jb_segments.cuh)
//   l1 bit b    = any event among flat bit positions [1024 b, 1024 b + 1024)
//
// k_scan layout (r1 -> r2 redesign after ncu: 33 % of the old kernel's stall samples were __syncthreads,
// 16 warps/SM): NO block-wide barrier in the steady state.  A persistent CTA loads the hash LUT once;
// then every WARP works on its own tile of 1024 positions: stages the tile's bases (+k-1 halo) in its
// private slice of shared memory, each lane rolls 32 consecutive windows (forward and reverse-complement
// keys updated incrementally, canonical = min), hashes with the header matrix as byte LUTs in shared
// memory, keeps kGroup independent probes in flight (one 32-byte sector load each; the DRAM fetch behind
// it is a 128-byte line, tools/gather_probe.cu), stages the counts in shared memory and streams them out
// coalesced.  Warps drift apart, so probe latency of one overlaps key arithmetic of another.
#pragma once
#include "jb_internal.cuh"
#include "jb_jobs.cuh"

namespace jb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kLanePos = 32;                        // consecutive positions per lane (one output word per lane)
constexpr int kTile = 32 * kLanePos;                // 1024 positions per warp tile
#ifndef JB_SCAN_GROUP
#define JB_SCAN_GROUP 4
#endif
#ifndef JB_SCAN_MINBLOCKS
#define JB_SCAN_MINBLOCKS 3
#endif
#ifndef JB_SCAN_GROUP_COMPACT
#define JB_SCAN_GROUP_COMPACT 3                     // (r2c: 4 in flight spill in the compact loop: 2.28 ms against 1.88)
#endif
#ifndef JB_SCAN_GROUP_INDEX
#define JB_SCAN_GROUP_INDEX 3
#endif
constexpr int kScanByIndex = 2;                     // k_scan<kScanByIndex>: lookups through the scan index (jb_index.cuh) instead of a table layout
constexpr int kGroupIndex = JB_SCAN_GROUP_INDEX;
constexpr int kGroup = JB_SCAN_GROUP;               // probes in flight per lane, wide table
constexpr int kGroupCompact = JB_SCAN_GROUP_COMPACT;
constexpr int kScanCtasPerSm = JB_SCAN_MINBLOCKS;
// 1: tiles arrive by cp.async.bulk (TMA) into a double buffer with an mbarrier per buffer, the next tile in flight while the
// current one is probed.  Measured against the per-lane staging on one box (profiles/r2g_scan_tma_ab.txt): k_scan 2.32 ms
// against 1.87 ms -- the second tile buffer takes the shared-memory carve-out from 196 KB to 228 KB per SM and the staging was
// never on the critical path (a warp spends ~1 us of a ~35 us tile on it).  Kept as a build option, off by default.
#ifndef JB_SCAN_TMA
#define JB_SCAN_TMA 0
#endif
constexpr int kSeqSlice = kTile + JB_MAX_K + 16;    // staged bases per warp (multiple of 16)
constexpr int kSeqBufs = JB_SCAN_TMA ? 2 : 1;       // TMA: double-buffered
constexpr int kCntSlice = 33 * 32;                  // counts, row-padded against bank conflicts
constexpr size_t kScanSmemPerWarp = (size_t)kSeqBufs * kSeqSlice + (size_t)kCntSlice * 4 + (JB_SCAN_TMA ? 16 : 0);   // (+ two mbarriers)

struct ScanUnit {          // one contig
  uint64_t seq_off;        // first base in the flat base buffer
  uint64_t len;            // bases
  uint64_t out_off;        // first count / first bit (multiple of 32) in the flat outputs
};
struct ScanTile {
  uint32_t unit;
  uint32_t pad;
  uint64_t start;          // first position of the tile inside the unit (multiple of kTile)
};

// tiles are numbered contig by contig: contig u owns tiles [tile_first[u], tile_first[u+1]),
// tile_first[u+1] - tile_first[u] = ceil((len_u + 1) / kTile)
__host__ __device__ inline uint32_t tiles_of(uint64_t len) { return (uint32_t)((len + 1 + kTile - 1) / kTile); }

__device__ __forceinline__ ScanTile decode_tile(const uint32_t* __restrict__ tile_first, uint32_t n_units, uint32_t ti) {
  uint32_t lo = 0, hi = n_units;                    // last u with tile_first[u] <= ti
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (tile_first[mid] <= ti) lo = mid; else hi = mid; }
  return ScanTile{lo, 0, (uint64_t)(ti - tile_first[lo]) * kTile};
}

struct ScanArgs {
  const uint8_t* bases;
  const ScanUnit* units;
  const uint32_t* tile_first;   // n_units + 1 entries
  uint32_t n_units;
  uint32_t n_tiles;
  uint32_t* counts;             // flat, counts[out_off + p]
  uint8_t* okb;                 // flat, 1 = window is k chars of ACGTacgt; may be null
  unsigned long long* probe_steps;   // optional statistics: total sector loads
};

inline size_t scan_smem_bytes(int key_bytes) { return (size_t)key_bytes * 256 * 8 + kScanWarps * kScanSmemPerWarp; }
// index scan: no hash LUT, but a ring of m-mer hashes per thread (k - m + 1 words, odd stride)
__host__ __device__ inline int scan_ring_stride(int k, int m) { return (k - m + 1) | 1; }
inline size_t scan_smem_bytes_index(int k, int m) { return kScanWarps * kScanSmemPerWarp + (size_t)kScanThreads * scan_ring_stride(k, m) * 4; }

struct EventArgs {
  const ScanUnit* units;
  const uint32_t* tile_first;
  uint32_t n_units;
  uint32_t n_tiles;
  const uint32_t* counts;
  const uint8_t* okb;
  uint32_t* events;
  uint32_t* weak;
  uint32_t* events_l1;      // pre-zeroed
  uint32_t* events_l2;      // pre-zeroed: bit w = level-1 word w is not empty
  uint32_t solid_thre;
  int k;
};

// Incremental recount after a fixing pass (passes 2.. never rescan everything): the copy jobs that stitched
// the new contigs say which stretches are untouched, so their windows keep their counts; only windows that
// touch edited bases, straddle a job boundary or run off a contig end are probed again.
struct CopyJob;
struct RecountArgs {
  const CopyJob* jobs;
  const unsigned long long* n_jobs;
  unsigned long long jobs_cap;  // arena size: the walkers' counter may run past it (the host then fails the call)
  const int64_t* dst_off;       // by launch slot: flat offset of the segment's new bases; < 0 = superseded
  const uint32_t* slot_unit;    // by launch slot: contig
  const ScanUnit* units_new;    // per contig: seq_off (= flat offset, unchanged between passes) and the NEW length
  const uint8_t* in_buf;        // the pass-start flat base buffer (jobs whose src lies in it are untouched stretches)
  uint64_t flat_bytes;
  const uint8_t* new_bases;     // the stitched flat base buffer
  const uint32_t* cnt_old; const uint8_t* okb_old;
  uint32_t* cnt_new; uint8_t* okb_new;
  unsigned long long* probes;   // statistics
  int force_rows;               // tests: never split into teams (jb_jobs.cuh, job_split)
};
struct RecountTailArgs {
  const ScanUnit* units_old; const ScanUnit* units_new; uint32_t n_units;
  const uint32_t* cnt_old; uint32_t* cnt_new; uint8_t* okb_new;
};

// launchers (jb_scan.cu)
// number of positions p in [0, len_u] of any contig where (count, ok) differ between two flat arrays (test aid)
int compare_counts(const ScanUnit* d_units, uint32_t n_units, const uint32_t* cnt_a, const uint8_t* okb_a, const uint32_t* cnt_b,
                   const uint8_t* okb_b, unsigned long long* d_scratch, unsigned long long* mismatches, cudaStream_t s);
int launch_recount(const TableView& view, const RecountArgs& a, const RecountTailArgs& t, cudaStream_t s);
int scan_grid(uint32_t n_tiles);
int launch_scan(const TableView& view, const ScanArgs& a, cudaStream_t s);
int launch_events(const EventArgs& a, cudaStream_t s);
std::vector<uint32_t> make_tile_prefix(const std::vector<ScanUnit>& units);

}  // namespace jb

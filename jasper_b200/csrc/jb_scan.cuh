// jb_scan.cuh -- K1/K2/K3 fused: the bulk "count every window" kernel.
//
// For every position p of every contig it computes C(seq[p:p+k]) exactly as
// qf[jf.MerDNA(seq[p:p+k]).get_canonical()] would (src/jasper.py:70-71; MerDNA padding semantics
// JF!include/jellyfish/mer_dna.hpp:525-542,642-644), and from those counts either
//   * writes them out (jb_count_positions), and/or
//   * writes the EVENT bitmask the walker hops over: bit p is 0 iff the reference's scan loop, when
//     it visits p, takes the plain `i += k-1` branch (src/jasper.py:55-100): the window is all
//     ACGTacgt, count >= solid_thre, and not (i > 0 and count < count(k bases before)/50).
//     Positions p < k of a contig are always flagged so the walker evaluates the `max(0,i-k)`
//     / `max(k,i)` clamps of src/jasper.py:80 itself.
//
// Layout: a persistent grid (one CTA per SM slot) loops over tiles of kTile positions; the tile's
// bases (+ k-1 halo, + k positions of look-back for the count(k bases before) rule) are staged in
// shared memory, each thread rolls kPerThread consecutive windows (2-bit pack, reverse complement,
// canonical min), hashes with the header matrix as byte LUTs held in shared memory, and probes the
// table with one 32-byte sector load per step, kBatch probes in flight per thread.
#pragma once
#include "jb_internal.cuh"

namespace jb {

constexpr int kScanThreads = 256;
constexpr int kPerThread = 8;
constexpr int kTile = kScanThreads * kPerThread;   // 2048 positions
constexpr int kBatch = 4;

struct ScanUnit {          // one contig
  uint64_t seq_off;        // first base in the flat base buffer
  uint64_t len;            // bases
  uint64_t out_off;        // first count / first event bit (multiple of 32) in the flat outputs
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
  uint32_t* counts;        // may be null
  uint32_t* events;        // may be null (bit words)
  uint32_t* weak;          // with events: bit p = window all-ACGT and count < solid_thre ("pure weak")
  uint32_t* events_l1;     // with events: bit b = any event among flat bit positions [1024 b, 1024 b + 1024); pre-zeroed
  uint32_t solid_thre;
  unsigned long long* probe_steps;   // optional statistics: total sector loads
};

template <bool kCounts, bool kEvents>
__global__ void __launch_bounds__(kScanThreads)
k_scan(TableView t, ScanArgs a) {
  constexpr int kMaxRegion = kTile + JB_MAX_K;            // positions whose count we need
  constexpr int kMaxChars = kMaxRegion + JB_MAX_K + 16;
  __shared__ uint64_t s_lut[kMaxKeyBytes * 256];
  __shared__ __align__(16) uint8_t s_seq[kMaxChars];
  __shared__ uint32_t s_cnt[kEvents ? kMaxRegion : 1];
  __shared__ uint8_t s_ok[kEvents ? kMaxRegion : 1];
  const int k = t.g.k;
  const int back = kEvents ? k : 0;
  for (int i = threadIdx.x; i < t.g.key_bytes * 256; i += kScanThreads) s_lut[i] = t.lut[i];
  unsigned long long steps_total = 0;

  for (uint32_t ti = blockIdx.x; ti < a.n_tiles; ti += gridDim.x) {
    const ScanTile tile = decode_tile(a.tile_first, a.n_units, ti);
    const ScanUnit u = a.units[tile.unit];
    // positions [tile.start, min(start + kTile, len + 1)) are produced (len + 1: the walker also
    // queries the empty tail window); the region additionally covers `back` positions before.
    const int64_t reg0 = (int64_t)tile.start - back;              // may be negative
    const int64_t n_out = (int64_t)min((uint64_t)kTile, u.len + 1 - tile.start);
    const int n_reg = (int)n_out + back;
    const int n_chars = n_reg + k - 1;
    __syncthreads();                                              // previous tile done with smem
    const uint8_t* src = a.bases + u.seq_off;
    for (int i = threadIdx.x; i < n_chars + 1; i += kScanThreads) {
      int64_t p = reg0 + i;
      s_seq[i] = (p >= 0 && p < (int64_t)u.len && i < n_chars) ? src[p] : 0;
    }
    __syncthreads();

    for (int chunk = threadIdx.x; chunk * kPerThread < n_reg; chunk += kScanThreads) {
      const int r0 = chunk * kPerThread;
      const int nr = min(kPerThread, n_reg - r0);
      Key keys[kPerThread];
      bool ok[kPerThread];
      // first window
      Key fwd = {0, 0};
      int nv = 0;
      for (int j = 0; j < k; ++j) {
        uint8_t c = s_seq[r0 + j];
        bool v = base_valid(c);
        nv = v ? nv + 1 : 0;
        fwd = key_push(fwd, v ? base_code(c) : 0u, t.mask);
      }
#pragma unroll
      for (int q = 0; q < kPerThread; ++q) {
        if (q < nr) {
          if (q > 0) {
            uint8_t c = s_seq[r0 + q + k - 1];
            bool v = base_valid(c);
            nv = v ? nv + 1 : 0;
            fwd = key_push(fwd, v ? base_code(c) : 0u, t.mask);
          }
          ok[q] = nv >= k;
          Key m = ok[q] ? fwd : key_from_chars(s_seq + r0 + q, k, k, t.mask);
          keys[q] = key_canonical(m, k);
        } else { ok[q] = false; keys[q] = Key{0, 0}; }
      }
      uint32_t cnt[kPerThread];
#pragma unroll
      for (int b0 = 0; b0 < kPerThread; b0 += kBatch) {
        const ulonglong4* sec[kBatch];
        unsigned long long w[kBatch][4];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          uint64_t pos = hash_lut(s_lut, t.g.key_bytes, keys[b0 + b]);
          sec[b] = reinterpret_cast<const ulonglong4*>(t.slots) + (home_slot(t.g, pos, keys[b0 + b]) >> 1);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b)
          if (b0 + b < nr)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          uint32_t r = 0;
          if (b0 + b < nr) {
            const Key key = keys[b0 + b];
            ++steps_total;
            while (true) {
              uint32_t c0 = (uint32_t)(w[b][1] >> 32), c1 = (uint32_t)(w[b][3] >> 32);
              if (c0 == 0) { r = 0; break; }
              if (w[b][0] == key.lo && (uint32_t)w[b][1] == (uint32_t)key.hi) { r = c0; break; }
              if (c1 == 0) { r = 0; break; }
              if (w[b][2] == key.lo && (uint32_t)w[b][3] == (uint32_t)key.hi) { r = c1; break; }
              ++sec[b];
              ++steps_total;
              asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                           : "=l"(w[b][0]), "=l"(w[b][1]), "=l"(w[b][2]), "=l"(w[b][3]) : "l"(sec[b]));
            }
          }
          cnt[b0 + b] = r;
        }
      }
#pragma unroll
      for (int q = 0; q < kPerThread; ++q) {
        if (q < nr) {
          if (kEvents) { s_cnt[r0 + q] = cnt[q]; s_ok[r0 + q] = ok[q]; }
          if (kCounts && r0 + q >= back) a.counts[u.out_off + tile.start + (r0 + q - back)] = cnt[q];
        }
      }
    }
    if (kEvents) {
      __syncthreads();
      for (int q = threadIdx.x; q < kTile; q += kScanThreads) {
        bool ev = false, wk = false;
        if (q < n_out) {
          const uint64_t p = tile.start + q;
          const uint32_t c = s_cnt[q + back];
          wk = s_ok[q + back] && c < a.solid_thre;
          ev = !s_ok[q + back] || p < (uint64_t)k || c < a.solid_thre;
          if (!ev) ev = (unsigned long long)c * JB_DIVISOR < (unsigned long long)s_cnt[q];   // count at p - k
        }
        const unsigned int word = __ballot_sync(0xffffffffu, ev);
        const unsigned int wword = __ballot_sync(0xffffffffu, wk);
        if ((threadIdx.x & 31) == 0 && q < n_out) {
          const uint64_t f = u.out_off + tile.start + q;       // flat bit index, multiple of 32
          a.events[f >> 5] = word;
          a.weak[f >> 5] = wword;
          if (word) atomicOr(&a.events_l1[f >> 15], 1u << ((f >> 10) & 31));
        }
      }
    }
  }
  if (a.probe_steps) {
    for (int o = 16; o; o >>= 1) steps_total += __shfl_xor_sync(0xffffffffu, steps_total, o);
    if ((threadIdx.x & 31) == 0 && steps_total) atomicAdd(a.probe_steps, steps_total);
  }
}

}  // namespace jb

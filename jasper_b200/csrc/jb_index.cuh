// jb_index.cuh -- the scan index: the records of the table once more, laid out for the ONE access pattern that dominates
// the path, k_scan's "every window of a contig in order" (src/jasper.py:56,70-71 over a whole sequence).
//
// The hash table (jb_kmer.cuh) answers a lookup with one random 32-byte sector, but a random sector costs a whole 128-byte
// DRAM line on this part (tools/gather_probe.cu), and consecutive windows of a contig hash to unrelated lines: 128 bytes of
// DRAM traffic per window, 4x the algorithmic 32.  Consecutive windows do share something, though: their MINIMIZER, the
// m-mer (m < k) with the smallest hash among the k-m+1 m-mers of the window, taken strand-symmetrically.  On random
// sequence a minimizer survives (k-m+2)/2 consecutive windows on average (9 at k = 37, m = 21).  So the index groups the
// records by minimizer:
//     bucket(key)  = mix(min over the m-mers x of key of hash(min(x, revcomp x)))  >> (32 - bucket_bits)
//     off[b]       = first slot of bucket b (exclusive prefix sums of the bucket capacities)
//     a bucket     = a small open-addressing table of its own: an even number (records * 5/4 + 2) of sixteen-byte slots
//                    {key.lo : 64, key.hi : 32, count : 32}, probed a PAIR (one 32-byte sector) at a time:
//                    pair = mulhi32(hash2(key), pairs), linear probing with wrap-around inside the bucket, count == 0 =
//                    empty (every bucket keeps at least one empty slot; records fill a pair front to back)
// A lookup is one offset pair (16 M x 4 bytes at the bench size: resident in the 126 MB L2, and a lane that walks a contig
// keeps the pair in registers while the minimizer stays) and one slot -- still a single independent load per window, like
// the hash probe, so the probes of several windows stay in flight together; but the ~9 windows of a minimizer run land in
// the same ~3 lines, of which only the first touch goes to DRAM.
// Only k_scan uses it (random single lookups -- verdicts, walkers, the recount -- stay on the hash table: computing a
// minimizer per lookup costs more than the line it saves).  Built on the device from the resident table right after the
// load when it fits beside it (jb_index.cu); results are the table's by construction (same records) and by test
// (tests/test_gpu.py: every golden and fuzz case with and without the index).
#pragma once
#include "jb_kmer.cuh"

namespace jb {

struct ScanIndex {
  const uint4* slots;        // {key.lo low, key.lo high, key.hi, count}
  const uint32_t* off32;     // n_buckets + 1 entries; null when the table has 2^32 slots or more
  const uint64_t* off64;     // ... then this one
  int bucket_bits;           // 0 = no index
  int m;                     // minimizer length
};

JB_HD int index_mmer_len(int k) { return k >= 35 ? 21 : (k >= 27 ? k - 14 : 13); }   // no index below k = 15

// 64 -> 32 bit finalizer (murmur3's fmix64, truncated): the order it induces on m-mers picks the minimizer
JB_HD uint32_t mmer_hash(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
JB_HD uint64_t key_bits_at(const Key& key, int shift, uint64_t mask) {      // (key >> shift) & mask, shift < 128, mask < 2^64
  uint64_t v;
  if (shift >= 64) v = key.hi >> (shift - 64);
  else if (shift == 0) v = key.lo;
  else v = (key.lo >> shift) | (key.hi << (64 - shift));
  return v & mask;
}
// hash of the m-mer that starts t bases into the window (fwd = the window, rc = its reverse complement)
JB_HD uint32_t mmer_hash_at(const Key& fwd, const Key& rc, int k, int m, int t, uint64_t mmask) {
  const uint64_t x = key_bits_at(fwd, 2 * (k - m - t), mmask), y = key_bits_at(rc, 2 * t, mmask);
  return mmer_hash(x < y ? x : y);
}
// smallest m-mer hash of the window and the offset of the LAST m-mer that attains it
JB_HD uint32_t window_min_hash(const Key& fwd, const Key& rc, int k, int m, int* at) {
  const uint64_t mmask = (1ull << (2 * m)) - 1;
  uint32_t best = 0xFFFFFFFFu;
  int where = 0;
  for (int t = 0; t <= k - m; ++t) {
    const uint32_t h = mmer_hash_at(fwd, rc, k, m, t, mmask);
    if (h <= best) { best = h; where = t; }
  }
  if (at) *at = where;
  return best;
}
JB_HD uint64_t index_bucket(uint32_t min_hash, int bucket_bits) { return (uint32_t)(min_hash * 0x9E3779B1u) >> (32 - bucket_bits); }
JB_HD uint32_t index_slot_hash(const Key& canonical) {
  uint64_t x = canonical.lo ^ (canonical.hi * 0x9E3779B97F4A7C15ull);
  x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32;
  return (uint32_t)x;
}
JB_HD uint64_t index_bucket_capacity(uint64_t records) { return records ? (records + records / 4 + 2) & ~1ull : 0; }   // slots, even

#if defined(__CUDACC__)
__device__ __forceinline__ void index_bounds(const ScanIndex& ix, uint64_t bucket, uint64_t* beg, uint32_t* cap) {
  if (ix.off32) { const uint32_t a = __ldg(ix.off32 + bucket), b = __ldg(ix.off32 + bucket + 1); *beg = a; *cap = b - a; }
  else { const uint64_t a = __ldg(ix.off64 + bucket), b = __ldg(ix.off64 + bucket + 1); *beg = a; *cap = (uint32_t)(b - a); }
}
// generic (slow-path) lookup: minimizer from scratch
static __device__ __noinline__ uint32_t index_find(const ScanIndex& ix, const Key& canonical, int k) {
  const Key rc = key_revcomp(canonical, k);
  const uint64_t bucket = index_bucket(window_min_hash(canonical, rc, k, ix.m, nullptr), ix.bucket_bits);
  uint64_t beg; uint32_t cap;
  index_bounds(ix, bucket, &beg, &cap);
  if (cap == 0) return 0;
  uint32_t s = 2 * __umulhi(index_slot_hash(canonical), cap / 2);
  while (true) {
    const uint4 r = __ldg(ix.slots + beg + s);
    if (r.w == 0) return 0;
    if ((((uint64_t)r.y << 32) | r.x) == canonical.lo && r.z == (uint32_t)canonical.hi) return r.w;
    s = s + 1 == cap ? 0 : s + 1;
  }
}
#endif

}  // namespace jb

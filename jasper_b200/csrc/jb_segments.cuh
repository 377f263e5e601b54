// jb_segments.cuh -- exact decomposition of one contig's sequential scan into independent segments.
//
// The reference's scan (src/jasper.py:55-100) is sequential only through the PHASE of its k-1
// stride.  An ANCHOR is a position a such that
//   (1) windows a .. a+k-2 are all "pure weak" (all-ACGT, count < solid_thre), and
//   (2) none of the `margin` positions before a raises an event.
// A scan arriving from the left in untouched sequence lands on one of a .. a+k-2 whatever its phase
// (stride k-1 cannot jump a run of k-1), finds count < solid_thre there, and calls handle_bad_kmers
// (src/jasper.py:150-223), whose backward walk stops at a-1 (count >= solid_thre) and whose forward
// walk crosses the same weak positions from any landing point: the state after that call does not
// depend on the phase.  So a walker may START at an anchor (as if it had landed on `a`) and the walker
// before it may STOP when it lands in the anchor run; (2) keeps the previous walker's edits away
// from the seam.  Both conditions are re-checked at run time by the walker (status kSeamViolation
// makes the host re-run that contig as one segment), so the result is exact, not speculative.
#pragma once
#include <stdint.h>

#include "jb_kmer.cuh"

namespace jb {

constexpr int kSeamViolation = -101;      // walker status: seam assumptions did not hold
// static margin: event-free positions required before an anchor (a heuristic that makes seam violations
// rare; exactness rests on the walkers' run-time checks, so tests shrink it to provoke re-runs)
JB_HD int64_t anchor_margin(int k) { return 4 * (int64_t)k + 32; }
// dynamic margin: the walker that starts at an anchor may read this far before it (k bases of look-back
// plus one 32-window cache refill), and the walker before it must have left that much untouched
JB_HD int64_t seam_read_margin(int k) { return (int64_t)k + 32; }

JB_HD bool bit_at(const uint32_t* words, int64_t p) { return (words[p >> 5] >> (p & 31)) & 1u; }

// ev / weak: bit words of this contig (bit 0 = position 0).  len = contig length.
JB_HD bool is_anchor(const uint32_t* ev, const uint32_t* weak, int64_t a, int k, int64_t len, int64_t margin) {
  if (a < margin + k || a < seam_read_margin(k) + k || a + k - 1 > len) return false;
  if (!bit_at(weak, a) || bit_at(weak, a - 1)) return false;
  for (int64_t p = a + 1; p <= a + k - 2; ++p) if (!bit_at(weak, p)) return false;
  for (int64_t p = a - margin; p < a; ++p) if (bit_at(ev, p)) return false;
  return true;
}

// First anchor inside [lo, hi) or -1 (scalar; the device kernel below spreads candidates over lanes).
JB_HD int64_t first_anchor(const uint32_t* ev, const uint32_t* weak, int64_t lo, int64_t hi, int k, int64_t len, int64_t margin) {
  for (int64_t a = lo; a < hi; ++a)
    if (is_anchor(ev, weak, a, k, len, margin)) return a;
  return -1;
}

#if defined(__CUDACC__)
struct SegUnit { uint64_t bit_off; uint64_t len; };

// one warp per block of seg_block positions: anchors[blk] = first anchor in the block or -1
__global__ void k_find_anchors(const uint32_t* __restrict__ events, const uint32_t* __restrict__ weak,
                               const SegUnit* __restrict__ units, const uint32_t* __restrict__ block_first, uint32_t n_units,
                               uint32_t n_blocks, int k, int seg_block, int margin, long long* __restrict__ anchors) {
  const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (blk >= n_blocks) return;
  // blocks are numbered contig by contig (make_blocks): contig u owns [block_first[u], block_first[u+1]),
  // its j-th block starts at (j + 1) * seg_block
  uint32_t ulo = 0, uhi = n_units;
  while (uhi - ulo > 1) { uint32_t mid = (ulo + uhi) >> 1; if (block_first[mid] <= blk) ulo = mid; else uhi = mid; }
  struct { uint64_t start; } b{(uint64_t)(blk - block_first[ulo] + 1) * (uint64_t)seg_block};
  const SegUnit u = units[ulo];
  const uint32_t* ev = events + (u.bit_off >> 5);
  const uint32_t* wk = weak + (u.bit_off >> 5);
  const int64_t len = (int64_t)u.len;
  const int64_t lo = (int64_t)b.start, hi = lo + seg_block < len ? lo + seg_block : len;
  long long best = -1;
  for (int64_t w0 = lo; w0 < hi && best < 0; w0 += 32 * 32) {
    // each lane takes one word of 32 positions: candidates are run starts (weak & ~weak<<1)
    const int64_t wp = w0 + (int64_t)lane * 32;
    long long mine = -1;
    if (wp < hi) {
      uint32_t cur = wk[wp >> 5];
      uint32_t prev_bit = wp > 0 ? (wk[(wp >> 5) - 1] >> 31) : 0u;
      uint32_t starts = cur & ~((cur << 1) | prev_bit);
      while (starts && mine < 0) {
        int bpos = __ffs(starts) - 1;
        starts &= starts - 1;
        int64_t a = wp + bpos;
        if (a < hi && is_anchor(ev, wk, a, k, len, margin)) mine = a;
      }
    }
    // lowest lane with a hit wins (positions increase with the lane index)
    unsigned int ball = __ballot_sync(0xffffffffu, mine >= 0);
    if (ball) best = __shfl_sync(0xffffffffu, mine, __ffs(ball) - 1);
  }
  if (lane == 0) anchors[blk] = best;
}
#endif

}  // namespace jb

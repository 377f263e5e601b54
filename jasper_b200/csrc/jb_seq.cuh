// jb_seq.cuh -- read access to the sequence a walker is editing, and window counts over it.
//
// The new sequence of a segment is   in[0:base) ++ closed copy jobs ++ out[wstart:olen) ++ in[ipos:)
// (jb_walker.cuh).  The walker publishes that view in its shared-memory scratch whenever an edit changes
// it; these helpers read it there and stay out of line on the device, so the
// byte-gather / key / hash / probe code exists once in the kernel image instead of once per call site
// (the walker calls C(window) from ~20 places), and the walker's own state stays in registers.
#pragma once
#include <stdint.h>

#include "jb_cand.cuh"

namespace jb {

enum : uint32_t { kFlagSeam = 1u, kFlagFar = 2u, kFlagProbed = 4u, kFlagTooWide = 8u };

JB_FREE_INLINE uint8_t view_get(const SeqView& v, int64_t x, uint32_t& flags) {
  if (x < v.base) {                                   // bases the previous segment's walker owns
    if (x < v.base - v.seam_margin) flags |= kFlagSeam;   // ...and may have edited: the contig is re-walked whole
    return v.in[x];
  }
  const int64_t y = x - v.base;
  if (y >= v.olen) return v.in[v.ipos + (y - v.olen)];
  if (y >= v.wstart) return v.out[y];
  flags |= kFlagFar;                                  // behind the window: the contig is re-walked fully materialised
  return (uint8_t)'A';
}

JB_FREE_INLINE uint32_t warp_or(uint32_t flags) {     // lanes read different bases: make the verdict warp-uniform
#if defined(__CUDA_ARCH__)
  return (__any_sync(0xffffffffu, flags & kFlagSeam) ? kFlagSeam : 0u) | (__any_sync(0xffffffffu, flags & kFlagFar) ? kFlagFar : 0u) |
         (flags & ~(kFlagSeam | kFlagFar));
#else
  return flags;
#endif
}

// dst[0:n) = seq[a:b) clamped to [0, L), dst[n] = 0; returns n | flags << 16
JB_FREE_NOINLINE uint32_t view_load(const SeqView* vp, uint8_t* dst, int64_t a, int64_t b) {
  const SeqView v = *vp;
  if (a < 0) a = 0;
  if (b > v.L) b = v.L;
  const int n = b > a ? (int)(b - a) : 0;
  uint32_t flags = 0;
  JB_SYNC();
  JB_LANES(l) { for (int q = l; q < n; q += 32) dst[q] = view_get(v, a + q, flags); if (l == 0) dst[n] = 0; }
  JB_SYNC();
  return (uint32_t)n | (warp_or(flags) << 16);
}

// s.wc[l] = C(seq[b0+l : b0+l+k]) for l in [0, 32); returns flags
template <class DB>
JB_FREE_NOINLINE uint32_t view_refill(const DB& db, WarpScratch& s, int64_t b0, int k, Key mask) {
  const SeqView v = s.view;
  uint32_t flags = 0;
  if (v.cnt && b0 >= v.clean_from) {
    // untouched territory: the bulk scan of this pass already counted these windows
    const int64_t p0 = b0 - (v.base + v.olen - v.ipos);        // pass-start coordinate of window b0
    if (p0 < v.base - v.seam_margin) flags |= kFlagSeam;        // (the previous segment's walker may have changed them)
    JB_SYNC();
    JB_LANES(l) { const int64_t p = p0 + l; s.wc[l] = v.cnt[p < v.in_len ? p : v.in_len]; }
    JB_SYNC();
    return flags;
  }
  const uint32_t r = view_load(&s.view, s.c, b0, b0 + 32 + k - 1);
  const int len = (int)(r & 0xFFFFu);
  pack_bytes(s.pc, s.c, len);
  JB_LANES(l) { s.wc[l] = count_packed(db, s.pc, s.c, len, l < len ? l : len, k, mask); }
  JB_SYNC();
  return (r >> 16) | kFlagProbed;
}

// C(seq[x : x+k]) through the 32-window cache (a miss refills it in direction dir); returns count | flags << 32
template <class DB>
JB_FREE_NOINLINE uint64_t view_count(const DB& db, WarpScratch& s, int64_t x, int dir, int k, Key mask) {
  if (x < 0) return s.empty_count;                   // the slices with negative starts are empty (L >= k)
  if (x >= s.wc_base && x < s.wc_base + s.wc_n) return s.wc[x - s.wc_base];
  const int64_t b0 = dir > 0 ? x : (x - 31 > 0 ? x - 31 : 0);
  const uint32_t f = view_refill(db, s, b0, k, mask);
  JB_SYNC();
  JB_LANES(l) { if (l == 0) { s.wc_base = b0; s.wc_n = 32; if (f & kFlagProbed) s.probes += 32; } }
  JB_SYNC();
  return (uint64_t)s.wc[x - b0] | ((uint64_t)(f & (kFlagSeam | kFlagFar)) << 32);
}

// s.pr[q] = C(seq[s.px[q] : +k]) for q < n <= 32 (starts non-negative); returns flags (kFlagTooWide: nothing done)
template <class DB>
JB_FREE_NOINLINE uint32_t view_points(const DB& db, WarpScratch& s, int n, int k, Key mask) {
  const SeqView v = s.view;
  int64_t lo = s.px[0], hi = s.px[0];
  for (int q = 1; q < n; ++q) { lo = s.px[q] < lo ? s.px[q] : lo; hi = s.px[q] > hi ? s.px[q] : hi; }
  uint32_t flags = 0;
  if (v.cnt && lo >= v.clean_from) {                 // untouched territory: read the bulk scan's counts
    const int64_t shift = v.base + v.olen - v.ipos;
    if (lo - shift < v.base - v.seam_margin) flags |= kFlagSeam;
    JB_SYNC();
    JB_LANES(l) { if (l < n) { const int64_t p = s.px[l] - shift; s.pr[l] = v.cnt[p < v.in_len ? p : v.in_len]; } }
    JB_SYNC();
    return flags;
  }
  if (hi + k - lo > kStrCap - 8) return kFlagTooWide;
  const uint32_t r = view_load(&s.view, s.c, lo, hi + k);
  const int len = (int)(r & 0xFFFFu);
  pack_bytes(s.pc, s.c, len);
  JB_LANES(l) { if (l < n) { const int m = (int)(s.px[l] - lo); s.pr[l] = count_packed(db, s.pc, s.c, len, m < len ? m : len, k, mask); } }
  JB_SYNC();
  return (r >> 16) | kFlagProbed;
}

}  // namespace jb

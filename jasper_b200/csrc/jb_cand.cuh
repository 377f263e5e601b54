// jb_cand.cuh -- candidate-edit scoring, one probe wave for many candidates.
//
// The reference tries its candidate fixes one after another and accepts the first whose sampled k-mers
// are all solid (src/jasper.py:392-431 fix_k_case_sub / fix_insert / fix_del, :340-388 fixdiploid, the
// single-base tails of :434-524; the test itself is check_sequence, :585-599).  Here the candidates of
// one call are DESCRIBED (kind, position, base) instead of materialised, every (candidate, sample)
// pair becomes one lane's probe, the pairs are laid out candidate-major over consecutive waves, and the
// first candidate (in the reference's enumeration order) without a failing sample wins -- the same
// answer as the sequential loop, in ceil(pairs / 32) dependent probe waves instead of one per candidate.
//
// The functions are free templates (not Walker members) and stay out of line on the device: they take
// only shared-memory scratch and scalars, so the walker's own state stays in registers and the heavy
// key/hash/probe code exists once in the kernel image (the all-inlined walker was 860 KB of SASS and
// stalled on instruction fetch).
#pragma once
#include <stdint.h>

#include "jb_kmer.cuh"

#if defined(__CUDACC__)
#define JB_FREE_NOINLINE __device__ __noinline__
#define JB_FREE_INLINE __device__ __forceinline__
#else
#define JB_FREE_NOINLINE
#define JB_FREE_INLINE inline
#endif
#if defined(__CUDA_ARCH__)
#define JB_LANES(l) for (int l = (int)(threadIdx.x & 31), once_ = 1; once_; once_ = 0)
#define JB_SYNC() __syncwarp()
#else
#define JB_LANES(l) for (int l = 0; l < 32; ++l)
#define JB_SYNC() ((void)0)
#endif

namespace jb {

constexpr int kStrCap = 4 * JB_MAX_K + 64;       // every tbf / trial / check string fits (<= 3k + 8)
constexpr int kMaxCands = 64;

enum CandKind : uint8_t { kSub = 0, kIns = 1, kDel = 2, kDip = 3 };

// An edit of the to_be_fixed string t (n chars):
//   kSub  t[pos] = base                     kIns  t[:pos] + base * m + t[pos:]
//   kDel  t[:pos] + t[pos+m:]               kDip  before + t{[n-k] = base, [k-1] = base2} + after   (fixdiploid)
struct Cand { uint8_t kind, base, base2, pad; int16_t pos, m; };

// The sequence a walker is editing (jb_seq.cuh reads it, jb_walker.cuh publishes it after every edit).
struct SeqView {
  const uint8_t* in;         // the contig at the start of the pass
  const uint8_t* out;        // staging: out[r] = new base r of this segment inside the materialised window
  const uint32_t* cnt;       // bulk-scan counts of the pass-start contig, [0, in_len]; may be null
  int64_t base, olen, ipos, wstart, L, in_len, clean_from, seam_margin;
};

struct WarpScratch {         // shared memory (device) / plain struct (host)
  uint8_t t[kStrCap];        // to_be_fixed
  uint8_t b[kStrCap];        // replacement text / good_k_mer_after
  uint8_t c[kStrCap];        // check string / window staging / diploid flanks
  uint32_t wc[32];           // window-count cache
  int64_t px[32];            // point queries
  uint32_t pr[32];
  uint32_t flag[32];
  uint32_t ext[32];          // extension verdicts of one bridging wave
  Cand cand[kMaxCands];
  uint16_t cand_first[kMaxCands + 1];   // first work item of every candidate
  uint32_t failmask[2];
  int32_t dip_nb, dip_bl;    // kDip: chars of flank before (s.c[0:nb)) and after (s.c[nb+n : nb+n+bl))
  uint8_t fx[128];           // text of the fix row being prepared: mutation at [0:64), original at [64:128)
  SeqView view;              // published copy of the walker's sequence state
  int64_t wc_base;           // window-count cache: wc[0:wc_n) = C(seq[wc_base + l : ...])
  int32_t wc_n;
  uint32_t empty_count;      // C('') = count of poly-A
  unsigned long long probes; // table probes issued by the out-of-line helpers
};

// C(buf[m : m+k]) with MerDNA's padding (JF!include/jellyfish/mer_dna.hpp:525-542, 642-644)
template <class DB>
JB_FREE_NOINLINE uint32_t count_str(const DB& db, const uint8_t* buf, int len, int m, int k, Key mask) {
  Key key = key_from_chars(buf + m, (int64_t)len - m, k, mask);
  return db.count(key_canonical(key, k));
}

JB_FREE_INLINE int cand_len(const WarpScratch& s, const Cand& c, int n) {
  if (c.kind == kIns) return n + c.m;
  if (c.kind == kDel) return n - c.m;
  if (c.kind == kDip) return s.dip_nb + n + s.dip_bl;
  return n;
}

JB_FREE_INLINE uint8_t cand_char(const WarpScratch& s, const Cand& c, int n, int k, int q) {
  if (c.kind == kSub) return q == c.pos ? c.base : s.t[q];
  if (c.kind == kIns) return q < c.pos ? s.t[q] : (q < c.pos + c.m ? c.base : s.t[q - c.m]);
  if (c.kind == kDel) return q < c.pos ? s.t[q] : s.t[q + c.m];
  const int r = q - s.dip_nb;                                  // kDip
  if (r < 0 || r >= n) return s.c[q];
  return r == n - k ? c.base : (r == k - 1 ? c.base2 : s.t[r]);
}

// count of the window starting at m of the candidate's check string (len chars)
template <class DB>
JB_FREE_INLINE uint32_t cand_count(const DB& db, const WarpScratch& s, const Cand& c, int n, int len, int m, int k, Key mask) {
  Key key = {0, 0};
  int j = 0;
  for (; j < k && m + j < len; ++j) {
    const uint8_t ch = cand_char(s, c, n, k, m + j);
    if (!base_valid(ch)) break;
    key = key_push(key, base_code(ch), mask);
  }
  key = key_shl_bases(key, k - j, mask);
  return db.count(key_canonical(key, k));
}

// number of check_sequence samples of a string of len chars: 0, max(0, len-k), then step, 2 step, ... < len-k
JB_FREE_INLINE int check_samples(int len, int k, int step) { return 2 + (len - k > 0 ? (len - k - 1) / step : 0); }
JB_FREE_INLINE int check_sample_at(int q, int len, int k, int step) { return q == 0 ? 0 : (q == 1 ? (len - k > 0 ? len - k : 0) : step * (q - 1)); }

// First candidate of s.cand[0:nc) (over s.t, n chars) that check_sequence accepts, or -1.
template <class DB>
JB_FREE_NOINLINE int first_passing(const DB& db, WarpScratch& s, int nc, int n, int k, int step, Key mask, int64_t thr,
                                   unsigned long long* lookups) {
  JB_SYNC();
  JB_LANES(l) {
    if (l == 0) {
      int run = 0;
      for (int c = 0; c < nc; ++c) { s.cand_first[c] = (uint16_t)run; run += check_samples(cand_len(s, s.cand[c], n), k, step); }
      s.cand_first[nc] = (uint16_t)run;
      s.failmask[0] = 0; s.failmask[1] = 0;
    }
  }
  JB_SYNC();
  const int total = s.cand_first[nc];
  int decided = 0;
  for (int w0 = 0; w0 < total; w0 += 32) {
    JB_LANES(l) {
      const int item = w0 + l;
      if (item < total) {
        int c = decided;
        while (s.cand_first[c + 1] <= item) ++c;
        const Cand cd = s.cand[c];
        const int len = cand_len(s, cd, n);
        const int m = check_sample_at(item - s.cand_first[c], len, k, step);
        if ((int64_t)cand_count(db, s, cd, n, len, m, k, mask) < thr) {
#if defined(__CUDA_ARCH__)
          atomicOr(&s.failmask[c >> 5], 1u << (c & 31));
#else
          s.failmask[c >> 5] |= 1u << (c & 31);
#endif
        }
      }
    }
    JB_SYNC();
    *lookups += (unsigned long long)(total - w0 < 32 ? total - w0 : 32);
    while (decided < nc && (int)s.cand_first[decided + 1] <= w0 + 32) {
      if (!((s.failmask[decided >> 5] >> (decided & 31)) & 1u)) return decided;
      ++decided;
    }
  }
  return -1;
}

// number of windows (all of them) of candidate s.cand[ci]'s string below thr  (the homopolymer repairs count every window,
// src/jasper.py:452-459, 497-505)
template <class DB>
JB_FREE_NOINLINE int count_weak_cand(const DB& db, WarpScratch& s, int ci, int n, int k, Key mask, int64_t thr, unsigned long long* lookups) {
  JB_SYNC();
  const Cand cd = s.cand[ci];
  const int len = cand_len(s, cd, n);
  const int total = len - k + 1;
  int bad = 0;
  for (int w0 = 0; w0 < total; w0 += 32) {
    JB_SYNC();
    JB_LANES(l) {
      const int q = w0 + l;
      s.flag[l] = (q < total && (int64_t)cand_count(db, s, cd, n, len, q, k, mask) < thr) ? 1u : 0u;
    }
    JB_SYNC();
    const int cnt = total - w0 < 32 ? total - w0 : 32;
    *lookups += (unsigned long long)cnt;
    for (int q = 0; q < cnt; ++q) bad += (int)s.flag[q];
  }
  return bad;
}

}  // namespace jb

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
constexpr int kSmallPaths = 32;                  // base_extension: paths kept in the warp's shared memory
constexpr int kFastBack = 64;                    // make_verdict: staged windows in front of the event ...
constexpr int kFastWin = kFastBack + 2 * JB_MAX_K + 24;   // ... and in all

enum CandKind : uint8_t { kSub = 0, kIns = 1, kDel = 2, kDip = 3 };

// An edit of the to_be_fixed string t (n chars):
//   kSub  t[pos] = base                     kIns  t[:pos] + base * m + t[pos:]
//   kDel  t[:pos] + t[pos+m:]               kDip  before + t{[n-k] = base, [k-1] = base2} + after   (fixdiploid)
struct Cand { uint8_t kind, base, base2, pad; int16_t pos, m; };

// ---- 2-bit packed strings ---------------------------------------------------------------------
// Building a k-mer from bytes costs ~20 instructions per base and every probe lane used to do it for its
// own window (45 % of the walker's instructions, ncu r1n).  A string that many lanes cut windows from is
// packed once by the whole warp: base q sits in word q/32 at bits 62-2(q%32) (first base on top), invalid
// bytes are recorded in a bitmask, and a window is three shifts.
constexpr int kPackBases = 256;
struct Packed {
  uint64_t w[kPackBases / 32 + 2];   // two zero words behind the data: windows may read past the end
  uint32_t inv[kPackBases / 32];     // bit l of inv[c]: byte 32c + l is not ACGTacgt (bytes past len count as invalid)
  int32_t len;
};

// bases [q0, q0 + n) of p as a 2n-bit integer, n <= 64, q0 + n <= kPackBases
JB_HD Key packed_window(const Packed& p, int q0, int n) {
  const int bit = 2 * q0, j = bit >> 6, r = bit & 63;
  const uint64_t a = p.w[j], b = p.w[j + 1], c = p.w[j + 2];
  const uint64_t X = r ? (a << r) | (b >> (64 - r)) : a;        // the 128 bits that start at base q0
  const uint64_t Y = r ? (b << r) | (c >> (64 - r)) : b;
  Key key;
  if (n <= 32) { key.hi = 0; key.lo = n ? X >> (64 - 2 * n) : 0; }
  else { const int s = 128 - 2 * n; key.lo = s ? (Y >> s) | (X << (64 - s)) : Y; key.hi = s ? X >> s : X; }
  return key;
}

// no invalid byte in [q0, q0 + n) and the range lies inside the string
JB_HD bool packed_clean(const Packed& p, int q0, int n) {
  if (n <= 0) return true;
  if (q0 < 0 || q0 + n > p.len) return false;
  const int c0 = q0 >> 5, c1 = (q0 + n - 1) >> 5;
  for (int c = c0; c <= c1; ++c) {
    uint32_t m = p.inv[c];
    if (c == c0) m &= 0xFFFFFFFFu << (q0 & 31);
    if (c == c1) { const int top = (q0 + n - 1) & 31; if (top < 31) m &= (2u << top) - 1u; }
    if (m) return false;
  }
  return true;
}

// (a << 2n) | b for keys of bases
JB_HD Key key_append(Key a, Key b, int n) {
  if (n <= 0) return a;
  Key r;
  const int s = 2 * n;
  if (s >= 128) return b;
  if (s >= 64) { r.hi = (a.lo << (s - 64)) | b.hi; r.lo = b.lo; if (s == 64) r.hi = a.lo | b.hi; }
  else { r.hi = (a.hi << s) | (a.lo >> (64 - s)) | b.hi; r.lo = (a.lo << s) | b.lo; }
  return r;
}

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
  Key pkm[2][kSmallPaths];   // base_extension: the path list while it is short (longer ones live in the bridging slot)
  uint32_t pnode[2][kSmallPaths];
  Cand cand[kMaxCands];
  uint16_t cand_first[kMaxCands + 1];   // first work item of every candidate
  uint32_t failmask[2];
  int32_t dip_nb, dip_bl;    // kDip: chars of flank before (s.c[0:nb)) and after (s.c[nb+n : nb+n+bl))
  uint8_t fx[128];           // text of the fix row being prepared: mutation at [0:64), original at [64:128)
  uint32_t fc[kFastWin];     // make_verdict: bulk-scan counts of the windows around an event (jb_verdict.cuh)
  int32_t fbad[2];           // make_verdict: last unavailable entry before / first one behind the event
  uint32_t hc[64];           // make_verdict: counts of every window of a homopolymer trial string
  uint32_t win[2];           // score_single_base: counts of the winner's samples 0 (m = 0) and 1 (m = len - k)
  SeqView view;              // published copy of the walker's sequence state
  Packed pt;                 // packed to_be_fixed (or, for kDip, the whole check string before + t + after)
  Packed pc;                 // packed staging buffer (window refills, point queries, bridging checks)
  int64_t wc_base;           // window-count cache: wc[0:wc_n) = C(seq[wc_base + l : ...])
  int32_t wc_n;
  uint32_t empty_count;      // C('') = count of poly-A
  unsigned long long probes; // table probes issued by the out-of-line helpers
};

// pack buf[0:len) (len <= kPackBases) with the whole warp
static JB_FREE_NOINLINE void pack_bytes(Packed& p, const uint8_t* buf, int len) {
  if (len > kPackBases) len = kPackBases;
  const int chunks = (len + 31) >> 5;
  JB_SYNC();
#if defined(__CUDA_ARCH__)
  const int l = threadIdx.x & 31;
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    const int q = 32 * c + l;
    const uint8_t ch = q < len ? buf[q] : (uint8_t)0;
    const bool v = q < len && base_valid(ch);
    const uint32_t code = v ? base_code(ch) : 0u;
    const uint32_t hi = __reduce_or_sync(0xffffffffu, l < 16 ? code << (30 - 2 * l) : 0u);
    const uint32_t lo = __reduce_or_sync(0xffffffffu, l >= 16 ? code << (62 - 2 * l) : 0u);
    const uint32_t inv = __ballot_sync(0xffffffffu, !v);
    if (l == 0) { p.w[c] = ((uint64_t)hi << 32) | lo; p.inv[c] = inv; }
  }
  if (l == 0) { p.w[chunks] = 0; p.w[chunks + 1] = 0; p.len = len; }
#else
  for (int c = 0; c < chunks; ++c) {
    uint64_t w = 0; uint32_t inv = 0;
    for (int l = 0; l < 32; ++l) {
      const int q = 32 * c + l;
      const bool v = q < len && base_valid(buf[q]);
      if (v) w |= (uint64_t)base_code(buf[q]) << (62 - 2 * l); else inv |= 1u << l;
    }
    p.w[c] = w; p.inv[c] = inv;
  }
  p.w[chunks] = 0; p.w[chunks + 1] = 0; p.len = len;
#endif
  JB_SYNC();
}

// C(window m of a packed string of len bytes), falling back to the byte loop near invalid bytes / the end
template <class DB>
JB_FREE_INLINE uint32_t count_packed(const DB& db, const Packed& p, const uint8_t* buf, int len, int m, int k, Key mask) {
  Key key;
  if (packed_clean(p, m, k)) key = packed_window(p, m, k);
  else key = key_from_chars(buf + m, (int64_t)len - m, k, mask);
  return db.count(key_canonical(key, k));
}

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

// bases [q0, q0 + cnt) of the packed to_be_fixed with base `code` substituted at position pos (if inside)
JB_FREE_INLINE Key sub_base(Key key, int q0, int cnt, int pos, uint32_t code) {
  if (pos < q0 || pos >= q0 + cnt) return key;
  const int sh = 2 * (q0 + cnt - 1 - pos);
  if (sh >= 64) key.hi = (key.hi & ~(3ull << (sh - 64))) | ((uint64_t)code << (sh - 64));
  else key.lo = (key.lo & ~(3ull << sh)) | ((uint64_t)code << sh);
  return key;
}

// fast path of cand_count: the window cut from the packed string s.pt; false = use the byte loop
JB_FREE_INLINE bool cand_key_packed(const WarpScratch& s, const Cand& c, int n, int len, int m, int k, Key* out) {
  if (m < 0 || m + k > len) return false;
  const Packed& p = s.pt;
  if (c.kind == kSub) {
    if (!packed_clean(p, m, k)) return false;
    *out = sub_base(packed_window(p, m, k), m, k, c.pos, base_code(c.base));
    return true;
  }
  if (c.kind == kDip) {                                // s.pt holds before + t + after
    if (!packed_clean(p, m, k)) return false;
    Key key = packed_window(p, m, k);
    key = sub_base(key, m, k, s.dip_nb + k - 1, base_code(c.base2));
    key = sub_base(key, m, k, s.dip_nb + n - k, base_code(c.base));     // the left base wins when both hit one position
    *out = key;
    return true;
  }
  // kIns / kDel: a stretch of t, then (kIns) copies of the inserted base, then another stretch of t
  const int P = c.pos, mi = c.m;
  const int lenA = P > m ? (P - m < k ? P - m : k) : 0;                  // trial positions [m, min(P, m + k))
  int cB = 0, cstart;                                                   // t index where the last stretch starts
  if (c.kind == kIns) {
    const int b0 = m > P ? m : P, b1 = (P + mi < m + k) ? P + mi : m + k;
    cB = b1 > b0 ? b1 - b0 : 0;
    const int q = m > P + mi ? m : P + mi;                              // first trial position behind the insertion
    cstart = q - mi;
  } else {
    const int q = m > P ? m : P;
    cstart = q + mi;
  }
  const int lenC = k - lenA - cB;
  if (lenC < 0) return false;
  if (!packed_clean(p, m, lenA) || !packed_clean(p, cstart, lenC)) return false;
  Key key = lenA ? packed_window(p, m, lenA) : Key{0, 0};
  if (cB) {
    const uint64_t rep = 0x5555555555555555ull * base_code(c.base);     // the base repeated
    Key b;
    if (cB <= 32) { b.hi = 0; b.lo = cB == 32 ? rep : (rep & ((1ull << (2 * cB)) - 1)); }
    else { b.lo = rep; b.hi = rep & ((1ull << (2 * (cB - 32))) - 1); }
    key = key_append(key, b, cB);
  }
  if (lenC) key = key_append(key, packed_window(p, cstart, lenC), lenC);
  *out = key;
  return true;
}

// count of the window starting at m of the candidate's check string (len chars)
template <class DB>
JB_FREE_INLINE uint32_t cand_count(const DB& db, const WarpScratch& s, const Cand& c, int n, int len, int m, int k, Key mask) {
  Key fast;
  if (cand_key_packed(s, c, n, len, m, k, &fast)) return db.count(key_canonical(fast, k));
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

// Single-base candidates of an isolated error, scored from the packed to_be_fixed string s.pt (n valid bases, the suspect
// base at k-1), in the reference's enumeration order:
//   ins == 0 (src/jasper.py:392-419): substitutions at k-1 by 'A','C','T','G' except the base that stands there
//                                     (skip_code), then the deletion of k-1;
//   ins == 1 (src/jasper.py:422-431): one base 'A','T','C','G' inserted in front of k-1.
// Every candidate owns S lanes (one per check_sequence sample, src/jasper.py:585-599), a wave holds 32 / S candidates,
// a lane builds its window as  t[m : k-1] ++ (the new base) ++ t[cstart : ...]  with shifts.  Returns the first
// candidate whose samples are all solid (0..3), -1 for none, -2 if the layout does not apply (nothing was probed).
// s.win = counts of the winner's samples 0 (window at 0) and 1 (window at len - k).  Compact on purpose: this and the
// probe are the walker's hot loop, which has to stay inside the SM's instruction cache (tools/icache_probe.cu).
template <class DB>
JB_FREE_NOINLINE int score_single_base(const DB& db, WarpScratch& s, int ins, uint32_t skip_code, int n, int k, int step, int64_t thr) {
  const int P = k - 1;
  const int len0 = ins ? n + 1 : n;                                  // trial length of candidates 0..2 (and of 3 when ins)
  const int ns0 = check_samples(len0, k, step);
  const int ns3 = ins ? ns0 : check_samples(n - 1, k, step);
  const int S = ns0 > ns3 ? ns0 : ns3;
  if (S > 16 || n - 1 < k) return -2;
  int G = 1;
  while ((G + 1) * S <= 32 && G < 4) ++G;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int c0 = 0; c0 < 4; c0 += G) {
    JB_SYNC();
    JB_LANES(l) {
      int ci = c0, q = l;
      while (q >= S && ci < c0 + G - 1) { q -= S; ++ci; }
      const bool del = !ins && ci == 3;
      const int len = del ? n - 1 : len0;
      uint32_t fail = 0, cnt = 0;
      if (ci < 4 && q < (ci == 3 ? ns3 : ns0)) {
        uint32_t code;                                               // the new base
        if (ins) code = ci == 0 ? 0u : (ci == 1 ? 3u : (ci == 2 ? 1u : 2u));          // 'A','T','C','G'
        else { const uint32_t o0 = ci + (ci >= (skip_code == 0u ? 0 : (skip_code == 1u ? 1 : (skip_code == 3u ? 2 : 3))) ? 1u : 0u);
               code = o0 == 0 ? 0u : (o0 == 1 ? 1u : (o0 == 2 ? 3u : 2u)); }         // 'A','C','T','G' without skip_code
        const int m = check_sample_at(q, len, k, step);
        const int lenA = P - m, mid = del ? 0 : 1, cstart = ins ? P : P + 1, lenC = k - lenA - mid;
        Key key = packed_window(s.pt, m, lenA);
        if (mid) { const Key b = {code, 0}; key = key_append(key, b, 1); }
        key = key_append(key, packed_window(s.pt, cstart, lenC), lenC);
        cnt = db.count(key_canonical(key, k));
        fail = (int64_t)cnt < thr ? 1u : 0u;
      }
      s.flag[l] = fail; s.pr[l] = cnt;
    }
    JB_SYNC();
    const int c1 = c0 + G < 4 ? c0 + G : 4;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int ci = c0; ci < c1; ++ci) {
      uint32_t f = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int q = 0; q < S; ++q) f |= s.flag[(ci - c0) * S + q];
      if (!f) {
        const uint32_t w0 = s.pr[(ci - c0) * S], w1 = s.pr[(ci - c0) * S + 1];
        JB_SYNC();
        JB_LANES(l) { if (l == 0) { s.win[0] = w0; s.win[1] = w1; } }
        JB_SYNC();
        return ci;
      }
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

// jb_verdict.cuh -- verdicts: the isolated error, decided in bulk before the walkers run.
//
// An isolated substitution / indel leaves one run of weak windows in untouched sequence, and everything the
// reference does with it -- delimiting the run (src/jasper.py:150-214), trying the single-base candidates
// (:392-431 through check_sequence :585-599) and the scan step that follows the fix (:56-100) -- reads nothing but
// pass-start bases and their counts.  k_verdicts therefore evaluates every weak run of the pass-start contigs
// once, one warp per run, in a kernel whose whole code fits the SM's instruction cache, and stores a 16-byte
// verdict; a walker that reaches the run while it still stands in untouched territory consumes the verdict
// (jb_walker.cuh consume_verdict) instead of executing that code itself.  The walker kernel is bound by
// instruction fetch (tools/icache_probe.cu: throughput drops 6x once the working set of divergent warps leaves
// the ~32 KB instruction cache), so moving the common case out of it is what makes it fast; exactness is
// unchanged: a verdict is only used when every position it read is still clean for that walker, and all other
// events run the generic code.
//
// The functions compile for the host too (tests/hostsim builds verdicts the same way).
#pragma once
#include <stdint.h>

#include "jb_cand.cuh"

namespace jb {

enum VerdictKind : uint8_t { kVDelimit = 0, kVSub = 1, kVDel = 2, kVIns = 3, kVNone = 4, kVHomoIns = 5 };

struct Verdict {
  uint32_t re;          // low 32 bits of the contig position of the first window behind the run (= the walker's i after delimiting)
  uint16_t d_jq;        // re - jq, jq = last solid window in front of the run
  uint16_t d_s0;        // s0 - jq: start of to_be_fixed
  uint16_t d_ga;        // ga - jq: good_after (after the src/jasper.py:204-214 adjustment)
  uint8_t kind;         // VerdictKind: kVDelimit = bounds only (the walker runs the reference's dispatch itself)
  uint8_t base;         // kVSub / kVIns: the new base
  uint8_t bad;          // the base at s0 + k - 1
  uint8_t post_ok;      // bit 0: the scan step behind the fix is the plain `i += k-1`; bit 1: so is the one after it
  uint16_t pad;
};
static_assert(sizeof(Verdict) == 16, "Verdict is meant to be 16 bytes");
constexpr uint32_t kVmapAmbiguous = 0xFFFFFFFFu;

// Second-stage verdict of a run whose first verdict is kVDelimit: what the reference's dispatch (src/jasper.py:226-338:
// homopolymer repairs, diploid, lower-case substitutions, and the de-Bruijn bridging of multi-error stretches) decides
// on the pass-start sequence.  k_verdicts2 runs the walker's own code for it (Walker::dry_fix), one warp per run, so
// that this work is spread over the whole chip instead of sitting in the walkers' sequential chains.  Patches longer
// than kV2Repl bases stay with the walkers.
constexpr int kV2Repl = 108;
struct Verdict2 {
  uint8_t valid;        // 0: the walker runs slow_fix itself
  uint8_t found;
  int16_t a_rel, e_rel; // seq = seq[:s0 + a_rel] + repl + seq[s0 + e_rel:]
  int16_t rlen;
  int16_t coord_rel;    // fix row: coordinate - s0
  uint8_t fb, fb_len, tag, orig, orig_len;    // mutation = fb * fb_len, original = tag + orig * orig_len
  uint8_t post_ok;      // the scan step behind the applied fix is the plain `i += k-1` (evaluated on the edited sequence)
  uint8_t pad[4];
  uint8_t repl[kV2Repl];
};
static_assert(sizeof(Verdict2) == 128, "Verdict2 is meant to be 128 bytes");      // two runs end inside the same 32 positions: no verdict for either

// the last staged window q in (blo, from] whose count is solid, or blo: the reference walks back one window at a time
// (src/jasper.py:152-158), a whole weak run of k-1 windows per event; the warp looks at 32 of them per step
JB_FREE_INLINE int last_solid_back(const WarpScratch& s, int from, int blo, int64_t solid) {
#if defined(__CUDA_ARCH__)
  const int lane = (int)(threadIdx.x & 31);
  for (int top = from; top > blo; top -= 32) {
    const int q = top - lane;
    const unsigned m = __ballot_sync(0xffffffffu, q > blo && (int64_t)s.fc[q] >= solid);
    if (m) return top - (__ffs((int)m) - 1);
  }
  return blo;
#else
  int jq = from;
  while (jq > blo && (int64_t)s.fc[jq] < solid) --jq;
  return jq;
#endif
}

// Verdict of the weak run that contains position a_pos (pass-start coordinates; the window at a_pos is valid and weak).
// False: no verdict (the run touches invalid windows, a contig end, or is longer than the staged window).
template <class DB>
JB_FREE_NOINLINE bool make_verdict(const DB& db, WarpScratch& s, const uint8_t* in, const uint32_t* cnt, const uint8_t* okb,
                                   int64_t in_len, int k, int step, int64_t solid, int64_t a_pos, Verdict* out, int64_t* re_out, bool second_step = true) {
  const int W = kFastBack + 2 * k + 24;
  const int64_t W0 = a_pos - kFastBack;
  const int64_t hi_p = in_len - 2 * (int64_t)k - 3;            // keeps every end-of-contig clamp of the reference out of play
  JB_SYNC();
  JB_LANES(l) { if (l == 0) { s.fbad[0] = -1; s.fbad[1] = W; } }
  JB_SYNC();
  JB_LANES(l) {
    // every load of the staged window first (validity and count side by side, all rounds in flight: r2aa had a third of the
    // kernel's stall samples on a validity load followed by a count load, round after round), then the bookkeeping
    constexpr int kRounds = (kFastWin + 31) / 32;
    uint8_t okv[kRounds];
    uint32_t cv[kRounds];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < kRounds; ++r) {
      const int q = l + 32 * r;
      const int64_t p = W0 + q;
      const bool in = q < W && p >= 0 && p <= hi_p;
      okv[r] = in ? okb[p] : (uint8_t)0;
      cv[r] = in ? cnt[p] : 0u;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < kRounds; ++r) {
      const int q = l + 32 * r;
      if (q >= W) continue;
      const bool ok = okv[r] != 0;
      s.fc[q] = ok ? cv[r] : 0u;
      if (!ok) {
#if defined(__CUDA_ARCH__)
        if (q < kFastBack) atomicMax(&s.fbad[0], q); else atomicMin(&s.fbad[1], q);
#else
        if (q < kFastBack) { if (q > s.fbad[0]) s.fbad[0] = q; } else if (q < s.fbad[1]) s.fbad[1] = q;
#endif
      }
    }
  }
  JB_SYNC();
  const int blo = s.fbad[0];                          // staged entries (blo, bhi) are all available
  int hi_lim = s.fbad[1] - k - 3;
  if (W - k - 4 < hi_lim) hi_lim = W - k - 4;
  const int a = kFastBack;
  if (a >= hi_lim || (int64_t)s.fc[a] >= solid) return false;
  // src/jasper.py:152-166: back to the last solid window, forward to the next one
  const int jq = last_solid_back(s, a - 1, blo, solid);
  if (jq <= blo) return false;
  int gq = a;
  while (gq < hi_lim && (int64_t)s.fc[gq] < solid) ++gq;
  if (gq >= hi_lim) return false;
  // src/jasper.py:176-196: the half-of-previous refinement of good_before (g = good_before - k + 1)
  uint64_t prev = s.fc[jq];
  int g = jq;
  if (!(2 * (int64_t)s.fc[g + 1] < solid && 2 * (int64_t)s.fc[g + 2] < solid)) {
    while (2 * (uint64_t)s.fc[g + 1] >= prev && g < gq) {
      if (2 * (int64_t)prev >= solid && 2 * (int64_t)s.fc[g + 1] < solid && 2 * (int64_t)s.fc[g + 2] < solid) break;
      prev = s.fc[g + 1];
      ++g;
    }
  }
  const int s0q = g + 1;
  int gaq = gq;
  if ((int64_t)s.fc[s0q + 1] < solid && (int64_t)s.fc[s0q + k - 2] < solid && (int64_t)s.fc[s0q + k - 1] < solid &&
      (int64_t)s.fc[s0q + k] >= solid) gaq = s0q + k;            // src/jasper.py:204-214
  const int nbad = gaq - s0q > 0 ? gaq - s0q : 0;
  Verdict v;
  v.re = (uint32_t)(W0 + gq); v.d_jq = (uint16_t)(gq - jq); v.d_s0 = (uint16_t)(s0q - jq); v.d_ga = (uint16_t)(gaq - jq);
  v.kind = kVDelimit; v.base = 0; v.bad = 0; v.post_ok = 0; v.pad = 0;
  if (nbad == k || nbad == k - 1) {
    // src/jasper.py:392-431: the single-base candidates of to_be_fixed = seq[s0 : ga + k - 1]
    const int n = nbad + k - 1;
    const int ins = nbad == k - 1 ? 1 : 0;
    const uint8_t* src = in + (W0 + s0q);
    JB_SYNC();
    JB_LANES(l) { for (int q = l; q < n; q += 32) s.t[q] = src[q]; }
    JB_SYNC();
    const uint8_t bad = s.t[k - 1];
    v.bad = bad;
    if (ins || bad == 'A' || bad == 'C' || bad == 'G' || bad == 'T') {      // (a lower-case suspect base has four substitutions)
      pack_bytes(s.pt, s.t, n);
      const int w = score_single_base(db, s, ins, base_code(bad), n, k, step, solid);
      if (w >= 0) {
        uint64_t occ, prevk;
        bool at;                                                  // the walker's i is where the reference's next step looks
        if (ins) { v.kind = kVIns; v.base = (uint8_t)("ATCG"[w]); at = gq == s0q + k - 1; occ = s.win[1]; prevk = s.fc[s0q - 1]; }
        else if (w == 3) { v.kind = kVDel; at = gq == s0q + k; occ = s.fc[s0q + k + 1]; prevk = s.win[0]; }
        else {
          int o = 0, left = w;                                    // the w-th of 'A','C','T','G' that is not the suspect base
          while (true) { if ((uint8_t)("ACTG"[o]) != bad) { if (left == 0) break; --left; } ++o; }
          v.kind = kVSub; v.base = (uint8_t)("ACTG"[o]); at = gq == s0q + k; occ = s.fc[s0q + k]; prevk = s.win[0];
        }
        // src/jasper.py:56-100 at the new i: neither weak nor below 1/50 of the window k bases back
        v.post_ok = (at && (int64_t)occ >= solid && !(occ * JB_DIVISOR < prevk)) ? 1 : 0;
        if (v.post_ok && second_step) {
          // ... and once more at i + k - 1: that window is untouched, the one k bases back still holds the new base
          // (substitution: the trial's last sample; insertion: one extra probe; deletion: untouched as well)
          const int64_t i2 = W0 + gq + (k - 1) - (ins ? 1 : 0) + (v.kind == kVDel ? 1 : 0);     // pass-start coordinate of the window
          if (i2 <= hi_p && okb[i2] != 0) {
            const uint64_t occ2 = cnt[i2];
            uint64_t prevk2;
            if (v.kind == kVSub) prevk2 = s.win[1];
            else if (v.kind == kVDel) prevk2 = s.fc[s0q + k];
            else {
              Key key = packed_window(s.pt, k - 2, 1);
              const Key b = {base_code(v.base), 0};
              key = key_append(key, b, 1);
              key = key_append(key, packed_window(s.pt, k - 1, k - 2), k - 2);
              prevk2 = db.count(key_canonical(key, k));
            }
            if ((int64_t)occ2 >= solid && !(occ2 * JB_DIVISOR < prevk2)) v.post_ok |= 2;
          }
        }
      } else if (w == -1 && !ins) {
        v.kind = kVNone;                                          // nothing passes and the reference stops there
      }
    }
  }
  else if (nbad > 1 && nbad < k - 1) {
    // src/jasper.py:434-459, first round: a base missing from a homopolymer run.  One more copy of the base in front of
    // k-1 and EVERY window of the trial string must be solid (nbad + 1 windows, each  t[m : k-1] ++ base ++ t[k-1 : ...]).
    // Anything but that immediate success is left to the second stage.
    const int n = nbad + k - 1;
    const uint8_t* src = in + (W0 + s0q);
    JB_SYNC();
    JB_LANES(l) { for (int q = l; q < n; q += 32) s.t[q] = src[q]; }
    JB_SYNC();
    const uint8_t sb = s.t[k - 2];
    pack_bytes(s.pt, s.t, n);
    const int total = nbad + 1;
    uint32_t weak_any = 0;
    for (int w0 = 0; w0 < total; w0 += 32) {
      JB_SYNC();
      JB_LANES(l) {
        const int m = w0 + l;
        uint32_t fail = 0;
        if (m < total) {
          const int lenA = (k - 1) - m;
          Key key = packed_window(s.pt, m, lenA);
          const Key b = {base_code(sb), 0};
          key = key_append(key, b, 1);
          key = key_append(key, packed_window(s.pt, k - 1, k - lenA - 1), k - lenA - 1);
          const uint32_t c2 = db.count(key_canonical(key, k));
          s.hc[m] = c2;
          fail = (int64_t)c2 < solid ? 1u : 0u;
        }
        s.flag[l] = fail;
      }
      JB_SYNC();
      for (int q = 0; q < 32; ++q) weak_any |= s.flag[q];
    }
    if (!weak_any) {
      v.kind = kVHomoIns; v.base = sb; v.bad = sb;
      // the two scan steps behind the fix: the windows that hold the new base are windows of the trial string
      const int64_t p1 = W0 + gq - k, p2 = W0 + gq + k - 2;
      if (gaq == gq && second_step && p1 >= 0 && okb[p1] != 0) {
        const uint64_t occ = s.hc[nbad], prevk = cnt[p1];
        if ((int64_t)occ >= solid && !(occ * JB_DIVISOR < prevk)) {
          v.post_ok = 1;
          if (p2 <= hi_p && okb[p2] != 0) {
            const uint64_t occ2 = cnt[p2], prevk2 = s.hc[nbad - 1];
            if ((int64_t)occ2 >= solid && !(occ2 * JB_DIVISOR < prevk2)) v.post_ok |= 2;
          }
        }
      }
    }
  }
  JB_SYNC();
  JB_LANES(l) { if (l == 0) *out = v; }
  JB_SYNC();
  *re_out = W0 + gq;
  return true;
}

}  // namespace jb

// jb_walker.cuh -- K4: the per-contig walker.  One WARP owns one contig and runs the reference's
// sequential state machine (src/jasper.py:50-104 scan, :150-223 handle_bad_kmers, :226-332
// fixing_sid, :340-583 the fix_* family, :585-599 check_sequence) with warp-uniform control flow;
// the 32 lanes cooperate on everything that is data-parallel:
//   * hopping over good sequence: 32 strided event bits per step (bitmask from k_scan),
//   * weak-run walks: 32 consecutive windows looked up per wave (speculative window cache),
//   * candidate scoring: all sampled k-mers of a trial string probed in one wave,
//   * de-Bruijn bridging: 8 paths x 4 extensions probed per wave,
//   * splicing edits into the output buffer.
// The mutating sequence is  out[0:olen] ++ in[ipos:]  (materialised prefix + untouched suffix);
// nothing is ever memmoved beyond the few bases between an edit and the materialised frontier.
//
// The same source compiles for the host (one "lane loop" instead of a warp): that build exists
// only for differential debugging in tests/ (tests/hostsim) and is never linked into the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/jasper_b200.h"
#include "jb_kmer.cuh"
#include "jb_cand.cuh"
#include "jb_jobs.cuh"
#include "jb_seq.cuh"
#include "jb_segments.cuh"
#include "jb_verdict.cuh"

// Inlined everywhere the kernel is 860 KB of SASS and the warps starve on instruction fetch (ncu: 60 %
// stall_no_inst); everything out of line costs registers and call overhead.  JB_WC marks the cold, bulky
// helpers (bridging, diploid, homopolymer repairs), JB_WN the hot ones; JB_NOINLINE_LEVEL picks who stays out.
#ifndef JB_NOINLINE_LEVEL
#define JB_NOINLINE_LEVEL 1
#endif
#if defined(__CUDACC__)
#define JB_W __device__
#if JB_NOINLINE_LEVEL >= 2
#define JB_WN __device__ __noinline__
#else
#define JB_WN __device__
#endif
#if JB_NOINLINE_LEVEL >= 1
#define JB_WC __device__ __noinline__
#else
#define JB_WC __device__
#endif
#else
#define JB_W
#define JB_WN
#define JB_WC
#endif

namespace jb {

constexpr int kPathCap = 4 * JB_MAX_PATHS + 64;  // a round can at most quadruple the path list
constexpr uint32_t kDeadNode = 0xFFFFFFFFu;

enum RowKind : uint32_t { kRowPlain = 0, kRowBridge = 1, kRowSpill = 2 };

// One fix event as logged by the walker, 32 bytes.  The text (mutation then original, the latter led by
// its s/i/d tag) sits inline when it fits in 8 bytes -- nearly always -- else in the aux arena; a bridged
// multi-error patch keeps its patch and the replaced bases in the aux arena for the host-side alignment.
struct DevRow {
  int64_t coord;
  uint32_t unit;
  uint32_t seg;              // segment of the contig that emitted it
  uint32_t seqno;            // emission order inside the segment
  uint8_t gen;               // which launch of the pass produced it (a re-run supersedes the first attempt)
  uint8_t kind;              // RowKind
  uint8_t fb_len, orig_len;  // kRowPlain / kRowSpill: lengths of mutation / original (tag included)
  union {
    char text[8];            // kRowPlain
    uint64_t aux_off;        // kRowSpill: text in the aux arena; kRowBridge: {u32 patch_len, u32 orig_len, patch, original}
  };
};
static_assert(sizeof(DevRow) == 32, "DevRow is meant to be 32 bytes");

struct RowSink {
  DevRow* rows;
  unsigned long long cap;
  unsigned long long* used;
  uint8_t* aux;
  unsigned long long aux_cap;
  unsigned long long* aux_used;
};

struct BridgeScratch {       // global memory, one per resident walker
  Key km[2][kPathCap];
  uint32_t node[2][kPathCap];
  uint32_t* trie;            // (parent << 2) | base
  uint32_t trie_cap;
  uint8_t* patch;            // room for the recovered patch
  uint32_t patch_cap;
};

constexpr uint8_t kTagBridge = 0xFF;      // Verdict2.tag of a bridged multi-error patch
enum PostKind : int { kPostNone = 0, kPostSub = 1, kPostDel = 2, kPostIns = 3 };
constexpr int kFarBack = -102;            // walker status: it needed bases behind its materialised window
constexpr int kTrieFull = -103;           // walker status: the de-Bruijn search outgrew its trie (the host re-walks the contig with a grown pool)
constexpr int64_t kMinGap = 256;          // clean stretches longer than this are not materialised
JB_HD int default_keep_margin(int k) { return 2 * k + 32; }

struct BridgePool {          // bridging scratch is borrowed only while a multi-error patch is being searched
  BridgeScratch* slots;
  uint32_t n;
  uint32_t* busy;            // device: 0 = free, 1 = held
};

struct UnitIO {               // one SEGMENT of one contig (a whole contig is the one-segment case)
  const uint8_t* in;         // the contig at the start of the pass
  int64_t in_len;
  uint8_t* out;              // staging: out[r] = new base r of this segment, valid inside materialised windows
  int64_t out_cap;
  const uint32_t* cnt;       // C(window) for every position of the pass-start contig, [0, in_len]; may be null
  const uint8_t* okb;        // 1 = the pass-start window at that position is k chars of ACGTacgt; may be null
  const uint32_t* weak;      // weak-window bit words of this contig (valid window with count < solid); null = no verdicts
  const uint32_t* vmap;      // flat: word (ev_bit_base + re) >> 5 holds verdict id + 1 of the run that ends at re, 0, or kVmapAmbiguous
  const Verdict* verdicts;   // k_verdicts output
  const Verdict2* verdicts2; // k_verdicts2 output, same index; may be null
  const uint32_t* ev;        // event words of this contig (bit p of the pass-start sequence); may be null
  const uint32_t* ev_l1;     // level-1 summary over FLAT bit positions (1 bit per 1024); may be null
  const uint32_t* ev_l2;     // level-2 summary (1 bit per 32768 flat positions = one level-1 word); may be null
  uint64_t ev_bit_base;      // flat bit index of contig position 0
  int64_t base;              // old coordinate where this segment starts (0 or an anchor)
  int64_t next_anchor;       // old coordinate of the next segment's anchor, or -1 for the last segment
  uint32_t unit;             // contig index
  uint32_t seg;              // segment index inside the contig
  uint32_t gen;
  uint32_t slot;             // index of this descriptor in the launch's array (copy jobs refer to it)
  uint32_t materialize_all;  // 1: never skip clean stretches (the conservative mode of re-runs)
  int32_t keep_margin;       // bases kept materialised in front of an edit (default_keep_margin(k); tests shrink it)
  // results: old [base, end_old) became new_len bases, described by this segment's copy jobs
  int64_t end_old;
  int64_t new_len;
  int64_t wrong;
  int32_t status;
  uint32_t n_rows;
  uint32_t n_fast;             // events taken from a verdict (jb_verdict.cuh)
  uint32_t n_post;             // ... whose following scan step was taken from data in hand
  uint32_t n_v2;               // events whose fix (or "no fix") came from a second-stage verdict
  uint32_t pad2;
  unsigned long long lookups;
  unsigned long long events;
  unsigned long long cycles;   // device clock ticks this segment took (profiling aid)
};

// append one copy job (out of line: four call sites in the walker)
JB_FREE_NOINLINE bool push_job(JobSink jobs, uint32_t seg_slot, const uint8_t* src, int64_t rel, int64_t len) {
  JB_SYNC();
  unsigned long long slot = 0;
#if defined(__CUDA_ARCH__)
  if ((threadIdx.x & 31) == 0) slot = atomicAdd(jobs.used, 1ull);
  slot = __shfl_sync(0xffffffffu, slot, 0);
#else
  slot = (*jobs.used)++;
#endif
  if (slot >= jobs.cap) return false;
  JB_LANES(l) { if (l == 0) { CopyJob j; j.src = src; j.rel = rel; j.len = len; j.slot = seg_slot; j.pad = 0; jobs.jobs[slot] = j; } }
  JB_SYNC();
  return true;
}

template <class DB>
struct Walker {
  const DB& db;
  const int k;
  const int step;
  const int64_t solid;
  const bool fix;
  const Key mask;
  UnitIO& u;
  UnitIO c;                            // the segment descriptor, read once (u lives in global memory)
  WarpScratch& s;
  BridgePool pool;
  BridgeScratch* bs = nullptr;
  int bs_slot = -1;
  RowSink sink;
  JobSink jobs;

  int64_t i = 0, olen = 0, ipos = 0, L = 0, clean_from = 0;
  int64_t wstart = 0;                  // out[wstart:olen) is the materialised window; [0, wstart) is closed (copy jobs)
  int64_t keep_margin = 0;             // bases kept materialised in front of an edit
  bool far_back = false;
  int64_t base = 0, seam_margin = 0;   // seam_margin: how far before `base` this walker may read / how far before the
                                       // next anchor the previous walker must have stopped editing
  int64_t wrong = 0;
  int status = 0;
  bool brk = false;
  bool seam_bad = false;
  uint32_t vflags = 0;                 // kFlagSeam / kFlagFar raised by reads (lane-local until the end of run)
  uint32_t seqno = 0;
  unsigned long long lookups = 0, events = 0;
  uint32_t n_fast = 0, n_post = 0, n_v2 = 0;
  uint32_t vid = 0;                    // index of the verdict consume_verdict returned
  bool dry = false;                    // k_verdicts2: edits go to scratch, nothing is logged
  int v2_post = 0;                     // scan-step bits of the second-stage verdict just applied (bit 0: next step plain, bit 1: the one after)
  long long t_begin = 0;

  JB_W Walker(const DB& db_, int k_, int step_, uint32_t solid_, bool fix_, UnitIO& u_, WarpScratch& s_,
               BridgePool pool_, RowSink sink_, JobSink jobs_)
      : db(db_), k(k_), step(step_), solid((int64_t)solid_), fix(fix_), mask(key_mask(k_)), u(u_), s(s_), pool(pool_), sink(sink_), jobs(jobs_) {}

  // ---------------------------------------------------------------------------- bridging scratch
  JB_WC bool acquire_bridge() {
    if (!pool.slots || !pool.n) { status = JB_E_SCRATCH; return false; }
#if defined(__CUDA_ARCH__)
    int got = -1;
    if ((threadIdx.x & 31) == 0) {
      uint32_t at = (uint32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % pool.n;
      unsigned int backoff = 200;
      while (got < 0) {                                  // holders never wait on anything, so this terminates
        for (uint32_t tries = 0; tries < 16 && got < 0; ++tries) {
          if (*(volatile uint32_t*)&pool.busy[at] == 0u && atomicCAS(&pool.busy[at], 0u, 1u) == 0u) got = (int)at;
          else at = at + 1 == pool.n ? 0 : at + 1;
        }
        if (got < 0) { __nanosleep(backoff); if (backoff < 20000) backoff *= 2; }   // all busy: stay off the issue slots
      }
      __threadfence();
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    bs_slot = got;
#else
    bs_slot = 0;
#endif
    bs = pool.slots + bs_slot;
    return true;
  }
  JB_WC void release_bridge() {
    if (bs_slot < 0) return;
#if defined(__CUDA_ARCH__)
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { __threadfence(); atomicExch(&pool.busy[bs_slot], 0u); }
#endif
    bs = nullptr; bs_slot = -1;
  }

  // ---------------------------------------------------------------------------- sequence access (jb_seq.cuh)
  // new sequence (segment-local coordinates) = in[0:base) ++ out[0:olen) ++ in[ipos:)
  // the walker's registers hold the state; the out-of-line helpers read this published copy
  JB_W void publish() {
    JB_SYNC();
    JB_LANES(l) {
      if (l == 0) {
        SeqView v;
        v.in = c.in; v.out = c.out; v.cnt = c.cnt;
        v.base = base; v.olen = olen; v.ipos = ipos; v.wstart = wstart; v.L = L; v.in_len = c.in_len;
        v.clean_from = clean_from; v.seam_margin = seam_margin;
        s.view = v;
        s.wc_n = 0;                                   // cached window counts do not survive an edit
      }
    }
    JB_SYNC();
  }
  JB_W void note(uint32_t flags) { vflags |= flags & (kFlagSeam | kFlagFar); }

  JB_W uint8_t get(int64_t x) { return view_get(s.view, x, vflags); }

  // dst[0:n] = seq[a:b] clamped to [0, L); returns n
  JB_W int load(uint8_t* dst, int64_t a, int64_t b) {
    const uint32_t r = view_load(&s.view, dst, a, b);
    note(r >> 16);
    return (int)(r & 0xFFFFu);
  }

  JB_W uint32_t count_in(const uint8_t* buf, int len, int m) const { return count_str(db, buf, len, m, k, mask); }

  // ---------------------------------------------------------------------------- lookups
  // C(seq[x:x+k]); a miss refills the 32-window cache in direction `dir` (jb_seq.cuh)
  JB_W uint32_t Cw(int64_t x, int dir) {
    const uint64_t r = view_count(db, s, x, dir, k, mask);
    note((uint32_t)(r >> 32));
    return (uint32_t)r;
  }

  // n <= 32 windows at arbitrary starts s.px[0..n) (non-negative, within a span that fits s.c)
  JB_W void Cpts(int n) {
    const uint32_t f = view_points(db, s, n, k, mask);
    if (f & kFlagTooWide) {                          // cannot happen for the call sites below; stay exact anyway
      for (int q = 0; q < n; ++q) s.pr[q] = Cw(s.px[q], +1);
      return;
    }
    note(f);
    if (f & kFlagProbed) lookups += n;
  }

  // src/jasper.py:585-599 on a string in scratch memory
  JB_W bool check_sequence(const uint8_t* buf, int len, int64_t thr) {   // (bridging only; the fix_* family scores through jb_cand.cuh)
    // sample starts: 0, max(0,len-k), then step, 2*step, ... < len-k
    int n_mid = 0;
    for (int m = step; m < len - k; m += step) ++n_mid;
    int total = 2 + n_mid;
    bool ok = true;
    pack_bytes(s.pc, buf, len);
    for (int w0 = 0; w0 < total && ok; w0 += 32) {
      JB_SYNC();
      JB_LANES(l) {
        int q = w0 + l;
        uint32_t f = 0;
        if (q < total) {
          int m = q == 0 ? 0 : (q == 1 ? (len - k > 0 ? len - k : 0) : step * (q - 1));
          f = (int64_t)count_packed(db, s.pc, buf, len, m, k, mask) < thr ? 1u : 0u;
        }
        s.flag[l] = f;
      }
      JB_SYNC();
      int cnt = total - w0 < 32 ? total - w0 : 32;
      lookups += cnt;
      for (int q = 0; q < cnt; ++q) ok = ok && !s.flag[q];
    }
    return ok;
  }

  // ---------------------------------------------------------------------------- edits
  JB_W bool emit_job(const uint8_t* src, int64_t rel, int64_t len) {
    if (len <= 0 || dry) return true;                 // (a dry run, k_verdicts2, describes no output)
    if (!push_job(jobs, c.slot, src, rel, len)) { status = JB_E_SCRATCH; return false; }
    return true;
  }

  // materialise the new sequence up to upto_abs for an edit that starts at a_abs; a long clean stretch
  // in front of the edit is handed to k_concat as a copy job instead of being copied here
  JB_WN bool ensure_out(int64_t a_abs, int64_t upto_abs) {
    const int64_t upto = upto_abs - base;
    if (upto <= olen) return true;
    if (upto > c.out_cap) { status = JB_E_SCRATCH; return false; }
    if (!c.materialize_all) {
      const int64_t keep_from = (a_abs - base) - keep_margin;
      if (keep_from > olen + kMinGap) {
        if (!emit_job(c.out + wstart, wstart, olen - wstart)) return false;
        const int64_t gap = keep_from - olen;
        if (!emit_job(c.in + ipos, olen, gap)) return false;
        ipos += gap; olen += gap; wstart = olen;
      }
    }
    const int64_t n = upto - olen;
    JB_SYNC();
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int64_t q0 = 0; q0 < n; q0 += 128) {           // 4 independent loads per lane before the stores (a window is ~2k+32 bases)
      uint8_t v[4];
      JB_LANES(l) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { int64_t q = q0 + l + 32 * j; v[j] = q < n ? c.in[ipos + q] : (uint8_t)0; }
#pragma unroll
        for (int j = 0; j < 4; ++j) { int64_t q = q0 + l + 32 * j; if (q < n) c.out[olen + q] = v[j]; }
      }
    }
    JB_SYNC();
    ipos += n; olen = upto;
    return true;
  }

  // seq = seq[:a] + repl + seq[e:]
  JB_WN bool apply_edit(int64_t a_abs, int64_t e_abs, const uint8_t* repl, int64_t rlen) {
    if (a_abs < base) { status = kSeamViolation; return false; }
    if (a_abs - base < wstart) { status = kFarBack; return false; }
    if (!ensure_out(a_abs, e_abs)) return false;
    const int64_t a = a_abs - base, e = e_abs - base;   // offsets into out
    const int64_t delta = rlen - (e - a);
    const int64_t tail = olen - e;
    if (olen + delta > c.out_cap) { status = JB_E_SCRATCH; return false; }
    if (delta != 0 && tail > 0) {
      // shift out[e:olen) by delta, 32 bytes per step, in the safe direction
      if (delta > 0) {
        for (int64_t hi = tail; hi > 0; hi -= 32) {
          int64_t lo = hi - 32 > 0 ? hi - 32 : 0;
          uint8_t v = 0;
          JB_SYNC();
#if defined(__CUDA_ARCH__)
          { int l = threadIdx.x & 31; if (lo + l < hi) v = c.out[e + lo + l]; __syncwarp(); if (lo + l < hi) c.out[e + lo + l + delta] = v; }
#else
          { uint8_t tmp[32]; for (int l = 0; l < 32; ++l) if (lo + l < hi) tmp[l] = c.out[e + lo + l];
            for (int l = 0; l < 32; ++l) if (lo + l < hi) c.out[e + lo + l + delta] = tmp[l]; (void)v; }
#endif
        }
      } else {
        for (int64_t lo = 0; lo < tail; lo += 32) {
          int64_t hi = lo + 32 < tail ? lo + 32 : tail;
          uint8_t v = 0;
          JB_SYNC();
#if defined(__CUDA_ARCH__)
          { int l = threadIdx.x & 31; if (lo + l < hi) v = c.out[e + lo + l]; __syncwarp(); if (lo + l < hi) c.out[e + lo + l + delta] = v; }
#else
          { uint8_t tmp[32]; for (int l = 0; l < 32; ++l) if (lo + l < hi) tmp[l] = c.out[e + lo + l];
            for (int l = 0; l < 32; ++l) if (lo + l < hi) c.out[e + lo + l + delta] = tmp[l]; (void)v; }
#endif
        }
      }
    }
    JB_SYNC();
    JB_LANES(l) { for (int64_t q = l; q < rlen; q += 32) c.out[a + q] = repl[q]; }
    JB_SYNC();
    olen += delta; L += delta;
    int64_t shifted = clean_from + delta;
    clean_from = a_abs + rlen > shifted ? a_abs + rlen : shifted;
    publish();
    return true;
  }

  JB_WN void emit_row(int64_t coord, const uint8_t* fb, int fb_len, uint8_t tag, const uint8_t* orig, int orig_len) {
    JB_SYNC();
    const int total = fb_len + 1 + orig_len;
    unsigned long long slot = 0, off = 0;
#if defined(__CUDA_ARCH__)
    if ((threadIdx.x & 31) == 0) {
      slot = atomicAdd(sink.used, 1ull);
      if (total > 8) off = atomicAdd(sink.aux_used, (unsigned long long)total);
    }
    slot = __shfl_sync(0xffffffffu, slot, 0); off = __shfl_sync(0xffffffffu, off, 0);
#else
    slot = (*sink.used)++;
    if (total > 8) { off = *sink.aux_used; *sink.aux_used += (unsigned long long)total; }
#endif
    if (slot >= sink.cap || (total > 8 && off + total > sink.aux_cap)) { status = JB_E_SCRATCH; return; }
    DevRow& r = sink.rows[slot];
    char* text = total > 8 ? (char*)sink.aux + off : r.text;
    JB_LANES(l) {
      if (l == 0) {
        r.coord = coord; r.unit = c.unit; r.seg = c.seg; r.gen = (uint8_t)c.gen; r.seqno = seqno;
        r.kind = (uint8_t)(total > 8 ? kRowSpill : kRowPlain);
        r.fb_len = (uint8_t)fb_len; r.orig_len = (uint8_t)(orig_len + 1);
        if (total > 8) r.aux_off = off;
        text[fb_len] = (char)tag;
      }
    }
    JB_SYNC();                                       // (aux_off and inline text share storage: header first)
    JB_LANES(l) {
      for (int q = l; q < fb_len; q += 32) text[q] = (char)fb[q];
      for (int q = l; q < orig_len; q += 32) text[fb_len + 1 + q] = (char)orig[q];
    }
    JB_SYNC();
    ++seqno;
  }

  // ---------------------------------------------------------------------------- fix_* family
  // All take the to_be_fixed string in s.t (length n) that starts at sequence position s0.

  JB_W void set_cand(int idx, uint8_t kind, uint8_t b1, uint8_t b2, int pos, int m) {
    JB_LANES(l) { if (l == 0) { Cand c; c.kind = kind; c.base = b1; c.base2 = b2; c.pad = 0; c.pos = (int16_t)pos; c.m = (int16_t)m; s.cand[idx] = c; } }
  }
  JB_W void fill(uint8_t* dst, uint8_t v, int m) { JB_SYNC(); JB_LANES(l) { for (int q = l; q < m; q += 32) dst[q] = v; } JB_SYNC(); }

  // What a fix_* function decided: seq = seq[:a] + repl + seq[e:], and the csv row that describes it (text in
  // s.fx).  fixing_sid applies it in ONE place, so the splice / job / row code exists once in the kernel.
  struct Fix {
    bool found = false;
    int64_t a = 0, e = 0;
    const uint8_t* repl = nullptr;
    int64_t rlen = 0;
    int64_t coord = 0;
    int fb_len = 0, orig_len = 0;
    uint8_t tag = 0;                                  // 0: no plain row (bridged patches log their own)
    int post = 0;                                     // PostKind: a single-base fix scored by score_single_base (s.win holds its counts)
  };
  JB_W void fix_row(Fix& f, int64_t coord, uint8_t fb, int fb_len, uint8_t tag, uint8_t orig, int orig_len) {
    fill(s.fx, fb, fb_len);
    fill(s.fx + 64, orig, orig_len);
    f.coord = coord; f.fb_len = fb_len; f.tag = tag; f.orig_len = orig_len;
  }

  // src/jasper.py:392-406 then :409-419: up to three substitutions at k-1 in order 'ACTG', then the deletion of k-1
  JB_W Fix fix_k_case(int n, int64_t thr, int64_t s0, int64_t ga) {
    Fix f;
    const uint8_t bad = s.t[k - 1];
    const uint8_t before = get(ga - 1);
    const char order[4] = {'A', 'C', 'T', 'G'};
    int nc = 0;
    pack_bytes(s.pt, s.t, n);
    for (int o = 0; o < 4; ++o) if ((uint8_t)order[o] != bad) set_cand(nc++, kSub, (uint8_t)order[o], 0, k - 1, 0);
    set_cand(nc++, kDel, 0, 0, k - 1, 1);
    const int w = first_passing(db, s, nc, n, k, step, mask, thr, &lookups);
    if (w < 0) return f;
    const Cand c = s.cand[w];
    f.found = true; f.a = s0 + k - 1; f.e = s0 + k; f.repl = s.b;
    if (c.kind == kSub) { fill(s.b, c.base, 1); f.rlen = 1; fix_row(f, ga - 1, c.base, 1, 's', before, 1); }
    else { f.rlen = 0; fix_row(f, ga - 1, '-', 1, 'i', before, 1); }
    return f;
  }

  // src/jasper.py:422-431: one base inserted in front of k-1, order 'ATCG'
  JB_W Fix fix_del(int n, int64_t thr, int64_t s0, int64_t ga) {
    Fix f;
    const char order[4] = {'A', 'T', 'C', 'G'};
    pack_bytes(s.pt, s.t, n);
    for (int o = 0; o < 4; ++o) set_cand(o, kIns, (uint8_t)order[o], 0, k - 1, 1);
    const int w = first_passing(db, s, 4, n, k, step, mask, thr, &lookups);
    if (w < 0) return f;
    const uint8_t bb = s.cand[w].base;
    fill(s.b, bb, 1);
    f.found = true; f.a = s0 + k - 1; f.e = s0 + k - 1; f.repl = s.b; f.rlen = 1;
    fix_row(f, ga, bb, 1, 'd', '-', 1);
    return f;
  }

  // src/jasper.py:340-388
  JB_WC Fix fix_diploid(int n, int64_t thr, int64_t s0, int64_t gb, int64_t ga) {
    Fix f;
    const uint8_t left_bad = s.t[n - k], right_bad = s.t[k - 1];
    const int64_t g0 = gb - k + 1 > 0 ? gb - k + 1 : 0;
    const int na = (2 * k - 1 - n) / 2;
    int64_t a0, a1;
    if (ga + k - 1 + na < L) { a0 = ga + k - 1; a1 = ga + k - 1 + na; }
    else { a0 = (L - 1 < ga + k - 1) ? L - 1 : ga + k - 1; a1 = L; }
    const int bl = (int)(a1 > a0 ? a1 - a0 : 0);
    const int64_t b0 = g0 - bl + 1 > 0 ? g0 - bl + 1 : 0;
    const int nb = (int)(g0 + 1 - b0);
    // check string = before + trial + after: the flanks are staged once in s.c
    const uint8_t seq_ga1 = get(ga - 1), seq_gb1 = get(gb + 1);
    load(s.c, b0, b0 + nb);
    load(s.c + nb + n, a0, a0 + bl);
    JB_SYNC();
    JB_LANES(l) {
      for (int q = l; q < n; q += 32) s.c[nb + q] = s.t[q];          // the whole check string, for packing
      if (l == 0) { s.dip_nb = nb; s.dip_bl = bl; }
    }
    JB_SYNC();
    pack_bytes(s.pt, s.c, nb + n + bl);
    const char order[4] = {'A', 'C', 'T', 'G'};
    int nc = 0;
    for (int xo = 0; xo < 4; ++xo)
      for (int yo = 0; yo < 4; ++yo) {
        const uint8_t x = (uint8_t)order[xo], y = (uint8_t)order[yo];
        if (x == left_bad && y == right_bad) continue;
        if (x != left_bad && y != right_bad) continue;
        set_cand(nc++, kDip, x, y, 0, 0);
      }
    const int w = first_passing(db, s, nc, n, k, step, mask, thr, &lookups);
    if (w < 0) return f;
    const Cand c = s.cand[w];
    // one edit covering both positions: t[n-k : k] with x / y substituted
    JB_SYNC();
    JB_LANES(l) { for (int q = l; q < 2 * k - n; q += 32) s.b[q] = cand_char(s, c, n, k, nb + (n - k) + q); }
    JB_SYNC();
    f.found = true; f.a = s0 + n - k; f.e = s0 + k; f.repl = s.b; f.rlen = 2 * k - n;
    if (c.base == left_bad) fix_row(f, gb + 1, c.base2, 1, 's', seq_gb1, 1);       // "e": the right-hand base changed
    else fix_row(f, ga - 1, c.base, 1, 's', seq_ga1, 1);                           // "s": the left-hand base changed
    return f;
  }

  // src/jasper.py:434-477
  JB_WC Fix fix_same_base_del(int n, int64_t thr, int64_t s0) {
    Fix f;
    if (thr > solid) return f;
    const uint8_t sb = s.t[k - 2];
    const int start_bad = n - k + 1;
    int cur = start_bad, added = 0;
    pack_bytes(s.pt, s.t, n);
    while (added < start_bad) {
      ++added;
      JB_SYNC();
      set_cand(0, kIns, sb, 0, k - 1, added);
      const int nbw = count_weak_cand(db, s, 0, n, k, mask, thr, &lookups);
      if (nbw == 0) {
        fill(s.b, sb, added);
        f.found = true; f.a = s0 + k - 1; f.e = s0 + k - 1; f.repl = s.b; f.rlen = added;
        fix_row(f, k - 1 + s0, sb, added, 'd', '-', 1);
        return f;
      }
      if (nbw >= cur) break;
      cur = nbw;
    }
    const char order[4] = {'A', 'T', 'C', 'G'};
    JB_SYNC();
    for (int o = 0; o < 4; ++o) set_cand(o, kIns, (uint8_t)order[o], 0, k - 2, 1);
    const int w = first_passing(db, s, 4, n, k, step, mask, thr, &lookups);
    if (w < 0) return f;
    const uint8_t bb = s.cand[w].base;
    fill(s.b, bb, 1);
    f.found = true; f.a = s0 + k - 2; f.e = s0 + k - 2; f.repl = s.b; f.rlen = 1;
    fix_row(f, k - 2 + s0, bb, 1, 'd', '-', 1);
    return f;
  }

  // src/jasper.py:479-524
  JB_WC Fix fix_same_base_insertion(int n, int64_t thr, int64_t s0) {
    Fix f;
    if (thr > solid) return f;
    const uint8_t sb = s.t[k - 1];
    const int start_bad = n - k + 1;
    int cur = start_bad, removed = 0;
    pack_bytes(s.pt, s.t, n);
    while (removed < start_bad) {
      --cur; ++removed;
      if (n - removed == k) break;
      JB_SYNC();
      set_cand(0, kDel, 0, 0, k - 1, removed);
      const int nbw = count_weak_cand(db, s, 0, n, k, mask, thr, &lookups);
      if (nbw == 0) {
        f.found = true; f.a = s0 + k - 1; f.e = s0 + k - 1 + removed; f.repl = s.b; f.rlen = 0;
        fix_row(f, k - 1 + s0, '-', 1, 'i', sb, removed);
        return f;
      }
      if (nbw >= cur) break;
      cur = nbw;
    }
    int nc = 0;
    JB_SYNC();
    for (int m = n - k; m < n - 1 && nc < kMaxCands; ++m) set_cand(nc++, kDel, 0, 0, m, 1);
    const int w = first_passing(db, s, nc, n, k, step, mask, thr, &lookups);
    if (w < 0) return f;
    const int m = s.cand[w].pos;
    f.found = true; f.a = s0 + m; f.e = s0 + m + 1; f.repl = s.b; f.rlen = 0;
    fix_row(f, m + s0, '-', 1, 'i', s.t[m], 1);
    return f;
  }

  // ---------------------------------------------------------------------------- base_extension
  // src/jasper.py:527-583.  Paths live as (packed last k-1 bases, trie leaf); s.t holds good_kmer_before
  // (k chars), s.b holds good_k_mer_after (k chars).  Returns patch length (>= 1) with the patch in
  // bs->patch, 0 for "None".
  JB_W uint32_t trie_new(uint32_t parent, uint32_t base, uint32_t& n_nodes) {
    if (n_nodes >= bs->trie_cap) { status = kTrieFull; return 0; }
    uint32_t id = n_nodes++;
    JB_LANES(l) { if (l == 0) bs->trie[id] = (parent << 2) | base; }
    return id;
  }

  JB_WC int64_t base_extension(int len_tbf, int64_t thr) {
    if (thr > solid) return 0;
    const uint8_t* G = s.t;          // good_kmer_before
    const uint8_t* A = s.b;          // good_k_mer_after
    const int64_t span = (int64_t)len_tbf - 2 * k;
    const int64_t max_ext = (int64_t)rint((double)span * 1.2) + JB_MIN_OVERLAP + JB_SLACK;
    const int64_t min_patch = (int64_t)rint((double)span / 1.2) - JB_SLACK;
    // Codes of "ACGT" upper case for the overlap test
    const char up[4] = {'A', 'C', 'G', 'T'};
    uint32_t n_nodes = 1;            // node 0 = root = the single start path (its char is G[k-1], kept raw)
    // The path list (packed last k-1 bases, trie leaf) ping-pongs between two buffers.  While it is short -- nearly always:
    // one path, now and then a bubble -- the buffers are the warp's own shared memory (s.pkm / s.pnode); a list that has
    // more than kSmallPaths / 4 entries at the start of a round (a round can quadruple it) moves to the bridging slot's
    // arrays in global memory for the rest of the search.  (r2aa: every round of a one-path search paid three or four
    // dependent round trips to global memory for a 20-byte list.)
    Key* km_cur = s.pkm[0]; Key* km_nxt = s.pkm[1];
    uint32_t* nd_cur = s.pnode[0]; uint32_t* nd_nxt = s.pnode[1];
    bool small = true;
    // positions of G that are not ACGTacgt (bit q = G[q]); k <= 48
    uint64_t g_inv = 0;
    for (int q = 0; q < k; ++q) if (!base_valid(G[q])) g_inv |= 1ull << q;
    // km1 at round 1 = G[1:k]
    Key km0 = {0, 0};
    for (int q = 1; q < k; ++q) km0 = key_push(km0, base_valid(G[q]) ? base_code(G[q]) : 0u, mask);
    const Key mask_km1 = key_mask(k - 1);
    km0.lo &= mask_km1.lo; km0.hi &= mask_km1.hi;
    JB_SYNC();
    JB_LANES(l) { if (l == 0) { km_cur[0] = km0; nd_cur[0] = 0; bs->trie[0] = 0; } }
    JB_SYNC();
    int n_paths = 1;
    for (int64_t it = 1; it < max_ext; ++it) {
      // ---- paths = [l for l in paths if len(l) > 0]
      {
        int w = 0;
        for (int p0 = 0; p0 < n_paths; p0 += 32) {
          JB_SYNC();
          JB_LANES(l) { s.flag[l] = (p0 + l < n_paths && nd_cur[p0 + l] != kDeadNode) ? 1u : 0u; }
          JB_SYNC();
          int run = w;
          JB_LANES(l) {
            if (s.flag[l]) {
              int off = 0;
              for (int q = 0; q < l; ++q) off += (int)s.flag[q];
              km_nxt[w + off] = km_cur[p0 + l];
              nd_nxt[w + off] = nd_cur[p0 + l];
            }
          }
          for (int q = 0; q < 32; ++q) run += (int)s.flag[q];
          w = run;
          JB_SYNC();
        }
        { Key* tk = km_cur; km_cur = km_nxt; km_nxt = tk; uint32_t* tn = nd_cur; nd_cur = nd_nxt; nd_nxt = tn; }
        n_paths = w;
        if (small && n_paths > kSmallPaths / 4) {        // too long for shared memory by the end of this round: move out
          JB_SYNC();
          JB_LANES(l) { for (int q = l; q < n_paths; q += 32) { bs->km[0][q] = km_cur[q]; bs->node[0][q] = nd_cur[q]; } }
          JB_SYNC();
          km_cur = bs->km[0]; km_nxt = bs->km[1]; nd_cur = bs->node[0]; nd_nxt = bs->node[1];
          small = false;
        }
      }
      if (n_paths > JB_MAX_PATHS) return 0;
      if (n_paths == 0) return 0;                      // the reference keeps looping over an empty list: None
      // first invalid base of km1 = G[it:k] ++ new bases (only G can hold invalid bytes)
      int first_inv = -1;
      if (it < k && (g_inv >> it)) { int q = 0; while (!((g_inv >> (it + q)) & 1ull)) ++q; first_inv = q; }
      const bool overlap_round = it >= JB_MIN_OVERLAP && it >= min_patch;
      const int last_path = n_paths;
      for (int p0 = 0; p0 < last_path; p0 += 8) {
        // ---- 8 paths x 4 bases per wave
        JB_SYNC();
        JB_LANES(l) {
          int p = p0 + (l >> 2);
          uint32_t ok = 0;
          if (p < last_path) {
            Key full = key_push(km_cur[p], (uint32_t)(l & 3), mask);   // km1 + base: k bases
            if (first_inv >= 0) {                                          // MerDNA stops at the invalid byte
              int keep = first_inv;                                        // bases kept
              Key trimmed = {0, 0};
              if (keep > 0) {
                int drop = 2 * (k - keep);
                if (drop >= 64) { trimmed.hi = (full.hi >> (drop - 64)) << (drop - 64); trimmed.lo = 0; }
                else { trimmed.hi = full.hi; trimmed.lo = (full.lo >> drop) << drop; }
              }
              full = trimmed;
            }
            ok = (int64_t)db.count(key_canonical(full, k)) >= thr ? 1u : 0u;
          }
          s.ext[l] = ok;
        }
        JB_SYNC();
        int np = last_path - p0 < 8 ? last_path - p0 : 8;
        lookups += 4 * np;
        for (int pp = 0; pp < np; ++pp) {
          const int p = p0 + pp;
          const uint32_t old_node = nd_cur[p];
          const Key old_km = km_cur[p];
          int extended = 0;
          for (int j = 0; j < 4; ++j) {
            if (!s.ext[pp * 4 + j]) continue;
            Key nk = key_push(old_km, (uint32_t)j, mask_km1);
            if (overlap_round) {
              // (km1 + base)[-5:] == good_k_mer_after[0:5]: four upper-case path bases then base j
              bool same = (uint8_t)up[j] == A[JB_MIN_OVERLAP - 1];
              for (int q = 0; q < JB_MIN_OVERLAP - 1 && same; ++q) {
                uint32_t code = (uint32_t)((old_km.lo >> (2 * (JB_MIN_OVERLAP - 2 - q))) & 3u);
                same = (uint8_t)up[code] == A[q];
              }
              if (same) {
                // body = path string of old_node (length it).  Rebuild the last k+3 chars of
                // stem + body into s.c, append base j and good_k_mer_after[-(k-5):]  (2k-1 chars)
                const int need = k + 3;                     // chars of (stem + body) that survive [-(2k-1):]
                JB_SYNC();
                JB_LANES(l) {
                  if (l == 0) {
                    uint32_t nd = old_node;
                    int64_t depth = it;                     // body has `it` chars; body[0] = G[k-1]
                    for (int q = need - 1; q >= 0; --q) {
                      uint8_t ch;
                      if (depth > 1) { ch = (uint8_t)up[bs->trie[nd] & 3u]; nd = bs->trie[nd] >> 2; --depth; }
                      else if (depth == 1) { ch = G[k - 1]; --depth; }
                      else { int64_t gi = (k - 1) - (int64_t)(need - (int64_t)it - q) ; ch = G[gi]; }
                      s.c[q] = ch;
                    }
                    s.c[need] = (uint8_t)up[j];
                    for (int q = 0; q < k - JB_MIN_OVERLAP; ++q) s.c[need + 1 + q] = A[JB_MIN_OVERLAP + q];
                  }
                }
                JB_SYNC();
                if (check_sequence(s.c, 2 * k - 1, thr)) {
                  if (it == JB_MIN_OVERLAP) return 0;
                  // patch = (body + base)[1:-5] : it - 5 chars, recovered through the trie
                  const int64_t plen = it - JB_MIN_OVERLAP;
                  if ((uint64_t)plen > bs->patch_cap) { status = JB_E_SCRATCH; return 0; }
                  JB_SYNC();
                  JB_LANES(l) {
                    if (l == 0) {
                      // (body + base) has it+1 chars; drop the last 5 => chars [1, it-4)
                      uint32_t nd = old_node;
                      int64_t idx = it - 1;                 // index of body's last char
                      // skip the 4 body chars that fall in the dropped tail (base j is the 5th)
                      for (int q = 0; q < JB_MIN_OVERLAP - 1; ++q) { nd = bs->trie[nd] >> 2; --idx; }
                      for (; idx >= 1; --idx) { bs->patch[idx - 1] = (uint8_t)up[bs->trie[nd] & 3u]; nd = bs->trie[nd] >> 2; }
                    }
                  }
                  JB_SYNC();
                  return plen;
                }
              }
            }
            if (!extended) {
              uint32_t nn = trie_new(old_node, (uint32_t)j, n_nodes);
              if (status) return 0;
              JB_SYNC();
              JB_LANES(l) { if (l == 0) { km_cur[p] = nk; nd_cur[p] = nn; } }
              extended = 1;
            } else {
              if (n_paths >= kPathCap) { status = JB_E_SCRATCH; return 0; }
              uint32_t nn = trie_new(old_node, (uint32_t)j, n_nodes);
              if (status) return 0;
              JB_SYNC();
              JB_LANES(l) { if (l == 0) { km_cur[n_paths] = nk; nd_cur[n_paths] = nn; } }
              ++n_paths;
            }
          }
          if (!extended) { JB_SYNC(); JB_LANES(l) { if (l == 0) nd_cur[p] = kDeadNode; } }
        }
        JB_SYNC();
      }
    }
    return 0;
  }

  // ---------------------------------------------------------------------------- fixing_sid
  // the fix a verdict describes (jb_verdict.cuh)
  JB_W Fix verdict_fix(const Verdict& v, int64_t s0, int64_t ga) {
    Fix f;
    if (v.kind == kVNone) return f;
    f.found = true; f.repl = s.b;
    if (v.kind == kVIns) {
      fill(s.b, v.base, 1);
      f.a = s0 + k - 1; f.e = s0 + k - 1; f.rlen = 1; f.post = kPostIns;
      fix_row(f, ga, v.base, 1, 'd', '-', 1);
    } else if (v.kind == kVHomoIns) {                    // src/jasper.py:434-459: one more copy of the homopolymer's base
      fill(s.b, v.base, 1);
      f.a = s0 + k - 1; f.e = s0 + k - 1; f.rlen = 1; f.post = kPostIns;
      fix_row(f, k - 1 + s0, v.base, 1, 'd', '-', 1);
    } else if (v.kind == kVDel) {
      f.a = s0 + k - 1; f.e = s0 + k; f.rlen = 0; f.post = kPostDel;
      fix_row(f, ga - 1, '-', 1, 'i', v.bad, 1);
    } else {
      fill(s.b, v.base, 1);
      f.a = s0 + k - 1; f.e = s0 + k; f.rlen = 1; f.post = kPostSub;
      fix_row(f, ga - 1, v.base, 1, 's', v.bad, 1);
    }
    return f;
  }

  // the fix a second-stage verdict describes; false = there is none the walker may use
  JB_W bool verdict2_fix(Fix& f, int64_t s0) {
    if (!c.verdicts2) return false;
    const Verdict2* v2 = c.verdicts2 + vid;
    if (!v2->valid) return false;
    // slow_fix read bases as far back as the diploid flank (< k + 2 in front of s0): they must be untouched for this walker too
    const int64_t shift = base + olen - ipos;
    int64_t lo_x = clean_from > shift ? clean_from : shift;
    if (base - seam_margin > lo_x) lo_x = base - seam_margin;
    if (s0 - k - 2 < lo_x) return false;
    ++n_v2;
    v2_post = v2->found ? v2->post_ok : 0;
    if (!v2->found) return true;                         // decided: nothing to apply
    const int rl = v2->rlen;
    JB_SYNC();
    JB_LANES(l) { for (int q = l; q < rl; q += 32) s.b[q] = v2->repl[q]; }
    JB_SYNC();
    f.found = true; f.a = s0 + v2->a_rel; f.e = s0 + v2->e_rel; f.repl = s.b; f.rlen = rl;
    if (v2->tag == kTagBridge) {                         // a bridged patch logs its own row (the replaced bases are still in place)
      if (!bridge_row(f.a - 1, f.e, s.b, rl)) { f.found = false; v2_post = 0; }
      f.tag = 0;
    } else fix_row(f, s0 + v2->coord_rel, v2->fb, v2->fb_len, v2->tag, v2->orig, v2->orig_len);
    return true;
  }

  // k_verdicts2: the reference's dispatch on the pass-start sequence, nothing applied (the walker is set up on the whole
  // pristine contig, so every read goes to pass-start bases)
  JB_W void dry_fix(const Verdict& v, int64_t re, Verdict2* out) {
    dry = true;
    begin();
    const int64_t jq = re - (int64_t)v.d_jq;
    const int64_t s0 = jq + (int64_t)v.d_s0, ga = jq + (int64_t)v.d_ga, gb = s0 + k - 2;
    const int64_t nbad = ga - s0 > 0 ? ga - s0 : 0;
    Fix f;
    if (nbad > k) {                                      // the multi-error patch: search only, the walker logs the row
      const int64_t plen = bridge_search(ga + k - 1 - s0, solid, gb, ga);
      if (plen > 0) { f.found = true; f.a = gb + 1; f.e = ga; f.repl = bs->patch; f.rlen = plen; f.tag = kTagBridge; }
    } else f = slow_fix(s0, nbad, solid, gb, ga);
    const bool ok = !status && !(f.found && (f.rlen > kV2Repl || f.tag == 0));
    JB_SYNC();
    JB_LANES(l) {
      if (l == 0) {
        out->found = f.found ? 1 : 0;
        out->a_rel = (int16_t)(f.a - s0); out->e_rel = (int16_t)(f.e - s0); out->rlen = (int16_t)f.rlen;
        out->coord_rel = (int16_t)(f.coord - s0);
        out->fb = s.fx[0]; out->fb_len = (uint8_t)f.fb_len; out->tag = f.tag; out->orig = s.fx[64]; out->orig_len = (uint8_t)f.orig_len;
        out->post_ok = 0;
      }
      if (ok && f.found) for (int q = l; q < (int)f.rlen; q += 32) out->repl[q] = f.repl[q];
    }
    JB_SYNC();
    // the scan step behind the fix (src/jasper.py:56-100 at i = re), on the edited sequence in scratch
    if (ok && f.found && c.out && c.out_cap && apply_edit(f.a, f.e, f.repl, f.rlen)) {
      i = re;
      int steps = 0;
      for (int st = 0; st < 2; ++st) {                   // the next two scan steps, while they are the plain `i += k-1`
        // both windows of the step in one staged stretch seq[i-k : i+k): 2k bases, packed, two probes (a compact
        // restatement of generic_body for clean bases; anything else ends the evaluation)
        if (!(i < L - k + 1) || i < k) break;
        const int n2 = load(s.t, i - k, i + k);
        if (n2 != 2 * k) break;
        pack_bytes(s.pc, s.t, n2);
        if (!packed_clean(s.pc, 0, n2)) break;
        const int64_t prevk = db.count(key_canonical(packed_window(s.pc, 0, k), k));
        const int64_t occ = db.count(key_canonical(packed_window(s.pc, k, k), k));
        lookups += 2;
        if (occ < solid || (uint64_t)occ * JB_DIVISOR < (uint64_t)prevk) break;
        ++steps;
        i += k - 1;
      }
      JB_LANES(l) { if (l == 0) out->post_ok = (uint8_t)(steps == 2 ? 3 : steps); }
    }
    status = 0;
    JB_SYNC();
    release_bridge();
    JB_LANES(l) { if (l == 0) out->valid = ok ? 1 : 0; }
    JB_SYNC();
  }

  // the reference's dispatch (src/jasper.py:226-338), out of line: the hot loop only comes here for events without a verdict
  JB_WC Fix slow_fix(int64_t s0, int64_t nbad, int64_t thr, int64_t gb, int64_t ga) {
    const int64_t n64 = ga + k - 1 - s0;
    Fix f;
    if (nbad == k) {
      int n = load(s.t, s0, ga + k - 1);
      f = fix_k_case(n, thr, s0, ga);
    } else if (nbad == k - 1) {
      int n = load(s.t, s0, ga + k - 1);
      f = fix_del(n, thr, s0, ga);
      if (!f.found) f = fix_diploid(n, thr, s0, gb, ga);
      if (!f.found) f = fix_same_base_insertion(n, thr, s0);
    } else if (nbad < k - 1 && nbad > 1 && n64 >= k) {
      int n = load(s.t, s0, ga + k - 1);
      f = fix_same_base_del(n, thr, s0);
      if (!f.found) f = fix_diploid(n, thr, s0, gb, ga);
      if (!f.found) f = fix_same_base_insertion(n, thr, s0);
    } else if (nbad > k) {
      f = fix_bridge(n64, thr, gb, ga);
    }
    return f;
  }

  // ---------------------------------------------------------------------------- fixing_sid
  // returns the PostKind of an applied single-base fix (kPostNone otherwise)
  JB_W int fixing_sid(int64_t s0, int64_t nbad, int64_t thr, int64_t gb, int64_t ga, const Verdict* v) {
    Fix f;
    if (v && v->kind != kVDelimit) f = verdict_fix(*v, s0, ga);
    else if (v && verdict2_fix(f, s0)) { }
    else f = slow_fix(s0, nbad, thr, gb, ga);
    int post = kPostNone;
    if (f.found && !status) {
      if (apply_edit(f.a, f.e, f.repl, f.rlen)) {
        post = f.post;
        if (f.tag) emit_row(f.coord, s.fx, f.fb_len, f.tag, s.fx + 64, f.orig_len);
      }
    }
    release_bridge();
    return status ? kPostNone : post;
  }

  // src/jasper.py:295-329: bridge a long weak stretch with a de-Bruijn walk; the patch stays in the borrowed
  // bridging slot until fixing_sid has spliced it in
  // the search of fix_bridge: patch length (>= 1; the patch is in bs->patch and the bridging slot stays held) or 0
  JB_WC int64_t bridge_search(int64_t n64, int64_t thr, int64_t gb, int64_t ga) {
    if (gb - k + 1 < 0) return 0;                          // good_kmer_before comes out shorter than k
    int ng = load(s.t, gb - k + 1, gb + 1);
    int na = load(s.b, ga, ga + k);
    if (ng < k || na < k) return 0;
    if (n64 > 0x7fffffff) return 0;
    if (thr > solid) return 0;                             // base_extension returns None (src/jasper.py:529)
    if (!acquire_bridge()) return 0;
    const int64_t plen = base_extension((int)n64, thr);
    return (status || plen <= 0) ? 0 : plen;
  }

  // row material of a bridged patch: the patch and the original seq[gb+1:ga] go to the aux arena (aligned on the host)
  JB_WC bool bridge_row(int64_t gb, int64_t ga, const uint8_t* patch, int64_t plen) {
    const int64_t olen_old = ga - (gb + 1);
    unsigned long long off = 0, rslot = 0;
    JB_SYNC();
#if defined(__CUDA_ARCH__)
    if ((threadIdx.x & 31) == 0) { off = atomicAdd(sink.aux_used, (unsigned long long)(plen + olen_old + 8)); rslot = atomicAdd(sink.used, 1ull); }
    off = __shfl_sync(0xffffffffu, off, 0); rslot = __shfl_sync(0xffffffffu, rslot, 0);
#else
    off = *sink.aux_used; *sink.aux_used += (unsigned long long)(plen + olen_old + 8); rslot = (*sink.used)++;
#endif
    if (off + plen + olen_old + 8 > sink.aux_cap || rslot >= sink.cap) { status = JB_E_SCRATCH; return false; }
    JB_LANES(l) {
      for (int64_t q = l; q < plen; q += 32) sink.aux[off + 8 + q] = patch[q];
      for (int64_t q = l; q < olen_old; q += 32) sink.aux[off + 8 + plen + q] = get(gb + 1 + q);
      if (l == 0) {
        DevRow& r = sink.rows[rslot];
        r.coord = gb + 1; r.unit = c.unit; r.seg = c.seg; r.gen = (uint8_t)c.gen; r.seqno = seqno; r.kind = (uint8_t)kRowBridge; r.fb_len = 0; r.orig_len = 0;
        r.aux_off = off;
        const uint32_t pl = (uint32_t)plen, ol = (uint32_t)olen_old;     // byte-wise: the arena is not aligned
        for (int q = 0; q < 4; ++q) { sink.aux[off + q] = (uint8_t)(pl >> (8 * q)); sink.aux[off + 4 + q] = (uint8_t)(ol >> (8 * q)); }
      }
    }
    JB_SYNC();
    seqno += 2;                                           // at most two csv rows come out of one patch
    return true;
  }

  JB_WC Fix fix_bridge(int64_t n64, int64_t thr, int64_t gb, int64_t ga) {
    Fix f;
    const int64_t plen = bridge_search(n64, thr, gb, ga);
    if (plen <= 0) return f;
    if (!bridge_row(gb, ga, bs->patch, plen)) return f;
    f.found = true; f.a = gb + 1; f.e = ga; f.repl = bs->patch; f.rlen = plen; f.tag = 0;
    return f;
  }

  // src/jasper.py:150-214 through the window cache / point queries (edited territory, rolling threshold, contig ends, N);
  // false = the reference returns before it reaches the fix
  JB_WC bool generic_delimit(int64_t rolling_thre, int64_t thre, int64_t& gb, int64_t& ga, int64_t& s0) {
    int64_t j = i - 1;
    uint32_t occ = Cw(j, -1);
    while ((int64_t)occ < thre && j >= 0) { --j; occ = Cw(j, -1); }
    gb = j + k - 1;
    uint64_t prev = occ;
    uint32_t cur = Cw(i, +1);
    if (j == -1) gb = -1;
    if (rolling_thre == 0) {
      while ((int64_t)cur < thre && i < L - k + 1) { ++i; cur = Cw(i, +1); }
    } else {
      while ((int64_t)cur < thre && i < L - k + 1) {
        if (i - j > k) { i = i + 1; return false; }
        ++i; cur = Cw(i, +1);
      }
    }
    ga = i;
    if (2 * (int64_t)Cw(gb - k + 2, +1) < solid && 2 * (int64_t)Cw(gb - k + 3, +1) < solid) {
      // "too low" -- flag only
    } else if (rolling_thre == 0) {
      while (2 * (uint64_t)Cw(gb - k + 2, +1) >= prev && gb - k + 1 < ga) {
        if (gb == -1) break;
        if (2 * (int64_t)prev >= thre && 2 * (int64_t)Cw(gb - k + 2, +1) < thre && 2 * (int64_t)Cw(gb - k + 3, +1) < thre) break;
        prev = Cw(gb - k + 2, +1);
        ++gb;
      }
      if (gb >= L - 1) { brk = true; return false; }
    }
    s0 = gb - k + 2 > 0 ? gb - k + 2 : 0;
    if (s0 + k + k >= L) { i = s0 + k + k; brk = true; return false; }
    s.px[0] = s0 + 1; s.px[1] = s0 + k - 2; s.px[2] = s0 + k - 1; s.px[3] = s0 + k;
    JB_SYNC();
    Cpts(4);
    if ((int64_t)s.pr[0] < thre && (int64_t)s.pr[1] < thre && (int64_t)s.pr[2] < thre && (int64_t)s.pr[3] >= thre) ga = s0 + k;
    return true;
  }

  // ---------------------------------------------------------------------------- verdicts (untouched territory)
  // The isolated error was decided before the walk (jb_verdict.cuh).  While the walker stands in bases no edit has
  // touched it looks the weak run up by its end and takes bounds, fix and the following scan step from the verdict;
  // false leaves the walker's state untouched and the generic code runs.
  JB_W bool consume_verdict(int64_t& gb, int64_t& ga, int64_t& s0, Verdict& v) {
    if (!c.vmap) return false;
    const int64_t shift = base + olen - ipos;           // new = old + shift for everything >= clean_from
    int64_t lo_x = clean_from > shift ? clean_from : shift;
    if (base - seam_margin > lo_x) lo_x = base - seam_margin;      // (further back the previous segment's walker may have edited)
    if (i < lo_x) return false;
    const int64_t p = i - shift;
    if (p > c.in_len) return false;
    int64_t re = -1;                                    // first position >= p whose weak bit is clear
    {
      int64_t w = p >> 5;
      uint32_t m = ~c.weak[w] & (0xFFFFFFFFu << (p & 31));
      for (int t = 0; t < 3; ++t) {
        if (m) {
#if defined(__CUDA_ARCH__)
          re = (w << 5) + __ffs((int)m) - 1;
#else
          re = (w << 5) + __builtin_ffs((int)m) - 1;
#endif
          break;
        }
        ++w;
        if ((w << 5) > c.in_len) break;
        m = ~c.weak[w];
      }
    }
    if (re < 0) return false;
    const uint32_t id = c.vmap[(c.ev_bit_base + (uint64_t)re) >> 5];
    if (id == 0u || id == kVmapAmbiguous) return false;
    v = c.verdicts[id - 1];
    vid = id - 1;
    if (v.re != (uint32_t)re) return false;
    const int64_t jq = re - (int64_t)v.d_jq;
    if (jq + shift < lo_x) return false;                // the verdict read counts this walker's edits have invalidated
    i = re + shift;
    s0 = jq + (int64_t)v.d_s0 + shift; ga = jq + (int64_t)v.d_ga + shift; gb = s0 + k - 2;
    ++n_fast;
    return true;
  }

  // ---------------------------------------------------------------------------- handle_bad_kmers
  JB_WN void handle_bad_kmers(int64_t rolling_thre) {
    ++events;
    const int64_t thre = rolling_thre > 0 ? rolling_thre : solid;
    int64_t gb = 0, ga = 0, s0 = 0;
    Verdict v;
    const bool fast = rolling_thre == 0 && consume_verdict(gb, ga, s0, v);
    if (!fast && !generic_delimit(rolling_thre, thre, gb, ga, s0)) return;
    const int64_t nbad = ga - s0 > 0 ? ga - s0 : 0;
    wrong += nbad;
    if (!fix || gb < 0) return;
    v2_post = 0;
    const int done = fixing_sid(s0, nbad, thre, gb, ga, fast ? &v : nullptr);
    if (!fast || status || brk) return;
    const int bits = done ? v.post_ok : v2_post;
    if (!(bits & 1)) return;
    // ---- the scan step behind a single-base fix (src/jasper.py:56-100 at the new i) was evaluated with the verdict
    if (c.next_anchor >= 0 && i - (L - c.in_len) >= c.next_anchor) return;       // the loop re-checks the seam first
    i += k - 1;
    ++n_post;
    if (!(bits & 2)) return;
    if (c.next_anchor >= 0 && i - (L - c.in_len) >= c.next_anchor) return;
    i += k - 1;                                                                     // the step after that was evaluated too
    ++n_post;
  }

  // ---------------------------------------------------------------------------- scan loop
  // Fast hop: while the walker is in untouched territory, 32 strided event bits decide in one step
  // how far the plain `i += k-1` branch carries it.
  JB_WN bool fast_hop() {
    if (!c.ev || i - k < clean_from) return false;
    const int64_t shift = base + olen - ipos;          // new = old + shift for everything >= clean_from
    const int64_t p = i - shift;
    const int64_t stride = k - 1;
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31);
    int l;
    {
      const int64_t q = p + (int64_t)lane * stride;
      const bool f = q > c.in_len || ((c.ev[q >> 5] >> (q & 31)) & 1u);
      const unsigned m = __ballot_sync(0xffffffffu, f);
      l = m ? __ffs(m) - 1 : 32;
    }
#else
    JB_LANES(l) {
      int64_t q = p + (int64_t)l * stride;
      uint32_t f = 1;
      if (q <= c.in_len) f = (c.ev[q >> 5] >> (q & 31)) & 1u;
      s.flag[l] = f;
    }
    int l = 0;
    while (l < 32 && !s.flag[l]) ++l;
#endif
    i += (int64_t)l * stride;
    if (l < 32) return false;                           // an event (or the end) at the new i
    if (c.ev_l1) {
      // 32 clear strides: consult the level-1 summary to leap over event-free 1024-position blocks
      const int64_t p2 = i - shift;
      const uint64_t blk = (c.ev_bit_base + (uint64_t)p2) >> 10;
      const uint64_t last_blk = (c.ev_bit_base + (uint64_t)c.in_len) >> 10;
#if defined(__CUDA_ARCH__)
      int m;
      {
        const uint64_t bq = blk + (uint64_t)lane;
        const bool f = bq > last_blk || ((c.ev_l1[bq >> 5] >> (bq & 31)) & 1u);
        const unsigned mm = __ballot_sync(0xffffffffu, f);
        m = mm ? __ffs(mm) - 1 : 32;
      }
#else
      JB_LANES(l2) {
        uint64_t bq = blk + (uint64_t)l2;
        s.flag[l2] = bq > last_blk ? 1u : ((c.ev_l1[bq >> 5] >> (bq & 31)) & 1u);
      }
      int m = 0;
      while (m < 32 && !s.flag[m]) ++m;
#endif
      if (m == 32 && c.ev_l2) {
        // 32 empty blocks: the level-2 summary leaps over event-free level-1 words (1 M positions per step); the blocks
        // looked at above run up to (blk + 31), so the search starts at the level-1 word that holds blk + 32
        const uint64_t w0 = (blk + 32) >> 5, last_w = last_blk >> 5;
#if defined(__CUDA_ARCH__)
        int m2;
        {
          const uint64_t wq = w0 + (uint64_t)lane;
          const bool f = wq > last_w || ((c.ev_l2[wq >> 5] >> (wq & 31)) & 1u);
          const unsigned mm = __ballot_sync(0xffffffffu, f);
          m2 = mm ? __ffs(mm) - 1 : 32;
        }
#else
        JB_LANES(l2) {
          uint64_t wq = w0 + (uint64_t)l2;
          s.flag[l2] = wq > last_w ? 1u : ((c.ev_l2[wq >> 5] >> (wq & 31)) & 1u);
        }
        int m2 = 0;
        while (m2 < 32 && !s.flag[m2]) ++m2;
#endif
        // blocks [blk, (w0 + m2) << 5) are empty: [blk, blk + 32) by level 1, whole level-1 words w0 .. w0 + m2 - 1 by level 2,
        // and the blocks of word (blk + 32) >> 5 below blk + 32 are among the 32 seen above
        const uint64_t nb = (w0 + (uint64_t)m2) << 5;
        if (nb > blk + 32) m = (int)(nb - blk);                     // at most 32 + 31 + 32 * 32 blocks
      }
      if (m >= 2) {
        const int64_t T = (int64_t)(((blk + (uint64_t)m) << 10) - c.ev_bit_base);   // first position of a non-empty block
        if (T - 1 > p2) i += ((T - 1 - p2) / stride) * stride;
      }
    }
    return true;                                        // all clear: keep hopping
  }

  // the loop body's window checks and its two counts through the generic machinery (src/jasper.py:56-71); true = the
  // window held an invalid byte and i moved on
  JB_WC bool generic_body(int64_t& occ, int64_t& prevk) {
    int n = load(s.t, i, i + k);                      // n == k here
    int firstN = -1, firstn = -1, bad = 0;
    for (int q = 0; q < n; ++q) {
      uint8_t c = s.t[q];
      if (c == 'N' && firstN < 0) firstN = q;
      if (c == 'n' && firstn < 0) firstn = q;
      if (!base_valid(c)) bad = 1;
    }
    if (firstN >= 0) {
      i += firstN + 1;
      // a run of N: every following window starts with 'N' and advances by one; jump the run
      while (i < L - k + 1 && get(i) == 'N') ++i;
      return true;
    }
    if (firstn >= 0) { i += firstn + 1; return true; }
    if (bad) { i += 1; return true; }
    s.px[0] = i; s.px[1] = i - k > 0 ? i - k : 0;
    JB_SYNC();
    Cpts(i > 0 ? 2 : 1);
    occ = s.pr[0];
    prevk = i > 0 ? (int64_t)s.pr[1] : 0;
    return false;
  }

  // src/jasper.py:80-100: the count dropped below 1/50 of the one k bases back; the threshold handle_bad_kmers gets, or -1
  JB_WC int64_t rolling_threshold(int64_t occ) {
    int64_t ind = i - k > 0 ? i - k : 0;
    int num = 0;
    while (ind < i) { ind += step; s.px[num] = ind; ++num; }
    JB_SYNC();
    Cpts(num);
    uint64_t sum = 0;
    for (int q = 0; q < num; ++q) sum += s.pr[q];
    const double mean = (double)sum / (double)num;
    const int64_t rt = (int64_t)rint(mean / (double)JB_DIVISOR);
    return occ < rt ? (int64_t)rint(mean / 2.0) : -1;
  }

  // segment set-up and result write-back run once per segment: out of line, so the event loop between them stays small
  JB_WC void begin() {
#if defined(__CUDA_ARCH__)
    t_begin = clock64();
#endif
    c = u;
    base = c.base; seam_margin = seam_read_margin(k); keep_margin = c.keep_margin; wstart = 0; far_back = false;
    i = base; olen = 0; ipos = base; L = c.in_len; clean_from = 0; wrong = 0; status = 0; brk = false; seam_bad = false; vflags = 0;
    publish();
    {
      // C('') = count of poly-A: the bulk scan's entry for the empty tail window, else one probe
      uint32_t ec;
      if (c.cnt) ec = c.cnt[c.in_len]; else { ec = count_str(db, s.c, 0, 0, k, mask); ++lookups; }
      JB_SYNC();
      JB_LANES(l) { if (l == 0) { s.empty_count = ec; s.probes = 0; } }
      JB_SYNC();
    }
  }

  JB_WC void finish(bool reached) {
    seam_bad = seam_bad || (vflags & kFlagSeam); far_back = far_back || (vflags & kFlagFar);
#if defined(__CUDA_ARCH__)
    seam_bad = __any_sync(0xffffffffu, seam_bad);      // lanes read different bases: make the verdicts warp-uniform
    far_back = __any_sync(0xffffffffu, far_back);
#endif
    if (!status && seam_bad) status = kSeamViolation;
    if (!status && far_back) status = kFarBack;
    if (!status && c.next_anchor >= 0 && !reached) status = kSeamViolation;   // scan ended (or broke) before the seam
    const int64_t end_old = c.next_anchor >= 0 ? c.next_anchor : c.in_len;
    if (!status && ipos > end_old) status = kSeamViolation;
    // new bases of old [base, end_old) = closed jobs ++ out[wstart:olen) ++ in[ipos:end_old)
    if (fix && !status) {
      emit_job(c.out + wstart, wstart, olen - wstart);
      emit_job(c.in + ipos, olen, end_old - ipos);
    }
    u.end_old = end_old;
    u.new_len = status ? 0 : olen + (end_old - ipos);
    u.wrong = wrong;
    u.status = status;
    u.n_rows = seqno;
    u.n_fast = n_fast; u.n_post = n_post; u.n_v2 = n_v2;
    u.lookups = lookups + s.probes;
    u.events = events;
#if defined(__CUDA_ARCH__)
    u.cycles = (unsigned long long)(clock64() - t_begin);
#else
    u.cycles = 0;
#endif
  }

  JB_W void run() {
    begin();
    bool reached = false;
    while (i < L - k + 1 && !brk && !status) {
#if defined(__CUDA_ARCH__)
      if (__any_sync(0xffffffffu, vflags != 0)) break;            // the result is void already: stop early
#else
      if (vflags) break;
#endif
      if (c.next_anchor >= 0) {
        const int64_t i_old = i - (L - c.in_len);
        if (i_old >= c.next_anchor) {
          // the next segment starts here: we must arrive clean and land inside its anchor run
          // (everything from seam_margin before the anchor on is untouched, so the next walker reads pass-start bases)
          const int64_t a_new = c.next_anchor + (L - c.in_len);
          if (clean_from > a_new - seam_margin || i_old > c.next_anchor + k - 2) status = kSeamViolation;
          reached = true;
          break;
        }
      }
      if (fast_hop()) continue;
      if (!(i < L - k + 1)) break;
      if (c.next_anchor >= 0 && i - (L - c.in_len) >= c.next_anchor) continue;   // re-check the seam first
      // ---- the reference's loop body at position i (src/jasper.py:56-100)
      int64_t occ = 0, prevk = 0;
      bool direct = false;
      if (c.okb && c.cnt && i - k >= clean_from && i - k >= base - seam_margin) {
        // untouched territory: validity and both counts come straight from the bulk scan
        const int64_t p = i - (base + olen - ipos);
        if (p >= k && p + k <= c.in_len && c.okb[p]) { occ = c.cnt[p]; prevk = c.cnt[p - k]; direct = true; }
      }
      if (!direct && generic_body(occ, prevk)) continue;
      int64_t handle = -1;                              // >= 0: handle_bad_kmers(handle) -- one call site keeps the image small
      if (occ < solid) handle = 0;
      else if (i > 0 && (uint64_t)occ * JB_DIVISOR < (uint64_t)prevk) handle = rolling_threshold(occ);
      if (handle >= 0) handle_bad_kmers(handle);
      else i += k - 1;
    }
    finish(reached);
  }
};

}  // namespace jb

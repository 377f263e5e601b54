// jb_jobs.cuh -- copy jobs: how a fixing pass describes the new bases of a segment (written by the walker,
// consumed by k_concat and by the incremental recount).
#pragma once
#include <stdint.h>

namespace jb {

// The new bases of a segment are described as copy jobs: materialised windows around edits (src = the
// segment's staging region) and untouched stretches (src = the pass-start contig).  `rel` = offset inside
// the segment's new bases; k_concat adds the segment's offset in the new contig.
//
// What the incremental recount needs beyond that (r2ab).  A materialised window is pass-start bases with a few edits in it:
// `old` / `old_len` name the pass-start stretch it replaces, `lead` / `trail` (info) how many bases of an untouched copy job
// lie directly in front of it and behind it in BOTH sequences (clamped to k-1).  k_recount_probe lines the window up with
// the stretch it replaces (recount_walk below): a window of the new contig that lies inside a run of bases the two have in
// common consists of the same k bases as a pass-start window and inherits its count; only the windows that overlap a
// difference are probed (k per single-base edit instead of the ~2k+32+k the whole materialised window holds).  An untouched
// job whose successor took its last k-1 windows over this way carries kJobTailOwned.
struct CopyJob { const uint8_t* src; int64_t rel; int64_t len; uint32_t slot; uint32_t info; const uint8_t* old; int64_t old_len; };
struct JobSink { CopyJob* jobs; unsigned long long cap; unsigned long long* used; };
constexpr uint32_t kJobTailOwned = 1u;    // untouched job: the next job of the segment accounts for its last min(len, k-1) windows
constexpr uint32_t kJobHasOld = 2u;       // materialised job: old / old_len / lead / trail are set
#if defined(__CUDACC__)
#define JB_JOBS_HD __host__ __device__ inline
#define JB_JOBS_D __device__ inline
#else
#define JB_JOBS_HD inline
#define JB_JOBS_D inline
#endif
JB_JOBS_HD uint32_t job_info(uint32_t flags, int64_t lead, int64_t trail) { return flags | ((uint32_t)lead << 8) | ((uint32_t)trail << 16); }
JB_JOBS_HD int job_lead(uint32_t info) { return (int)((info >> 8) & 0xFFu); }
JB_JOBS_HD int job_trail(uint32_t info) { return (int)((info >> 16) & 0xFFu); }

// Which windows of a materialised job keep a pass-start count.  The job's bases x[0:n) replace the pass-start stretch
// y[0:m); the job accounts for the windows that start at [-lead, n) (relative to x[0]).  The walk lines x up with y greedily:
// extend the common run, and at a difference try "one base substituted / inserted / deleted" and go on with the hypothesis
// that matches furthest; when none does (a multi-base patch), whatever x and y have in common at their ends is the last run.
// A window that lies inside one run -- the first run reaches `lead` bases back, a run that ends both x and y reaches `trail`
// bases on -- is inherited from the pass-start window at the same place in the run; everything else is probed.
// Exactness does not depend on how well the walk aligns: a run is only what `match` has compared and found equal.
//   match.fwd(xn, yo, lim)   length <= lim (<= kMatchStep) of the common prefix of x[xn:) and y[yo:)
//   match.bwd(xe, ye, lim)   length <= lim (<= kMatchStep) of the common suffix of x[:xe) and y[:ye)
//   match.resync(xn, yo, ls, li, ld, &hs, &hi, &hd)   the three hypotheses' runs behind a difference at (xn, yo), each <= its limit
// next() hands out the stretches of windows one at a time (the caller has ONE place that probes and one that copies):
//   kRangeInherit  windows [a, b) <- pass-start windows that start at y[yo + (q - a)]
//   kRangeProbe    windows [a, b) are looked up
// The comparisons a job may spend are bounded (kMatchBudget): a whole contig materialised by a re-walk is mostly probed.
constexpr int kMatchStep = 128;
constexpr int kMatchBudget = 4096;
constexpr int kResyncSpan = 64;       // how far a hypothesis is followed
constexpr int kResyncMin = 12;        // ... and how far it must hold to be believed
enum RangeKind : int { kRangeProbe = 1, kRangeInherit = 2 };
struct RecountRange { int kind; int64_t a, b, yo; };
struct RecountWalk {
  int64_t plain_a, plain_b;               // a plain job: its windows [plain_a, plain_b), all probed
  int n, m, done, pn, po, budget, qa, qb, qyo;
  int lead, trail, k, mode;               // mode: 0 extend the run at (pn, po), 2 a difference at (pn, po), 1 common end, 3 over
  bool first, queued;

  // a job without a pass-start stretch (or the test switch that ignores it): windows [from, n) are probed
  JB_JOBS_HD void init_plain(int64_t from, int64_t n_) { plain_a = from; plain_b = n_; n = m = done = pn = po = budget = 0; lead = trail = 0; k = 1; mode = 3; first = false; queued = false; qa = qb = qyo = 0; }
  // (job-relative positions in 32 bits: a job of 2^30 bases or more -- a whole chromosome materialised by a re-walk -- is probed)
  JB_JOBS_HD void init(int64_t n_, int64_t m_, int lead_, int trail_, int k_) {
    if (n_ >= (1ll << 30) || m_ >= (1ll << 30)) { init_plain(-(int64_t)lead_, n_); return; }
    plain_a = plain_b = 0; n = (int)n_; m = (int)m_; lead = lead_; trail = trail_; k = k_;
    done = -lead_; pn = po = 0; budget = kMatchBudget; mode = 0; first = true; queued = false; qa = qb = qyo = 0;
  }
  // a run x[s:s+len) == y[o:o+len): the windows inside it are inherited, the ones in front of it (from `done` on) probed
  JB_JOBS_HD bool settle(int s, int o, int len, bool ends, RecountRange* r) {
    const int se = first ? s - lead : s, oe = first ? o - lead : o;
    const int ee = s + len + (ends ? trail : 0);
    const int w0 = done > se ? done : se, w1 = ee - k + 1 < n ? ee - k + 1 : n;
    first = false;
    if (w1 <= w0) return false;
    if (w0 > done) {
      r->kind = kRangeProbe; r->a = done; r->b = w0; r->yo = 0;
      queued = true; qa = w0; qb = w1; qyo = oe + (w0 - se);
    } else {
      r->kind = kRangeInherit; r->a = w0; r->b = w1; r->yo = oe + (w0 - se);
    }
    done = w1;
    return true;
  }
  template <class Match>
  JB_JOBS_D bool next(const Match& match, RecountRange* r) {
    if (queued) { queued = false; r->kind = kRangeInherit; r->a = qa; r->b = qb; r->yo = qyo; return true; }
    while (mode != 3) {
      if (mode == 0) {
        const int lim = n - pn < m - po ? n - pn : m - po;
        int len = 0;
        while (len < lim && budget > 0) {
          const int step = lim - len < kMatchStep ? lim - len : kMatchStep;
          const int got = match.fwd(pn + len, po + len, step);
          budget -= step;
          len += got;
          if (got < step) break;
        }
        const int s = pn, o = po;
        pn += len; po += len;
        mode = (budget <= 0 || pn >= n || po >= m) ? 3 : 2;
        if (settle(s, o, len, pn == n && po == m, r)) return true;
      } else if (mode == 2) {
        const int rn = n - pn, ro = m - po;                      // both >= 1: x[pn] != y[po]
        int ls = rn - 1 < ro - 1 ? rn - 1 : ro - 1, li = rn - 1 < ro ? rn - 1 : ro, ld = rn < ro - 1 ? rn : ro - 1;
        if (ls > kResyncSpan) ls = kResyncSpan;
        if (li > kResyncSpan) li = kResyncSpan;
        if (ld > kResyncSpan) ld = kResyncSpan;
        int hs = 0, hi = 0, hd = 0;
        if (ls >= kResyncMin || li >= kResyncMin || ld >= kResyncMin) {       // (an edit at the very end of the window: nothing to follow)
          match.resync(pn, po, ls, li, ld, &hs, &hi, &hd);
          budget -= 3 * kResyncSpan;
        }
        mode = 0;
        if (hs >= kResyncMin && hs >= hi && hs >= hd) { pn += 1; po += 1; }
        else if (hi >= kResyncMin && hi >= hd) pn += 1;
        else if (hd >= kResyncMin) po += 1;
        else mode = 1;                                               // no single-base explanation
      } else {
        const int rn = n - pn, ro = m - po;
        const int blim = rn < ro ? rn : (rn == ro ? ro - 1 : ro);   // (equally long rests: x[pn] != y[po] would be part of the common end)
        int S = 0;
        while (S < blim && budget > 0) {
          const int step = blim - S < kMatchStep ? blim - S : kMatchStep;
          const int got = match.bwd(n - S, m - S, step);
          budget -= step;
          S += got;
          if (got < step) break;
        }
        mode = 3;
        if (S > 0 && settle(n - S, m - S, S, true, r)) return true;
      }
    }
    if (done < n) { r->kind = kRangeProbe; r->a = done; r->b = n; r->yo = 0; done = n; return true; }
    if (plain_a < plain_b) { r->kind = kRangeProbe; r->a = plain_a; r->b = plain_b; r->yo = 0; plain_a = plain_b; return true; }
    return false;
  }
};
#if defined(__CUDACC__)
// How the kernels that consume the job list (k_concat, k_recount_copy, k_recount_probe) share it out.  They are launched
// on a fixed 1-D grid BEFORE the host knows how many jobs the walkers wrote (no host round trip between k_walk and them),
// so the split is decided on the device from the job count: many short jobs (a fixing pass over a dirty assembly leaves
// ~3 per edit, a few hundred positions each) -> teams of kJobTeam threads per job; few long ones -> the job list is dealt
// to rows of `gx` CTAs, each row sharing one job at a time.
constexpr int kJobTeam = 64;
constexpr unsigned long long kJobsForTeams = 16384;
struct JobSplit { int64_t tid, nth; unsigned long long j0, jstep; };
__device__ __forceinline__ JobSplit job_split(unsigned long long n_jobs, bool force_rows, unsigned max_gx, int team = kJobTeam) {
  JobSplit r;
  if (n_jobs >= kJobsForTeams && !force_rows) {
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    r.tid = (int64_t)(g % (unsigned)team); r.nth = team;
    r.j0 = g / (unsigned)team; r.jstep = (unsigned long long)gridDim.x * blockDim.x / (unsigned)team;
  } else {
    unsigned long long want = 65536 / (n_jobs ? n_jobs : 1);
    unsigned gx = (unsigned)(want < 4 ? 4 : (want > 128 ? 128 : want));
    if (gx > max_gx) gx = max_gx;
    if (gx > gridDim.x) gx = gridDim.x;
    const unsigned rows = gridDim.x / gx, row = blockIdx.x / gx;
    r.tid = (int64_t)(blockIdx.x % gx) * blockDim.x + threadIdx.x; r.nth = (int64_t)gx * blockDim.x;
    r.j0 = row < rows ? row : ~0ull; r.jstep = rows;              // (CTAs behind the last whole row have nothing to do)
  }
  return r;
}
#endif

}  // namespace jb

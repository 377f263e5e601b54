// walker_hostsim.cpp -- TEST-ONLY differential build of the walker (jasper_b200/csrc/jb_walker.cuh)
// for the CPU: the same source the GPU runs, with a warp emulated by a 32-iteration lane loop and
// table probes answered through a callback (the oracle's lookup).  It lets `-m "not gpu"` tests
// compare the device state machine against oracle/jasper_oracle.py without a GPU.
// It is never linked into libjasper_b200.so and nothing in jasper_b200/ loads it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../jasper_b200/csrc/jb_plan.hpp"

using namespace jb;

typedef uint64_t (*lookup_fn)(const void* db, const uint64_t key[2]);

struct HostDB {
  lookup_fn fn;
  const void* db;
  uint32_t count(const Key& c) const { uint64_t kk[2] = {c.lo, c.hi}; return (uint32_t)fn(db, kk); }
};

struct SimResult {
  int status = 0;
  std::vector<PassOut> passes;
  std::vector<std::string> seqs;
  std::string text;     // rows as "ite\tseq\tcoord\tmutation\toriginal\n"
  unsigned long long lookups = 0, events = 0, segments = 0, reruns = 0, far_backs = 0, fast = 0, post = 0, v2 = 0, pool_growths = 0;
};

// event / weak bitmasks exactly as k_scan<.,true> defines them (jb_scan.cuh)
static void scan_bits(const HostDB& db, const std::string& seq, int k, uint32_t solid, std::vector<uint32_t>* ev, std::vector<uint32_t>* weak,
                      std::vector<uint32_t>* cnt_out, std::vector<uint8_t>* ok_out) {
  const int64_t L = (int64_t)seq.size();
  std::vector<uint32_t>& cnt = *cnt_out;
  cnt.assign(L + 1, 0);
  std::vector<uint8_t>& ok = *ok_out;
  ok.assign(L + 1, 0);
  Key mask = key_mask(k);
  for (int64_t p = 0; p <= L; ++p) {
    Key m = key_from_chars((const uint8_t*)seq.data() + p, L - p, k, mask);
    cnt[p] = db.count(key_canonical(m, k));
    bool v = p + k <= L;
    for (int q = 0; v && q < k; ++q) v = base_valid((uint8_t)seq[p + q]);
    ok[p] = v;
  }
  ev->assign((L + 1 + 31) / 32 + 1, 0);
  weak->assign((L + 1 + 31) / 32 + 1, 0);
  for (int64_t p = 0; p <= L; ++p) {
    bool e = !ok[p] || p < k || cnt[p] < solid;
    if (!e) e = (unsigned long long)cnt[p] * JB_DIVISOR < (unsigned long long)cnt[p - k];
    if (e) (*ev)[p >> 5] |= 1u << (p & 31);
    if (ok[p] && cnt[p] < solid) (*weak)[p >> 5] |= 1u << (p & 31);
  }
}

// use_events: 0 = the walker evaluates every position itself, 1 = event bitmask + 1-per-1024 summary,
// >= 64 = additionally split every contig into segments at the first anchor of every `use_events` positions
extern "C" SimResult* hostsim_polish(lookup_fn fn, const void* dbh, int k, uint32_t solid, int passes, int fix,
                                     int n_seqs, const char* const* seqs, const uint64_t* lens, int use_events) {
  HostDB db{fn, dbh};
  auto* res = new SimResult();
  const size_t n = (size_t)n_seqs;
  std::vector<std::string> cur(n);
  for (size_t i = 0; i < n; ++i) cur[i].assign(seqs[i], lens[i]);
  std::vector<DevRow> rows(1 << 16);
  std::vector<uint8_t> aux(1 << 22);
  static BridgeScratch bs;                     // large: keep off the stack
  std::vector<uint32_t> trie(1u << 20);
  std::vector<uint8_t> patch(1u << 18);
  bs.trie = trie.data(); bs.trie_cap = (uint32_t)trie.size(); bs.patch = patch.data(); bs.patch_cap = (uint32_t)patch.size();
  BridgePool pool{&bs, 1, nullptr};
  WarpScratch ws;
  const int64_t seg_block = use_events >= 64 ? use_events : 0;
  int64_t margin = anchor_margin(k);
  if (const char* e = getenv("JB_ANCHOR_MARGIN")) margin = atol(e);   // tests shrink it to provoke seam re-runs
  int keep_margin = default_keep_margin(k);
  if (const char* e = getenv("JB_KEEP_MARGIN")) keep_margin = atoi(e);
  std::vector<int64_t> caps(n);
  for (size_t i = 0; i < n; ++i) { int64_t L = (int64_t)cur[i].size(); caps[i] = L + L / 8 + 4096; }
  for (int ite = 0; ite <= passes; ++ite) {
    const bool do_fix = fix && ite != passes;
    PassOut po;
    unsigned long long used = 0, aux_used = 0;
    RowSink sink{rows.data(), rows.size(), &used, aux.data(), aux.size(), &aux_used};
    unsigned long long jobs_used = 0;
    std::vector<CopyJob> jobv(1 << 18);
    JobSink jobs{jobv.data(), jobv.size(), &jobs_used};
    std::vector<int64_t> len(n);
    std::vector<std::vector<uint32_t>> ev(n), weak(n), l1(n), l2(n), cnts(n);
    std::vector<std::vector<uint8_t>> stage(n), okbs(n);
    for (size_t i = 0; i < n; ++i) {
      len[i] = (int64_t)cur[i].size();
      po.total += len[i] - k + 1;
      stage[i].resize((size_t)caps[i] + 64);
      if (!use_events) continue;
      scan_bits(db, cur[i], k, solid, &ev[i], &weak[i], &cnts[i], &okbs[i]);
      // level-1 summary over this contig's own bit positions (ev_bit_base = 0)
      l1[i].assign((size_t)(len[i] + 1) / 32768 + 2, 0);
      for (int64_t p = 0; p <= len[i]; ++p)
        if (bit_at(ev[i].data(), p)) l1[i][p >> 15] |= 1u << ((p >> 10) & 31);
      l2[i].assign(l1[i].size() / 32 + 2, 0);
      for (size_t w = 0; w < l1[i].size(); ++w) if (l1[i][w]) l2[i][w >> 5] |= 1u << (w & 31);
    }
    // verdicts, the way k_verdicts builds them: one per weak run, published under the run's end
    std::vector<std::vector<uint32_t>> vmaps(n);
    std::vector<std::vector<Verdict>> verdicts(n);
    std::vector<std::vector<Verdict2>> verdicts2(n);
    const bool use_verdicts2 = do_fix && !getenv("JB_SIM_NO_VERDICTS2");
    const bool use_verdicts = use_events && !getenv("JB_SIM_NO_OKB");
    for (size_t i = 0; i < n && use_verdicts; ++i) {
      vmaps[i].assign((size_t)(len[i] + 1) / 32 + 4, 0);
      for (int64_t re = 1; re <= len[i]; ++re) {
        if (bit_at(weak[i].data(), re) || !bit_at(weak[i].data(), re - 1)) continue;
        Verdict v; int64_t re2 = -1;
        if (!make_verdict(db, ws, (const uint8_t*)cur[i].data(), cnts[i].data(), okbs[i].data(), len[i], k, sampling_step(k), (int64_t)solid, re - 1, &v, &re2)) continue;
        if (re2 != re) continue;
        verdicts[i].push_back(v);
        uint32_t& cell = vmaps[i][(size_t)re >> 5];
        cell = cell == 0 ? (uint32_t)verdicts[i].size() : kVmapAmbiguous;
      }
      // second stage, the way k_verdicts2 does it: the walker's own dispatch on the pristine contig
      verdicts2[i].resize(verdicts[i].size());
      for (size_t q = 0; q < verdicts[i].size() && use_verdicts2; ++q) {
        const Verdict& v = verdicts[i][q];
        Verdict2& v2 = verdicts2[i][q];
        memset(&v2, 0, sizeof v2);
        if (v.kind != kVDelimit) continue;
        UnitIO u;
        memset(&u, 0, sizeof u);
        u.in = (const uint8_t*)cur[i].data(); u.in_len = len[i];
        u.cnt = cnts[i].data(); u.okb = okbs[i].data();
        u.next_anchor = -1; u.keep_margin = keep_margin;
        const int64_t s0v = (int64_t)v.re - (int64_t)v.d_jq + (int64_t)v.d_s0;
        static uint8_t dry_out[1024];
        u.base = s0v > 64 ? s0v - 64 : 0; u.out = dry_out; u.out_cap = sizeof dry_out;
        Walker<HostDB> w(db, k, sampling_step(k), solid, true, u, ws, pool, sink, jobs);
        w.dry_fix(v, (int64_t)v.re, &v2);
        if (getenv("JB_SIM_TRACE")) {
          const int nb = (int)v.d_ga - (int)v.d_s0;
          fprintf(stderr, "[sim] v2 pass %d nbad %d (k %d) valid %d found %d a_rel %d e_rel %d rlen %d tag %c post %d\n", ite, nb, k, v2.valid, v2.found, v2.a_rel, v2.e_rel, v2.rlen,
                  v2.tag == 0xFF ? 'B' : (v2.tag ? v2.tag : '-'), v2.post_ok);
        }
      }
    }
    std::vector<BlockRef> blocks;
    if (seg_block) blocks = make_blocks(len, seg_block);
    std::vector<long long> anchors(blocks.size(), -1);
    for (size_t q = 0; q < blocks.size(); ++q) {
      const size_t u = blocks[q].unit;
      const int64_t lo = (int64_t)blocks[q].start, hi = std::min<int64_t>(lo + seg_block, len[u]);
      anchors[q] = first_anchor(ev[u].data(), weak[u].data(), lo, hi, k, len[u], margin);
    }
    if (getenv("JB_SIM_TRACE") && ite > 0 && use_events) {
      for (size_t u = 0; u < n; ++u)
        for (int64_t a = 1; a <= len[u]; ++a) {
          if (!bit_at(weak[u].data(), a) || bit_at(weak[u].data(), a - 1)) continue;
          int64_t run = 0; while (a + run <= len[u] && bit_at(weak[u].data(), a + run)) ++run;
          int64_t evb = -1; for (int64_t p2 = a - 1; p2 >= 0 && p2 >= a - margin; --p2) if (bit_at(ev[u].data(), p2)) { evb = a - p2; break; }
          fprintf(stderr, "[sim]   pass %d weak run at %lld len %lld, nearest event before: %lld back, anchor=%d\n", ite, (long long)a, (long long)run, (long long)evb,
                  (int)is_anchor(ev[u].data(), weak[u].data(), a, k, len[u], margin));
        }
    }
    std::vector<SegPlan> plan = plan_segments(len, caps, blocks, anchors.data());
    auto make_seg = [&](const SegPlan& p, uint32_t gen, uint32_t slot) {
      UnitIO u;
      memset(&u, 0, sizeof u);
      const size_t i = p.unit;
      u.in = (const uint8_t*)cur[i].data(); u.in_len = len[i];
      u.out = stage[i].data() + p.stage_off; u.out_cap = p.stage_cap;
      u.cnt = use_events ? cnts[i].data() : nullptr;
      u.okb = (use_events && !getenv("JB_SIM_NO_OKB")) ? okbs[i].data() : nullptr;
      u.weak = use_events ? weak[i].data() : nullptr;
      u.vmap = use_verdicts ? vmaps[i].data() : nullptr; u.verdicts = verdicts[i].data();
      u.verdicts2 = (use_verdicts && use_verdicts2) ? verdicts2[i].data() : nullptr;
      u.ev = use_events ? ev[i].data() : nullptr; u.ev_l1 = use_events ? l1[i].data() : nullptr; u.ev_l2 = (use_events && !getenv("JB_SIM_NO_L2")) ? l2[i].data() : nullptr; u.ev_bit_base = 0;
      u.base = p.base; u.next_anchor = p.next_anchor; u.unit = p.unit; u.seg = p.seg; u.gen = gen;
      u.slot = slot; u.materialize_all = gen ? 1u : 0u;
  u.keep_margin = keep_margin;
      return u;
    };
    std::vector<UnitIO> segs(plan.size());
    std::vector<uint8_t> redo(n, 0);
    bool any_redo = false;
    for (size_t q = 0; q < plan.size(); ++q) {
      segs[q] = make_seg(plan[q], 0, (uint32_t)q);
      Walker<HostDB> w(db, k, sampling_step(k), solid, do_fix, segs[q], ws, pool, sink, jobs);
      w.run();
      if (segs[q].status) {
        if (segs[q].status == kTrieFull || segs[q].status == kFarBack || (plan[q].n_segs > 1 && (segs[q].status == kSeamViolation || segs[q].status == JB_E_SCRATCH))) {
          redo[plan[q].unit] = 1; any_redo = true;
          if (segs[q].status == kFarBack) ++res->far_backs;
        }
        else { res->status = segs[q].status; return res; }
      }
    }
    res->segments += plan.size();
    if (getenv("JB_SIM_TRACE")) {
      unsigned long long ev = 0, fast = 0; size_t na = 0;
      for (const UnitIO& u : segs) { ev += u.events; fast += u.n_fast; }
      for (long long a : anchors) na += a >= 0;
      fprintf(stderr, "[sim] pass %d: %zu segments, %zu blocks, %zu anchors, %llu events, %llu from verdicts\n", ite, plan.size(), blocks.size(), na, ev, fast);
    }
    size_t n_slots = plan.size();
    if (any_redo) {
      std::vector<UnitIO> merged;
      for (size_t q = 0; q < segs.size();) {
        const uint32_t unit = segs[q].unit;
        if (!redo[unit]) { merged.push_back(segs[q++]); continue; }
        UnitIO u = make_seg(whole_contig(unit, caps[unit]), 1, (uint32_t)n_slots++);
        {
          Walker<HostDB> w(db, k, sampling_step(k), solid, do_fix, u, ws, pool, sink, jobs);
          w.run();
        }
        if (u.status == kTrieFull && trie.size() < (1u << 26)) {      // the product's ladder: once more with the grown pool
          trie.assign(1u << 26, 0);
          bs.trie = trie.data(); bs.trie_cap = (uint32_t)trie.size();
          ++res->pool_growths;
          u = make_seg(whole_contig(unit, caps[unit]), 2, (uint32_t)n_slots++);
          Walker<HostDB> w2(db, k, sampling_step(k), solid, do_fix, u, ws, pool, sink, jobs);
          w2.run();
        }
        if (u.status) { res->status = u.status; return res; }
        merged.push_back(u);
        ++res->reruns;
        while (q < segs.size() && segs[q].unit == unit) ++q;
      }
      segs.swap(merged);
    }
    for (const UnitIO& u : segs) { po.wrong += u.wrong; res->lookups += u.lookups; res->events += u.events; res->fast += u.n_fast; res->post += u.n_post; res->v2 += u.n_v2; }
    if (do_fix) {
      std::vector<int64_t> new_lens(len);
      std::vector<SegPlace> place = place_segments(segs, &new_lens);
      std::vector<size_t> first_seg(n, 0);
      std::vector<std::string> next(n);
      std::vector<std::vector<uint8_t>> covered(n);
      for (size_t i = 0; i < n; ++i) { next[i].assign((size_t)new_lens[i], '?'); covered[i].assign((size_t)new_lens[i], 0); }
      std::vector<int64_t> dst_off(n_slots, -1);
      std::vector<uint32_t> slot_unit(n_slots, 0);
      for (size_t q = 0; q < segs.size(); ++q) {
        const UnitIO& u = segs[q];
        if (u.seg == 0) first_seg[u.unit] = q;
        dst_off[u.slot] = place[q].new_off; slot_unit[u.slot] = u.unit;
      }
      // what k_concat does: every copy job of an accepted segment; the jobs must tile the new contig exactly
      for (unsigned long long j = 0; j < jobs_used; ++j) {
        const CopyJob& job = jobv[j];
        if (dst_off[job.slot] < 0) continue;
        const uint32_t unit = slot_unit[job.slot];
        const int64_t at = dst_off[job.slot] + job.rel;
        if (at < 0 || at + job.len > new_lens[unit]) { res->status = -100; return res; }
        for (int64_t q = 0; q < job.len; ++q) {
          if (covered[unit][at + q]) { res->status = -100; return res; }
          covered[unit][at + q] = 1; next[unit][at + q] = (char)job.src[q];
        }
      }
      for (size_t i = 0; i < n; ++i) {
        if (new_lens[i] > caps[i]) { res->status = -100; return res; }
        for (int64_t q = 0; q < new_lens[i]; ++q) if (!covered[i][q]) { res->status = -100; return res; }
        cur[i].swap(next[i]);
      }
      std::vector<DevRow> got(rows.begin(), rows.begin() + used);
      std::vector<uint8_t> a2(aux.begin(), aux.begin() + aux_used);
      // the product's path (no copy, no sort) and the sorting path must agree
      PassOut alt;
      const bool ok1 = rows_in_order(rows.data(), used, aux.data(), segs, place, first_seg, &po);
      settle_rows(&got, segs, place, first_seg);
      const bool ok2 = rows_from_log(got, a2, &alt);
      if (ok1 != ok2 || alt.rows.size() != po.rows.size()) { res->status = -100; return res; }
      for (size_t q = 0; q < alt.rows.size(); ++q) {
        jb_fix_row_t x, y;
        alt.expand(alt.rows[q], &x); po.expand(po.rows[q], &y);
        if (memcmp(&x, &y, sizeof x)) { res->status = -100; return res; }
      }
      if (!ok1) res->status = JB_E_REFERENCE_CRASH;
    }
    for (const HostRow& r : po.rows) {
      jb_fix_row_t x;
      po.expand(r, &x);
      char line[256];
      snprintf(line, sizeof line, "%d\t%u\t%lld\t%s\t%s\n", ite, x.seq_index, (long long)x.coord, x.mutation, x.original);
      res->text += line;
    }
    res->passes.push_back(po);
    if (res->status) break;
  }
  res->seqs = cur;
  return res;
}

extern "C" int hostsim_status(const SimResult* r) { return r->status; }
extern "C" const char* hostsim_rows_text(const SimResult* r) { return r->text.c_str(); }
extern "C" int hostsim_iterations(const SimResult* r) { return (int)r->passes.size(); }
extern "C" void hostsim_tally(const SimResult* r, int ite, long long* wrong, long long* total) { *wrong = r->passes[ite].wrong; *total = r->passes[ite].total; }
extern "C" const char* hostsim_seq(const SimResult* r, int i, uint64_t* len) { *len = r->seqs[i].size(); return r->seqs[i].data(); }
extern "C" unsigned long long hostsim_lookups(const SimResult* r) { return r->lookups; }
extern "C" unsigned long long hostsim_events(const SimResult* r) { return r->events; }
extern "C" unsigned long long hostsim_fast(const SimResult* r) { return r->fast; }
extern "C" unsigned long long hostsim_post(const SimResult* r) { return r->post; }
extern "C" unsigned long long hostsim_v2(const SimResult* r) { return r->v2; }
extern "C" unsigned long long hostsim_segments(const SimResult* r) { return r->segments; }
extern "C" unsigned long long hostsim_reruns(const SimResult* r) { return r->reruns; }
extern "C" unsigned long long hostsim_pool_growths(const SimResult* r) { return r->pool_growths; }
extern "C" unsigned long long hostsim_far_backs(const SimResult* r) { return r->far_backs; }
extern "C" void hostsim_free(SimResult* r) { delete r; }

// also expose the shared k-mer arithmetic for unit tests
extern "C" void hostsim_mer(const char* s, uint64_t len, int k, uint64_t out[2], uint64_t canon[2], uint64_t rc[2]) {
  Key mask = key_mask(k);
  Key m = key_from_chars((const uint8_t*)s, (int64_t)len, k, mask);
  Key c = key_canonical(m, k), r = key_revcomp(m, k);
  out[0] = m.lo; out[1] = m.hi; canon[0] = c.lo; canon[1] = c.hi; rc[0] = r.lo; rc[1] = r.hi;
}
extern "C" int hostsim_align(const char* a, const char* b, char* out_a, char* out_b, int cap) {
  std::string x, y;
  if (!first_global_alignment(a, b, &x, &y)) return 0;
  snprintf(out_a, cap, "%s", x.c_str()); snprintf(out_b, cap, "%s", y.c_str());
  return 1;
}

// Self-check of the packed-key fast path (jb_cand.cuh) against the byte loop it replaces: random strings with
// invalid bytes, every candidate kind, every window.  Returns mismatches; *fast = windows the fast path served.
extern "C" long long hostsim_packed_selfcheck(unsigned seed, int k, long long* fast) {
  static WarpScratch s;
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + 12345;
  auto rnd = [&]() { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; return (uint32_t)((x * 0x2545F4914F6CDD1Dull) >> 33); };
  const char alpha[] = "ACGTacgt";
  long long bad = 0; *fast = 0;
  const Key mask = key_mask(k);
  for (int round = 0; round < 200; ++round) {
    const int n = k + (int)(rnd() % (unsigned)(k + 10));
    for (int q = 0; q < n; ++q) s.t[q] = (rnd() % 40 == 0) ? (uint8_t)"NnRy-"[rnd() % 5] : (uint8_t)alpha[rnd() % 8];
    const int nb = (int)(rnd() % (unsigned)(k / 2 + 1)), bl = (int)(rnd() % (unsigned)(k / 2 + 1));
    for (int q = 0; q < nb; ++q) s.c[q] = (uint8_t)alpha[rnd() % 8];
    for (int q = 0; q < bl; ++q) s.c[nb + n + q] = (rnd() % 30 == 0) ? (uint8_t)'N' : (uint8_t)alpha[rnd() % 8];
    s.dip_nb = nb; s.dip_bl = bl;
    for (int kind = 0; kind < 4; ++kind) {
      Cand c; c.kind = (uint8_t)kind; c.base = (uint8_t)"ACGT"[rnd() % 4]; c.base2 = (uint8_t)"ACGT"[rnd() % 4]; c.pad = 0;
      c.pos = (int16_t)(rnd() % (unsigned)n); c.m = (int16_t)(1 + rnd() % (unsigned)(k - 1));
      if (kind == kDel && c.pos + c.m > n) c.m = (int16_t)(n - c.pos);
      if (kind == kDel && c.m <= 0) continue;
      if (kind == kDip) { for (int q = 0; q < n; ++q) s.c[nb + q] = s.t[q]; pack_bytes(s.pt, s.c, nb + n + bl); }
      else pack_bytes(s.pt, s.t, n);
      const int len = cand_len(s, c, n);
      for (int m = 0; m <= len; ++m) {
        Key slow = {0, 0}; int j = 0;
        for (; j < k && m + j < len; ++j) { uint8_t ch = cand_char(s, c, n, k, m + j); if (!base_valid(ch)) break; slow = key_push(slow, base_code(ch), mask); }
        slow = key_shl_bases(slow, k - j, mask);
        Key f;
        if (cand_key_packed(s, c, n, len, m, k, &f)) { ++*fast; if (f.lo != slow.lo || f.hi != slow.hi) ++bad; }
      }
    }
    // plain buffers: count_packed's choice against key_from_chars
    const int len = n;
    pack_bytes(s.pc, s.t, len);
    for (int m = 0; m <= len; ++m) {
      Key slow = key_from_chars(s.t + m, (int64_t)len - m, k, mask);
      if (packed_clean(s.pc, m, k)) { ++*fast; Key f = packed_window(s.pc, m, k); if (f.lo != slow.lo || f.hi != slow.hi) ++bad; }
    }
  }
  return bad;
}

// walker_hostsim.cpp -- TEST-ONLY differential build of the walker (jasper_b200/csrc/jb_walker.cuh)
// for the CPU: the same source the GPU runs, with a warp emulated by a 32-iteration lane loop and
// table probes answered through a callback (the oracle's lookup).  It lets `-m "not gpu"` tests
// compare the device state machine against oracle/jasper_oracle.py without a GPU.
// It is never linked into libjasper_b200.so and nothing in jasper_b200/ loads it.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../jasper_b200/csrc/jb_rows.hpp"

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
  unsigned long long lookups = 0, events = 0;
};

// event bitmask exactly as k_scan<.,true> defines it (jb_scan.cuh)
static std::vector<uint32_t> event_bits(const HostDB& db, const std::string& seq, int k, uint32_t solid) {
  const int64_t L = (int64_t)seq.size();
  std::vector<uint32_t> cnt(L + 1);
  std::vector<uint8_t> ok(L + 1);
  Key mask = key_mask(k);
  for (int64_t p = 0; p <= L; ++p) {
    Key m = key_from_chars((const uint8_t*)seq.data() + p, L - p, k, mask);
    cnt[p] = db.count(key_canonical(m, k));
    bool v = p + k <= L;
    for (int q = 0; v && q < k; ++q) v = base_valid((uint8_t)seq[p + q]);
    ok[p] = v;
  }
  std::vector<uint32_t> ev((L + 1 + 31) / 32 + 1, 0);
  for (int64_t p = 0; p <= L; ++p) {
    bool e = !ok[p] || p < k || cnt[p] < solid;
    if (!e) e = (unsigned long long)cnt[p] * JB_DIVISOR < (unsigned long long)cnt[p - k];
    if (e) ev[p >> 5] |= 1u << (p & 31);
  }
  return ev;
}

extern "C" SimResult* hostsim_polish(lookup_fn fn, const void* dbh, int k, uint32_t solid, int passes, int fix,
                                     int n_seqs, const char* const* seqs, const uint64_t* lens, int use_events) {
  HostDB db{fn, dbh};
  auto* res = new SimResult();
  std::vector<std::string> cur(n_seqs);
  for (int i = 0; i < n_seqs; ++i) cur[i].assign(seqs[i], lens[i]);
  std::vector<DevRow> rows(1 << 16);
  std::vector<uint8_t> aux(1 << 22);
  static BridgeScratch bs;                     // large: keep off the stack
  std::vector<uint32_t> trie(1u << 20);
  std::vector<uint8_t> patch(1u << 18);
  bs.trie = trie.data(); bs.trie_cap = (uint32_t)trie.size(); bs.patch = patch.data(); bs.patch_cap = (uint32_t)patch.size();
  WarpScratch ws;
  for (int ite = 0; ite <= passes; ++ite) {
    const bool do_fix = fix && ite != passes;
    PassOut po;
    unsigned long long used = 0, aux_used = 0;
    RowSink sink{rows.data(), rows.size(), &used, aux.data(), aux.size(), &aux_used};
    for (int i = 0; i < n_seqs; ++i) {
      const int64_t L = (int64_t)cur[i].size();
      po.total += L - k + 1;
      std::vector<uint32_t> ev;
      if (use_events) ev = event_bits(db, cur[i], k, solid);
      std::vector<uint8_t> out((size_t)(L + L / 8 + 4096));
      UnitIO u;
      memset(&u, 0, sizeof u);
      u.in = (const uint8_t*)cur[i].data(); u.in_len = L; u.out = out.data(); u.out_cap = (int64_t)out.size();
      u.ev = use_events ? ev.data() : nullptr; u.unit = (uint32_t)i;
      Walker<HostDB> w(db, k, sampling_step(k), solid, do_fix, u, ws, &bs, sink);
      w.run();
      if (u.status) { res->status = u.status; return res; }
      po.wrong += u.wrong;
      res->lookups += u.lookups; res->events += u.events;
      if (do_fix) {
        std::string next((const char*)out.data(), (size_t)u.fin_olen);
        next.append(cur[i].data() + u.fin_ipos, (size_t)(L - u.fin_ipos));
        if ((int64_t)next.size() != u.out_len) { res->status = -100; return res; }
        cur[i].swap(next);
      }
    }
    if (do_fix) {
      std::vector<DevRow> got(rows.begin(), rows.begin() + used);
      std::vector<uint8_t> a2(aux.begin(), aux.begin() + aux_used);
      if (!rows_from_log(got, a2, &po.rows)) res->status = JB_E_REFERENCE_CRASH;
    }
    for (const HostRow& r : po.rows) {
      char line[256];
      snprintf(line, sizeof line, "%d\t%u\t%lld\t%s\t%s\n", ite, r.seq_index, (long long)r.coord, r.mutation.c_str(), r.original.c_str());
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

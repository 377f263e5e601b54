"""Shared test plumbing: the oracle (checker), golden fixtures, the host-sim walker build."""
import ctypes as C
import glob
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import jasper_oracle as jo  # noqa: E402
from jasper_b200 import synth  # noqa: E402

GOLDEN_NAMES = ["k25_mix", "k37_mix", "k31_dirty", "k21_even"]
# fuzz cases through the UNMODIFIED reference, chosen for the rare branches (tests/golden/make_golden_branches.py)
DEEP_BRIDGE = "branches/deep_bridge"     # a de-Bruijn search of > 1 M path extensions (tests/golden/make_golden_deep_bridge.py)
BRANCH_GOLDENS = ["branches/s32", "branches/s99", "branches/s278", "branches/s377", "branches/s5", "branches/s9", "branches/s39", "branches/s66"]
ALL_BRANCHES = ["add_base", "bridge", "drop_base", "grow_run", "grow_single", "rolling_ok", "rolling_weak", "shrink_run", "shrink_single", "substitution",
                "two_ended_e", "two_ended_s", "unfixed_k", "unfixed_long", "unfixed_short"]
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _newer(target, *sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


def oracle_lib():
    """oracle/libjf_oracle.so (C restatement of the Jellyfish lookup path), built on demand with gcc."""
    so = os.path.join(ROOT, "oracle", "libjf_oracle.so")
    src = os.path.join(ROOT, "oracle", "jf_oracle.c")
    if not _newer(so, src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", so, src])
    lib = C.CDLL(so)
    lib.jfo_open.restype = C.c_void_p
    lib.jfo_open.argtypes = [C.c_char_p]
    lib.jfo_close.argtypes = [C.c_void_p]
    lib.jfo_last_error.restype = C.c_char_p
    lib.jfo_count.restype = C.c_uint32
    lib.jfo_count.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.jfo_lookup_key.restype = C.c_uint64
    lib.jfo_lookup_key.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.jfo_count_positions.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)]
    lib.jfo_k.argtypes = [C.c_void_p]
    lib.jfo_n_records.restype = C.c_uint64
    lib.jfo_n_records.argtypes = [C.c_void_p]
    lib.jfo_record.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.jfo_hash_pos.restype = C.c_uint64
    lib.jfo_hash_pos.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.jfo_mer_from_chars.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]
    lib.jfo_canonical.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]
    lib.jfo_revcomp.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]
    lib.jfo_mer_to_string.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_char_p]
    return lib


class OracleDB(object):
    """C(s) through the C oracle."""

    def __init__(self, path):
        self.lib = oracle_lib()
        self.h = self.lib.jfo_open(os.fsencode(path))
        if not self.h:
            raise RuntimeError(self.lib.jfo_last_error().decode())
        self.k = self.lib.jfo_k(self.h)
        self.n_records = self.lib.jfo_n_records(self.h)

    def count(self, s):
        b = s.encode("latin-1") if isinstance(s, str) else s
        return self.lib.jfo_count(self.h, b, len(b))

    def count_positions(self, seq):
        b = seq.encode("latin-1") if isinstance(seq, str) else seq
        out = np.empty(len(b) + 1, dtype=np.uint32)
        self.lib.jfo_count_positions(self.h, b, len(b), out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def records(self):
        n = self.n_records
        lo = np.empty(n, dtype=np.uint64)
        hi = np.empty(n, dtype=np.uint64)
        cnt = np.empty(n, dtype=np.uint64)
        key = (C.c_uint64 * 2)()
        val = C.c_uint64()
        for i in range(n):
            self.lib.jfo_record(self.h, i, key, C.byref(val))
            lo[i], hi[i], cnt[i] = key[0], key[1], val.value
        return lo, hi, cnt


def reference_binding():
    """The reference's own compiled binding (oracle/_ref), or None when it was not built."""
    if not glob.glob(os.path.join(ROOT, "oracle", "_ref", "_dna_jellyfish*.so")):
        return None
    import dna_jellyfish
    return dna_jellyfish


def hostsim_lib():
    """tests/hostsim/libwalker_hostsim.so: the walker source compiled for the CPU (test-only)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libwalker_hostsim.so")
    srcs = [os.path.join(d, "walker_hostsim.cpp")] + glob.glob(os.path.join(ROOT, "jasper_b200", "csrc", "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "jasper_b200", "csrc", "*.hpp")) + [os.path.join(ROOT, "jasper_b200", "csrc", "jb_align.cpp")]
    if not _newer(so, *srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", so,
                               os.path.join(d, "walker_hostsim.cpp"), os.path.join(ROOT, "jasper_b200", "csrc", "jb_align.cpp")])
    sim = C.CDLL(so)
    sim.hostsim_polish.restype = C.c_void_p
    sim.hostsim_polish.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_int]
    sim.hostsim_rows_text.restype = C.c_char_p
    sim.hostsim_rows_text.argtypes = [C.c_void_p]
    sim.hostsim_status.argtypes = [C.c_void_p]
    sim.hostsim_iterations.argtypes = [C.c_void_p]
    sim.hostsim_seq.restype = C.POINTER(C.c_char)
    sim.hostsim_seq.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    sim.hostsim_tally.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    sim.hostsim_free.argtypes = [C.c_void_p]
    sim.hostsim_segments.restype = C.c_ulonglong
    sim.hostsim_segments.argtypes = [C.c_void_p]
    sim.hostsim_reruns.restype = C.c_ulonglong
    sim.hostsim_reruns.argtypes = [C.c_void_p]
    sim.hostsim_pool_growths.restype = C.c_ulonglong
    sim.hostsim_pool_growths.argtypes = [C.c_void_p]
    for fn in (sim.hostsim_events, sim.hostsim_fast, sim.hostsim_post, sim.hostsim_v2):
        fn.restype = C.c_ulonglong
        fn.argtypes = [C.c_void_p]
    sim.hostsim_packed_selfcheck.restype = C.c_longlong
    sim.hostsim_packed_selfcheck.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_longlong)]
    sim.hostsim_mer.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    sim.hostsim_align.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    return sim


def table_sim_lib():
    """tests/hostsim/libtable_hostsim.so: the compact table's arithmetic (jb_kmer.cuh, jb_geom.hpp) on the CPU (test-only)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libtable_hostsim.so")
    csrc = os.path.join(ROOT, "jasper_b200", "csrc")
    srcs = [os.path.join(d, "table_hostsim.cpp"), os.path.join(csrc, "jb_header.cpp"), os.path.join(csrc, "jb_geom.hpp"),
            os.path.join(csrc, "jb_kmer.cuh"), os.path.join(csrc, "jb_header.hpp")]
    if not _newer(so, *srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, srcs[0], srcs[1]])
    sim = C.CDLL(so)
    sim.hostsim_table_check.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_char_p, C.c_int]
    return sim


TABLE_STATS = ("layout", "lines", "tag_bits", "cnt_bits", "n_stash", "max_disp_lines", "lines_per_hit_x1000", "lines_per_miss_x1000", "records")


def table_sim_check(sim, path, load=0.0, layout=-1, max_cval_bits=0):
    """-> (stats dict, message); raises AssertionError when the table arithmetic breaks an invariant."""
    st = (C.c_longlong * 9)()
    msg = C.create_string_buffer(512)
    rc = sim.hostsim_table_check(os.fsencode(path), load, layout, max_cval_bits, st, msg, 512)
    assert rc == 0, msg.value.decode()
    return dict(zip(TABLE_STATS, st)), msg.value.decode()


def colliding_db(path, k=25, n_random=20000, n_collide=100, r=16, seed=3, big_counts=0):
    """A .jf whose identity 'matrix' sends n_collide keys to ONE hash position: their run crosses more lines than a compact
    probe follows, so some end up as tombstones (records kept in the stash).  big_counts: that many records get counts
    beyond any compact count field."""
    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 1 << (2 * k), n_random, dtype=np.uint64)
    low = np.uint64(int(rng.integers(0, 1 << r)))
    hi_part = rng.choice(1 << (2 * k - r), n_collide, replace=False).astype(np.uint64)
    lo = np.unique(np.concatenate([lo, (hi_part << np.uint64(r)) | low]))
    cnt = rng.integers(1, 200, len(lo)).astype(np.uint64)
    if big_counts:
        cnt[rng.choice(len(lo), big_counts, replace=False)] = rng.integers(1 << 29, 1 << 32, big_counts).astype(np.uint64)
    synth.write_jf(path, lo, np.zeros(len(lo), dtype=np.uint64), cnt, k, r=r, identity=True)
    return lo, cnt


def hostsim_polish(sim, odb, seqs, k, thr, passes, use_events=1, fix=1, stats=None):
    """Run the CPU build of the device walker; returns (status, rows_text, seqs, tallies).
    use_events: 0 = no event bitmask, 1 = bitmask, >= 64 = also split contigs into segments (one anchor per that many bases)."""
    n = len(seqs)
    enc = [s.encode("latin-1") for _, s in seqs]
    arr = (C.c_char_p * max(1, n))(*enc)
    lens = (C.c_uint64 * max(1, n))(*[len(b) for b in enc])
    fn = C.cast(odb.lib.jfo_lookup_key, C.c_void_p)
    r = sim.hostsim_polish(fn, odb.h, k, thr, passes, fix, n, arr, lens, use_events)
    try:
        st = sim.hostsim_status(r)
        txt = sim.hostsim_rows_text(r).decode("latin-1")
        out = []
        if st == 0:
            for i in range(n):
                ln = C.c_uint64()
                p = sim.hostsim_seq(r, i, C.byref(ln))
                out.append(C.string_at(p, ln.value).decode("latin-1"))
        tallies = []
        for ite in range(sim.hostsim_iterations(r)):
            w, t = C.c_longlong(), C.c_longlong()
            sim.hostsim_tally(r, ite, C.byref(w), C.byref(t))
            tallies.append((w.value, t.value))
        if stats is not None:
            stats["segments"] = sim.hostsim_segments(r)
            stats["reruns"] = sim.hostsim_reruns(r)
            stats["pool_growths"] = sim.hostsim_pool_growths(r)
            stats["events"] = sim.hostsim_events(r)
            stats["fast"] = sim.hostsim_fast(r)
            stats["post"] = sim.hostsim_post(r)
            stats["v2"] = sim.hostsim_v2(r)
        return st, txt, out, tallies
    finally:
        sim.hostsim_free(r)


def rows_text(passes, names):
    """Oracle PassReports -> the same text form hostsim / the product rows are compared in."""
    out = []
    for ite, p in enumerate(passes):
        for row in p.rows:
            out.append("%d\t%d\t%d\t%s\t%s\n" % (ite, names.index(row[0]), row[1], row[2], row[3]))
    return "".join(out)


def product_rows_text(rows_per_pass, names):
    out = []
    for ite, rows in enumerate(rows_per_pass):
        for row in rows:
            out.append("%d\t%d\t%d\t%s\t%s\n" % (ite, names.index(row[0]), row[1], row[2], row[3]))
    return "".join(out)


def load_golden(name, tmpdir):
    d = os.path.join(GOLDEN_DIR, name)
    meta = json.load(open(os.path.join(d, "case.json")))
    db = os.path.join(tmpdir, name.replace("/", "_") + ".jf")
    if not os.path.exists(db):
        with gzip.open(os.path.join(d, "db.jf.gz")) as f, open(db, "wb") as o:
            o.write(f.read())
    meta["db"] = db
    meta["dir"] = d
    meta["fasta"] = os.path.join(d, "asm.fa")
    lk = os.path.join(d, "lookups.json")
    meta["lookups"] = json.load(open(lk)) if os.path.exists(lk) else []
    return meta


def expected_file(case, fn):
    with open(os.path.join(case["dir"], "expected", fn), "rb") as f:
        return f.read()


# ----------------------------------------------------------------------------- synthetic fuzz cases
def fuzz_case(seed, k, big=False):
    """A small random polishing problem with every kind of dirt the reference tolerates."""
    rng = np.random.default_rng(seed)
    ncont = int(rng.integers(1, 4))
    haps, recs = [], []
    het = float(rng.choice([0, 0, 0.003, 0.01]))
    for ci in range(ncont):
        n = int(rng.integers(200, 6000)) * (8 if big else 1)
        g = synth.random_genome(n, rng, homopolymers=int(rng.integers(0, 12)), repeat_copies=int(rng.choice([0, 0, 6, 20])),
                                repeat_len=int(rng.integers(50, 300)))
        haps.append(g)
        if het:
            h2 = g.copy()
            m = rng.random(n) < het
            h2[m] = (h2[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
            haps.append(h2)
        a = synth.inject_errors(g, float(rng.choice([0.002, 0.005, 0.01, 0.03])), rng)
        s = bytearray(synth.codes_to_bytes(a))
        if rng.random() < 0.3:
            for _ in range(int(rng.integers(1, 4))):
                ln = int(rng.integers(1, 40)); at = int(rng.integers(0, max(1, len(s) - ln)))
                s[at:at + ln] = (b"N" if rng.random() < 0.6 else b"n") * ln
        if rng.random() < 0.3:
            for _ in range(int(rng.integers(1, 4))):
                ln = int(rng.integers(1, 100)); at = int(rng.integers(0, max(1, len(s) - ln)))
                s[at:at + ln] = bytes(s[at:at + ln]).lower()
        if rng.random() < 0.2:
            for _ in range(int(rng.integers(1, 4))):
                at = int(rng.integers(0, len(s))); s[at] = int(rng.choice(list(b"RYKMSWryB-")))
        recs.append(("c%d" % ci, bytes(s).decode()))
    if rng.random() < 0.25:
        recs.append(("tiny", "ACGT" * int(rng.integers(0, 4))))
    cov = float(rng.choice([8, 15, 30]))
    reads = np.concatenate([synth.simulate_reads([h], cov / (2 if het else 1), 100, float(rng.choice([0.0, 0.002, 0.01])), rng)
                            for h in haps])
    lo, hi, cnt = synth.kmer_db_from_reads(reads, k)
    return recs, (lo, hi, cnt), int(rng.choice([2, 3, 4, 6]))


FUZZ_KS = [25, 37, 31, 20, 48, 8, 12, 33]


def write_fuzz_db(tmpdir, seed, k, big=False):
    recs, (lo, hi, cnt), thr = fuzz_case(seed, k, big)
    path = os.path.join(str(tmpdir), "fuzz_%d_%d.jf" % (seed, k))
    synth.write_jf(path, lo, hi, cnt, k, seed=seed)
    return recs, path, thr


def oracle_polish(odb, seqs, k, thr, passes):
    o = jo.PolishOracle(odb.count, k, thr)
    try:
        return o.polish(seqs, passes), o
    except jo.ReferenceCrash:
        return None, o

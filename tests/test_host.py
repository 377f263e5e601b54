"""CPU-side checks of the product: the C-ABI library loads and exports everything include/*.h
declares, fails loudly without a GPU, and the walker state machine (the very source the GPU runs,
compiled for the host in tests/hostsim -- test-only) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
import jasper_oracle as jo
from jasper_b200 import capi, synth

ROOT = helpers.ROOT


def _built():
    return os.path.exists(capi.library_path())


def test_library_exports_every_declared_symbol():
    if not _built():
        from jasper_b200 import build
        build.build()
    header = open(os.path.join(ROOT, "include", "jasper_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    lib = C.CDLL(capi.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTS) == declared


def test_no_cpu_fallback(golden_cases):
    """Without a CUDA device the product refuses to compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load_library()
    assert lib.jb_device_count() == 0
    with pytest.raises(capi.JasperError) as e:
        capi.QueryMerFile(golden_cases["k25_mix"]["db"])
    assert e.value.code == -4 and "no CPU path" in str(e.value)


def test_solid_threshold_entry_point():
    rows = [(1, 100), (2, 50), (3, 20), (4, 10), (5, 12), (6, 30)]
    assert capi.solid_threshold(rows) == synth.solid_threshold_from_histogram(rows) == 2
    with pytest.raises(capi.JasperError):
        capi.solid_threshold([(1, 100), (2, 50), (3, 60)])
    with pytest.raises(capi.JasperError):
        capi.solid_threshold([(1, 100), (2, 50), (3, 40)])


def test_kmer_arithmetic_matches_oracle(hostsim_lib, oracle_lib):
    """jb_kmer.cuh (2-bit pack with MerDNA padding, reverse complement, canonical) vs the C oracle."""
    rng = np.random.default_rng(3)
    alphabet = b"ACGTacgtNnRY-\0"
    for k in (8, 20, 25, 31, 32, 33, 37, 48):
        for _ in range(300):
            s = bytes(alphabet[int(x)] for x in rng.integers(0, 9 if rng.random() < 0.7 else len(alphabet), int(rng.integers(0, 60))))
            m, c, r = (C.c_uint64 * 2)(), (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
            hostsim_lib.hostsim_mer(s, len(s), k, m, c, r)
            om, oc, orc = (C.c_uint64 * 2)(), (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
            oracle_lib.jfo_mer_from_chars(s, len(s), k, om)
            oracle_lib.jfo_canonical(om, k, oc)
            oracle_lib.jfo_revcomp(om, k, orc)
            assert (m[0], m[1]) == (om[0], om[1]), (k, s)
            assert (r[0], r[1]) == (orc[0], orc[1]), (k, s)
            assert (c[0], c[1]) == (oc[0], oc[1]), (k, s)


def test_alignment_matches_shim(hostsim_lib):
    """jb_align.cpp vs oracle/Bio/pairwise2.py (both restate Bio.pairwise2; parity unpinned vs Biopython)."""
    from Bio import pairwise2
    rng = np.random.default_rng(11)
    for _ in range(400):
        n = int(rng.integers(1, 40))
        a = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))
        b = list(a)
        for _ in range(int(rng.integers(0, 5))):
            at = int(rng.integers(0, max(1, len(b))))
            op = rng.integers(0, 3)
            if op == 0 and b:
                b[at] = "ACGT"[int(rng.integers(0, 4))]
            elif op == 1:
                b.insert(at, "ACGT"[int(rng.integers(0, 4))])
            elif b:
                del b[at]
        b = "".join(b)
        if not b:
            continue
        want = pairwise2.first_global_alignment(a, b)[0]
        oa, ob = C.create_string_buffer(256), C.create_string_buffer(256)
        assert hostsim_lib.hostsim_align(a.encode(), b.encode(), oa, ob, 256) == 1
        assert (oa.value.decode(), ob.value.decode()) == (want[0], want[1]), (a, b)


def test_alignment_shortcut_equals_the_matrix(hostsim_lib, monkeypatch):
    """Equal-length pairs with up to two mismatches skip the matrix when the gapless alignment is the only optimum; the
    result must equal the full traceback on every such pair, ties (shifted repeats, homopolymers) included."""
    rng = np.random.default_rng(5)
    pairs = []
    for _ in range(6000):
        n = int(rng.integers(1, 30))
        alpha = "ACGT"[:int(rng.integers(1, 5))]                 # small alphabets make repeats and ties frequent
        a = "".join(alpha[int(x)] for x in rng.integers(0, len(alpha), n))
        b = list(a)
        for _ in range(int(rng.integers(0, 4))):
            b[int(rng.integers(0, n))] = "ACGT"[int(rng.integers(0, 4))]
        pairs.append((a.encode(), "".join(b).encode()))

    def run():
        out = []
        oa, ob = C.create_string_buffer(256), C.create_string_buffer(256)
        for a, b in pairs:
            ok = hostsim_lib.hostsim_align(a, b, oa, ob, 256)
            out.append((ok, oa.value, ob.value) if ok else (0, b"", b""))
        return out
    fast = run()
    monkeypatch.setenv("JB_ALIGN_NO_FAST", "1")
    slow = run()
    monkeypatch.delenv("JB_ALIGN_NO_FAST")
    assert fast == slow


def test_alignment_band_equals_the_matrix(hostsim_lib, monkeypatch):
    """The matrix is only filled in a band around the diagonal (every gap character and mismatch costs 1, so an optimal
    alignment of cost d stays within |i - j| <= d; the band is widened when the cost found does not fit).  Same first
    alignment as the full matrix on pairs with substitutions, insertions and deletions, small alphabets (ties) included, and
    on pairs that need the widened band."""
    rng = np.random.default_rng(9)
    pairs = []
    for it in range(5000):
        n = int(rng.integers(1, 120))
        alpha = "ACGT"[:int(rng.integers(1, 5))]
        a = [alpha[int(x)] for x in rng.integers(0, len(alpha), n)]
        b = list(a)
        for _ in range(int(rng.integers(0, 5)) if it % 10 else int(rng.integers(5, 40))):      # every tenth pair: many edits
            op, at = int(rng.integers(0, 3)), int(rng.integers(0, max(1, len(b))))
            if op == 0 and b:
                b[at] = "ACGT"[int(rng.integers(0, 4))]
            elif op == 1:
                b.insert(at, "ACGT"[int(rng.integers(0, 4))])
            elif b:
                del b[at]
        if b:
            pairs.append(("".join(a).encode(), "".join(b).encode()))

    def run():
        out = []
        oa, ob = C.create_string_buffer(512), C.create_string_buffer(512)
        for a, b in pairs:
            ok = hostsim_lib.hostsim_align(a, b, oa, ob, 512)
            out.append((ok, oa.value, ob.value) if ok else (0, b"", b""))
        return out
    banded = run()
    monkeypatch.setenv("JB_ALIGN_NO_BAND", "1")
    full = run()
    monkeypatch.delenv("JB_ALIGN_NO_BAND")
    assert banded == full


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
@pytest.mark.parametrize("use_events", [0, 1, 64, 300, 2048])
def test_walker_source_matches_golden(golden_cases, hostsim_lib, name, use_events):
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    seqs = jo.read_fasta(case["fasta"])
    res, _ = helpers.oracle_polish(odb, seqs, case["k"], case["threshold"], case["passes"])
    st, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], use_events)
    assert st == 0
    names = [n for n, _ in seqs]
    assert txt == helpers.rows_text(res.passes, names)
    assert out == [s for _, s in res.seqs]
    assert tallies == [(p.wrong, p.total) for p in res.passes]
    # and therefore the reference's own bytes
    rows0 = [[names[int(f[1])], int(f[2]), f[3], f[4]] for f in (l.split("\t") for l in txt.splitlines()) if f[0] == "0"]
    assert jo.format_fix_csv(rows0) == helpers.expected_file(case, "_iter0_asm.fa.fix.csv")
    assert jo.format_fasta(list(zip(names, out))) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))


@pytest.mark.parametrize("name", helpers.BRANCH_GOLDENS)
def test_walker_source_matches_rare_branch_goldens(tmp_path, hostsim_lib, name):
    """The device walker's source, built for the host, against the unmodified reference's files on the rare-branch cases."""
    case = helpers.load_golden(name, str(tmp_path))
    odb = helpers.OracleDB(case["db"])
    seqs = jo.read_fasta(case["fasta"])
    names = [n for n, _ in seqs]
    for use_events in (0, 300):
        st, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], use_events)
        assert st == 0
        for ite in range(case["passes"]):
            rows = [[names[int(f[1])], int(f[2]), f[3], f[4]] for f in (l.split("\t") for l in txt.splitlines()) if f[0] == str(ite)]
            assert jo.format_fix_csv(rows) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
        assert jo.format_fasta(list(zip(names, out))) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
        assert ("%d %d\n" % tallies[0]).encode() == helpers.expected_file(case, "0qValCalcHelper.csv")
        assert ("%d %d\n" % tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])


@pytest.mark.parametrize("seed", range(24))
def test_walker_source_fuzz(tmp_path, hostsim_lib, seed):
    """Random dirty genomes (N runs, lower case, IUPAC, repeats, heterozygosity, clustered errors)."""
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    recs, path, thr = helpers.write_fuzz_db(tmp_path, seed, k)
    odb = helpers.OracleDB(path)
    res, _ = helpers.oracle_polish(odb, recs, k, thr, 2)
    for use_events in (0, 1, 64, 512):       # >= 64: split into segments at anchors, one per that many positions
        st, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, recs, k, thr, 2, use_events)
        if res is None:
            assert st == -7          # JB_E_REFERENCE_CRASH
            continue
        assert st == 0
        assert txt == helpers.rows_text(res.passes, [n for n, _ in recs])
        assert out == [s for _, s in res.seqs]
        assert tallies == [(p.wrong, p.total) for p in res.passes]


@pytest.mark.parametrize("seed", range(100, 110))
def test_segment_seams_are_exact(tmp_path, hostsim_lib, seed, monkeypatch):
    """Splitting a contig at anchors and walking the segments independently must not change a byte:
    compare against the same walker run sequentially over bigger fuzz genomes (many seams), with the
    default anchor margin and with margin 0, where neighbouring fixes do collide with seams and the
    walkers' run-time seam checks must catch every collision (the contig is then re-walked whole)."""
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    recs, path, thr = helpers.write_fuzz_db(tmp_path, seed, k, big=True)
    odb = helpers.OracleDB(path)
    base = helpers.hostsim_polish(hostsim_lib, odb, recs, k, thr, 2, 1)
    if seed < 103:                               # and the sequential walker against the oracle itself
        res, _ = helpers.oracle_polish(odb, recs, k, thr, 2)
        if res is not None:
            assert base[0] == 0 and base[2] == [s for _, s in res.seqs]
            assert base[1] == helpers.rows_text(res.passes, [n for n, _ in recs])
    total = {"segments": 0, "reruns": 0}
    for margin, keep in ((None, None), ("0", None), (None, "0")):
        for var, val in (("JB_ANCHOR_MARGIN", margin), ("JB_KEEP_MARGIN", keep)):
            if val is None:
                monkeypatch.delenv(var, raising=False)
            else:
                monkeypatch.setenv(var, val)      # 0: seams collide with fixes / walkers need bases behind their window
        for block in (64, 1024):
            st = {}
            got = helpers.hostsim_polish(hostsim_lib, odb, recs, k, thr, 2, block, stats=st)
            assert got == base, (margin, block)
            for key in total:
                total[key] += st[key]
    assert total["segments"] > 3 * len(recs) * 4          # the contigs really were split


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_jasper_py_host_side_writes_reference_bytes(golden_cases, name, tmp_path):
    """The host half of the drop-in jasper.py (FASTA parse, csv / FASTA writers) against the golden files,
    fed with the oracle's result in place of the GPU's."""
    from jasper_b200 import jasper as host
    case = golden_cases[name]
    seq_dict = host.parse_fasta(case["fasta"])
    assert list(seq_dict.items()) == jo.read_fasta(case["fasta"])
    odb = helpers.OracleDB(case["db"])
    res, _ = helpers.oracle_polish(odb, list(seq_dict.items()), case["k"], case["threshold"], case["passes"])
    for ite in range(case["passes"]):
        out = str(tmp_path / host._prefixed("asm.fa.fix.csv", ite))
        host.write_fix_csv(out, res.passes[ite].rows)
        assert open(out, "rb").read() == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    fa = str(tmp_path / "x.fa")
    host.write_fasta(fa, res.seqs)
    assert open(fa, "rb").read() == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))


def test_jasper_py_parse_fasta_quirks(tmp_path):
    from jasper_b200 import jasper as host
    p = tmp_path / "q.fa"
    p.write_text("ACGT\n>a desc here\nAC GT\r\nTT\n>b\n\n>a\nGG\n>\nCC\n")
    d = host.parse_fasta(str(p))
    assert list(d.items()) == [("a", "GG"), ("b", ""), ("", "CC")]       # dict semantics: first position, last value
    assert host._prefixed("x.csv", 3) == "_iter3_x.csv" and host._prefixed("d/x.csv", 0) == "d_iter0_x.csv"


@pytest.mark.parametrize("k", [8, 12, 20, 25, 31, 32, 33, 37, 48])
def test_packed_keys_equal_the_byte_loop(hostsim_lib, k):
    """jb_cand.cuh cuts candidate k-mers out of 2-bit packed strings; every window of every candidate kind must
    be the k-mer the byte loop (MerDNA semantics) builds, and the fast path must actually serve most windows."""
    fast = C.c_longlong()
    bad = hostsim_lib.hostsim_packed_selfcheck(1234 + k, k, C.byref(fast))
    assert bad == 0
    assert fast.value > 5000


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_fast_path_is_taken_and_changes_nothing(golden_cases, hostsim_lib, name, monkeypatch):
    """Events in untouched territory take bounds, fix and the following scan step from verdicts (jb_verdict.cuh: first
    stage for single-base fixes, second stage = the walker's own dispatch run dry); with the validity bytes
    withheld the walker runs its generic code for everything.  Both must give the same bytes, and the fast path must
    really carry most events."""
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    seqs = jo.read_fasta(case["fasta"])
    for use_events in (1, 300):
        st_fast, st_slow = {}, {}
        monkeypatch.delenv("JB_SIM_NO_OKB", raising=False)
        fast = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], use_events, stats=st_fast)
        monkeypatch.setenv("JB_SIM_NO_OKB", "1")
        slow = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], use_events, stats=st_slow)
        monkeypatch.delenv("JB_SIM_NO_OKB", raising=False)
        assert fast == slow
        assert st_slow["fast"] == 0 and st_slow["events"] == st_fast["events"]
        st_one = {}
        monkeypatch.setenv("JB_SIM_NO_VERDICTS2", "1")           # first-stage verdicts only
        one = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], use_events, stats=st_one)
        monkeypatch.delenv("JB_SIM_NO_VERDICTS2", raising=False)
        assert one == fast and st_one["v2"] == 0 and st_fast["v2"] > 0
        print(name, use_events, st_fast)
        if name != "k31_dirty":
            assert st_fast["fast"] * 2 > st_fast["events"], st_fast
            assert st_fast["post"] * 4 > st_fast["events"], st_fast


@pytest.mark.parametrize("name", ["k31_dirty", "k25_mix"])
def test_fix_rows_assembled_on_several_threads(golden_cases, hostsim_lib, name, monkeypatch):
    """The fix log becomes csv rows on up to four host threads (index, convert, splice); small inputs use one.  Forcing
    three and five threads must give the same rows (hostsim also checks them against the sorting path)."""
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    seqs = jo.read_fasta(case["fasta"])
    base = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], 300)
    for n in ("3", "5"):
        monkeypatch.setenv("JB_ROW_THREADS", n)
        assert helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], 300) == base
    monkeypatch.delenv("JB_ROW_THREADS", raising=False)


# ----------------------------------------------------------------------------- compact table arithmetic (CPU twin)
@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
@pytest.mark.parametrize("load", [0.0, 0.9])
def test_compact_table_arithmetic_on_golden_dbs(golden_cases, name, load):
    """Sequential placement + the probe rule + the low-block inversion, on the real `jellyfish count` files."""
    sim = helpers.table_sim_lib()
    st, msg = helpers.table_sim_check(sim, golden_cases[name]["db"], load=load)
    if name == "k37_mix":                       # 74-bit keys in a 2^17 hash: the remainder alone takes 57 bits
        assert st["layout"] == 0 and "no room" in msg
        return
    assert st["layout"] == 1 and st["records"] > 0
    assert (st["n_stash"] == 0 if st["cnt_bits"] >= 20 and load == 0.0 else st["n_stash"] < st["records"] // 100)      # (k31_dirty: a 9-bit count field, 22 counts beyond it)
    assert st["max_disp_lines"] <= (3 if load == 0.0 else 12)
    if load == 0.0:
        assert st["lines_per_hit_x1000"] < 1060 and st["lines_per_miss_x1000"] < 1100      # 32-byte sectors read per lookup, default load
    # a 3-bit count field: most records escape to the stash and are still found with their exact counts
    st2, _ = helpers.table_sim_check(sim, golden_cases[name]["db"], load=load, max_cval_bits=3)
    assert st2["n_stash"] > 0


def test_compact_table_tombstones_and_singular_matrix(tmp_path):
    sim = helpers.table_sim_lib()
    p = str(tmp_path / "collide.jf")
    helpers.colliding_db(p, big_counts=50)
    st, _ = helpers.table_sim_check(sim, p)
    assert st["layout"] == 1 and st["max_disp_lines"] > 3 and st["n_stash"] >= 50
    # a matrix whose low block is singular cannot be quotiented: wide layout, with the reason
    rng = np.random.default_rng(4)
    lo = np.unique(rng.integers(0, 1 << 50, 5000, dtype=np.uint64))
    q = str(tmp_path / "singular.jf")
    synth.write_jf(q, lo, np.zeros(len(lo), dtype=np.uint64), np.ones(len(lo), dtype=np.uint64), 25, jellyfish_like=False)
    st, msg = helpers.table_sim_check(sim, q)
    assert st["layout"] == 0 and "singular" in msg


def test_long_clean_stretches_hop_by_the_level2_summary(tmp_path, hostsim_lib, monkeypatch):
    """A 0.8 Mb contig with a handful of errors: the walker leaps the clean megabases through the 1-per-32768 summary
    (fast_hop); with and without it, and against the oracle, the result is the same."""
    rng = np.random.default_rng(77)
    k, thr = 25, 3
    g = synth.random_genome(800_000, rng)
    lo, hi = synth.canonical_keys(g, k)
    lo, hi, cnt = synth.unique_keys(lo, hi, np.full(len(lo), 20, dtype=np.int64))
    path = os.path.join(str(tmp_path), "clean.jf")
    synth.write_jf(path, lo, hi, cnt, k, seed=5)
    a = g.copy()
    for at in (40_011, 41_500, 500_007, 799_950):          # substitutions: two near each other, hundreds of kilobases between the rest
        a[at] = (a[at] + 1) % 4
    a = np.delete(a, [230_123])                                     # and one deletion
    recs = [("long", synth.codes_to_bytes(a).decode()), ("short", synth.codes_to_bytes(g[:5000]).decode())]
    odb = helpers.OracleDB(path)
    res, _ = helpers.oracle_polish(odb, recs, k, thr, 2)
    assert res is not None and res.seqs[0][1] == synth.codes_to_bytes(g).decode()
    got = {}
    for no_l2 in ("", "1"):
        if no_l2:
            monkeypatch.setenv("JB_SIM_NO_L2", "1")
        else:
            monkeypatch.delenv("JB_SIM_NO_L2", raising=False)
        for use_events in (1, 4096):
            st, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, recs, k, thr, 2, use_events)
            assert st == 0
            assert txt == helpers.rows_text(res.passes, [n for n, _ in recs])
            assert out == [s for _, s in res.seqs]
            assert tallies == [(p.wrong, p.total) for p in res.passes]
            got[(no_l2, use_events)] = (txt, out, tallies)
    assert len(set(map(repr, got.values()))) == 1


def test_deep_bridge_outgrows_the_standard_trie(tmp_path, hostsim_lib):
    """tests/golden/branches/deep_bridge (bytes of the unmodified reference): thousands of live paths for hundreds of rounds.
    The walker source runs out of the standard 1 M-node trie, reports kTrieFull, and the re-walk with the grown pool gives the
    reference's bytes."""
    case = helpers.load_golden(helpers.DEEP_BRIDGE, str(tmp_path))
    seqs = jo.read_fasta(case["fasta"])
    odb = helpers.OracleDB(case["db"])
    names = [n for n, _ in seqs]
    st = {}
    status, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, seqs, case["k"], case["threshold"], case["passes"], 300, stats=st)
    assert status == 0 and st["pool_growths"] >= 1
    for ite in range(case["passes"]):
        rows = [[names[int(f[1])], int(f[2]), f[3], f[4]] for f in (l.split("\t") for l in txt.splitlines()) if f[0] == str(ite)]
        assert jo.format_fix_csv(rows) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(list(zip(names, out))) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    assert ("%d %d\n" % tallies[0]).encode() == helpers.expected_file(case, "0qValCalcHelper.csv")
    assert ("%d %d\n" % tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])

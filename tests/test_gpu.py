"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle."""
import os
import sys

import numpy as np
import pytest

import helpers
import jasper_oracle as jo
from jasper_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["compact", "wide"])
def dbs(golden_cases, request):
    """Every golden database resident twice over the module's tests: in the compact layout (8-byte quotient slots, sixteen
    per 128-byte line; k37_mix's 2^17 hash cannot be quotiented into 64 bits and stays wide) and in the wide fallback."""
    old = os.environ.get("JB_TABLE_LAYOUT")
    os.environ["JB_TABLE_LAYOUT"] = request.param
    out = {}
    try:
        for name, case in golden_cases.items():
            out[name] = capi.QueryMerFile(case["db"])
            want = capi.LAYOUT_COMPACT if (request.param == "compact" and name != "k37_mix") else capi.LAYOUT_WIDE
            assert out[name].info["layout"] == want
    finally:
        if old is None:
            del os.environ["JB_TABLE_LAYOUT"]
        else:
            os.environ["JB_TABLE_LAYOUT"] = old
    yield out
    for q in out.values():
        q.close()


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_table_holds_exactly_the_file_records(golden_cases, dbs, name):
    """load -> dump == the record list of the file, in file order (the loader preserves (pos,key) order)."""
    odb = helpers.OracleDB(golden_cases[name]["db"])
    lo, hi, cnt = odb.records()
    qf = dbs[name]
    assert qf.k == odb.k and qf.info["n_records"] == odb.n_records
    glo, ghi, gcnt = qf.dump()
    assert np.array_equal(glo, lo) and np.array_equal(ghi, hi) and np.array_equal(gcnt.astype(np.uint64), cnt & 0xFFFFFFFF)


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_histogram_matches_jellyfish_histo(golden_cases, dbs, name):
    rows = [tuple(int(x) for x in l.split()) for l in open(os.path.join(golden_cases[name]["dir"], "histo.csv"))]
    assert dbs[name].histogram() == rows


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_lookup_matches_binding_vectors_and_oracle(golden_cases, dbs, name):
    case = golden_cases[name]
    qf = dbs[name]
    strings = [s for s, _, _ in case["lookups"]]
    assert qf.lookup(strings) == [c for _, _, c in case["lookups"]]
    odb = helpers.OracleDB(case["db"])
    rng = np.random.default_rng(1)
    seq = "".join(s for _, s in jo.read_fasta(case["fasta"]))
    alphabet = "ACGTacgtNnRY-"
    probes = []
    for _ in range(5000):
        if rng.random() < 0.6:
            at = int(rng.integers(0, len(seq) - 100)); probes.append(seq[at:at + int(rng.integers(0, 2 * qf.k))])
        else:
            probes.append("".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), int(rng.integers(0, 60)))))
    assert qf.lookup(probes) == [odb.count(s) for s in probes]
    assert qf[capi.MerDNA(probes[0]).get_canonical()] == odb.count(probes[0])


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
@pytest.mark.parametrize("index", ["index", "table"])
def test_count_positions_matches_oracle(golden_cases, dbs, name, index, monkeypatch):
    """k_scan against the oracle for every window, through the scan index (records grouped by minimizer, jb_index.cuh) and,
    with JB_SCAN_NO_INDEX, through the hash table."""
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    qf = dbs[name]
    if index == "index":                                  # (the index is an option: JB_SCAN_INDEX at open time)
        monkeypatch.setenv("JB_SCAN_INDEX", "1")
        qf = capi.QueryMerFile(case["db"])
        assert qf.info["index_bytes"] > 0
    else:
        assert qf.info["index_bytes"] == 0
    try:
        for _, s in jo.read_fasta(case["fasta"]):
            got = qf.count_positions(s)
            assert np.array_equal(got, odb.count_positions(s))
        assert np.array_equal(qf.count_positions(""), odb.count_positions(""))
    finally:
        if index == "index":
            qf.close()


@pytest.mark.parametrize("seed", [3, 7, 12, 21, 30])
@pytest.mark.parametrize("layout", ["compact", "wide"])
def test_scan_index_equals_the_table_on_dirty_input(tmp_path, monkeypatch, seed, layout):
    """Fuzz genomes (N runs, lower case, IUPAC, repeats; k = 20 ... 48): the counts of every window through the index equal
    those through the hash table, and the polished result is the same, from a compact and from a wide table."""
    monkeypatch.setenv("JB_TABLE_LAYOUT", layout)
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    if k < 15:
        pytest.skip("no scan index below k = 15")
    recs, path, thr = helpers.write_fuzz_db(tmp_path, seed, k, big=True)
    monkeypatch.setenv("JB_SCAN_INDEX", "1")
    qf = capi.QueryMerFile(path)
    try:
        assert qf.info["index_bytes"] > 0
        with_index = [qf.count_positions(s) for _, s in recs]
        pol_index = qf.polish(recs, k, thr, 2)
        monkeypatch.setenv("JB_SCAN_NO_INDEX", "1")
        without = [qf.count_positions(s) for _, s in recs]
        pol_table = qf.polish(recs, k, thr, 2)
        for a, b in zip(with_index, without):
            assert np.array_equal(a, b)
        assert pol_index.status == pol_table.status
        assert pol_index.seqs == pol_table.seqs and pol_index.rows == pol_table.rows and pol_index.tallies == pol_table.tallies
    finally:
        qf.close()


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_polish_reproduces_reference_bytes(golden_cases, dbs, name):
    """fix.csv per pass, polished FASTA and tallies: byte-identical to the unmodified jasper.py run."""
    case = golden_cases[name]
    seqs = jo.read_fasta(case["fasta"])
    res = dbs[name].polish(seqs, case["k"], case["threshold"], case["passes"])
    assert res.status == 0
    for ite in range(case["passes"]):
        assert jo.format_fix_csv(res.rows[ite]) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    assert ("%d %d\n" % res.tallies[0]).encode() == helpers.expected_file(case, "0qValCalcHelper.csv")
    assert ("%d %d\n" % res.tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])
    assert res.counters["kernel_launches"] >= 2 * (case["passes"] + 1)


@pytest.mark.parametrize("name", helpers.BRANCH_GOLDENS)
@pytest.mark.parametrize("layout", ["compact", "wide"])
def test_polish_reproduces_reference_bytes_on_rare_branches(tmp_path, monkeypatch, name, layout):
    """tests/golden/branches/*: the unmodified reference on fuzz cases that take fixdiploid's "s" side, grow_single, unfixed_k and the
    rolling threshold (every branch >= 5 times in all, tests/test_oracle.py checks the counts): same bytes from the CUDA path,
    through both table layouts, with and without the verdict kernels."""
    monkeypatch.setenv("JB_TABLE_LAYOUT", layout)
    case = helpers.load_golden(name, str(tmp_path))
    seqs = jo.read_fasta(case["fasta"])
    qf = capi.QueryMerFile(case["db"])
    try:
        for no_verdicts in (None, "1"):
            if no_verdicts:
                monkeypatch.setenv("JB_NO_VERDICTS", no_verdicts)
            res = qf.polish(seqs, case["k"], case["threshold"], case["passes"])
            assert res.status == 0
            for ite in range(case["passes"]):
                assert jo.format_fix_csv(res.rows[ite]) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
            assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
            assert ("%d %d\n" % res.tallies[0]).encode() == helpers.expected_file(case, "0qValCalcHelper.csv")
            assert ("%d %d\n" % res.tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])
    finally:
        monkeypatch.delenv("JB_NO_VERDICTS", raising=False)
        qf.close()


def test_deep_bridge_rewalks_with_the_grown_pool(tmp_path, monkeypatch):
    """tests/golden/branches/deep_bridge (bytes of the unmodified reference): one de-Bruijn search of more than a million path
    extensions.  A standard bridging slot (256 K trie nodes) overflows, the walker reports kTrieFull, the library grows the pool and
    re-walks that contig alone -- same bytes as the reference; with the grown pool capped too, the call fails with JB_E_SCRATCH and
    says which contig."""
    case = helpers.load_golden(helpers.DEEP_BRIDGE, str(tmp_path))
    seqs = jo.read_fasta(case["fasta"])
    qf = capi.QueryMerFile(case["db"])
    try:
        before = capi.counter("pool_growths")
        res = qf.polish(seqs, case["k"], case["threshold"], case["passes"])
        assert res.status == 0
        assert capi.counter("pool_growths") > before
        assert res.counters["segment_reruns"] >= 1
        for ite in range(case["passes"]):
            assert jo.format_fix_csv(res.rows[ite]) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
        assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
        assert ("%d %d\n" % res.tallies[0]).encode() == helpers.expected_file(case, "0qValCalcHelper.csv")
        assert ("%d %d\n" % res.tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])
        monkeypatch.setenv("JB_BIG_TRIE_NODES", str((1 << 20) + 4096))
        with pytest.raises(capi.JasperError) as ei:
            qf.polish(seqs, case["k"], case["threshold"], case["passes"])
        assert ei.value.code == -8 and "contig 0" in str(ei.value)          # JB_E_SCRATCH
    finally:
        qf.close()


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
@pytest.mark.parametrize("seg_block,margin", [("0", None), ("64", None), ("256", "0"), ("2048", None)])
def test_polish_golden_is_independent_of_segmentation(golden_cases, dbs, name, seg_block, margin, monkeypatch):
    """The same bytes whether contigs are walked whole (JB_SEG_BLOCK=0) or split into segments at anchors."""
    monkeypatch.setenv("JB_SEG_BLOCK", seg_block)
    if margin is not None:
        monkeypatch.setenv("JB_ANCHOR_MARGIN", margin)
    case = golden_cases[name]
    seqs = jo.read_fasta(case["fasta"])
    res = dbs[name].polish(seqs, case["k"], case["threshold"], case["passes"])
    assert res.status == 0
    for ite in range(case["passes"]):
        assert jo.format_fix_csv(res.rows[ite]) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    assert ("%d %d\n" % res.tallies[-1]).encode() == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])
    if seg_block != "0":
        assert res.counters["segments"] > (case["passes"] + 1) * len(seqs)


@pytest.mark.parametrize("seed", range(100, 112))
def test_polish_big_fuzz_segmented_matches_hostsim_and_oracle(tmp_path, hostsim_lib, seed, monkeypatch):
    """Bigger dirty genomes, many seams: device result == sequential host build of the walker (== oracle for
    the first seeds), for the default anchor margin and for margin 0 (seam re-runs)."""
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    recs, path, thr = helpers.write_fuzz_db(tmp_path, seed, k, big=True)
    odb = helpers.OracleDB(path)
    st, txt, out, tallies = helpers.hostsim_polish(hostsim_lib, odb, recs, k, thr, 2, 1)
    names = [n for n, _ in recs]
    if seed < 104:
        want, _ = helpers.oracle_polish(odb, recs, k, thr, 2)
        if want is not None:
            assert st == 0 and out == [s for _, s in want.seqs] and txt == helpers.rows_text(want.passes, names)
    qf = capi.QueryMerFile(path)
    try:
        reruns = 0
        for seg_block, margin, keep in (("128", None, None), ("128", "0", None), ("1024", "0", "0"), ("100", None, "0")):
            monkeypatch.setenv("JB_SEG_BLOCK", seg_block)
            for var, val in (("JB_ANCHOR_MARGIN", margin), ("JB_KEEP_MARGIN", keep)):
                if val is None:
                    monkeypatch.delenv(var, raising=False)
                else:
                    monkeypatch.setenv(var, val)
            res = qf.polish(recs, k, thr, 2)
            assert res.status == (0 if st == 0 else st)
            if st == 0:
                assert helpers.product_rows_text(res.rows, names) == txt
                assert [s for _, s in res.seqs] == out
                assert res.tallies == tallies
            reruns += res.counters["segment_reruns"]
    finally:
        qf.close()


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_device_side_placement_changes_nothing(golden_cases, dbs, name, monkeypatch):
    """k_place puts the segments' new bases on the device and k_concat / the recount are queued behind it before the host has
    seen the walkers' results; with JB_NO_SPECULATION the host places and launches them itself (the path a re-walk takes).
    Same result either way, also when every seam is forced to fail (margin 0: the speculative launches then find nothing to
    do and the host's placement counts)."""
    case = golden_cases[name]
    seqs = jo.read_fasta(case["fasta"])
    fast = dbs[name].polish(seqs, case["k"], case["threshold"], 3)
    assert fast.status == 0 and fast.counters["speculative_passes"] >= 1
    monkeypatch.setenv("JB_NO_SPECULATION", "1")
    slow = dbs[name].polish(seqs, case["k"], case["threshold"], 3)
    monkeypatch.delenv("JB_NO_SPECULATION")
    assert slow.counters["speculative_passes"] == 0
    monkeypatch.setenv("JB_SEG_BLOCK", "256"); monkeypatch.setenv("JB_ANCHOR_MARGIN", "0")
    seams = dbs[name].polish(seqs, case["k"], case["threshold"], 3)
    for other in (slow, seams):
        assert other.status == 0
        assert other.seqs == fast.seqs and other.tallies == fast.tallies and other.rows == fast.rows


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_incremental_recount_equals_full_scan(golden_cases, dbs, name, monkeypatch):
    """Passes 2.. inherit window counts through k_recount: (a) the device checks every inherited count against a
    full scan of the new sequence (JB_VERIFY_RECOUNT), (b) forcing full rescans gives the same bytes."""
    case = golden_cases[name]
    seqs = jo.read_fasta(case["fasta"])
    monkeypatch.setenv("JB_VERIFY_RECOUNT", "1")
    a = dbs[name].polish(seqs, case["k"], case["threshold"], 3)
    assert a.status == 0 and a.counters["recount_probes"] > 0
    monkeypatch.delenv("JB_VERIFY_RECOUNT")
    monkeypatch.setenv("JB_FULL_RESCAN", "1")
    b = dbs[name].polish(seqs, case["k"], case["threshold"], 3)
    assert b.counters["recount_probes"] == 0
    assert (a.rows, a.seqs, a.tallies) == (b.rows, b.seqs, b.tallies)


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_verdicts_change_nothing(golden_cases, dbs, name, monkeypatch):
    """k_verdicts decides the isolated errors in bulk and the walkers consume its verdicts; with JB_NO_VERDICTS the
    walkers decide everything themselves.  Same bytes either way, and the verdicts really carry most events."""
    case = golden_cases[name]
    seqs = jo.read_fasta(case["fasta"])
    a = dbs[name].polish(seqs, case["k"], case["threshold"], case["passes"])
    assert a.status == 0 and a.counters["verdicts"] > 0
    assert a.counters["fast_events"] * 2 > a.counters["events"], a.counters
    assert a.counters["verdict2_events"] > 0, a.counters
    monkeypatch.setenv("JB_NO_VERDICTS2", "1")               # first stage only: the walkers run the dispatch for the rest
    c = dbs[name].polish(seqs, case["k"], case["threshold"], case["passes"])
    assert c.counters["verdict2_events"] == 0 and c.counters["fast_events"] == a.counters["fast_events"]
    assert (a.rows, a.seqs, a.tallies) == (c.rows, c.seqs, c.tallies)
    monkeypatch.setenv("JB_NO_VERDICTS", "1")
    b = dbs[name].polish(seqs, case["k"], case["threshold"], case["passes"])
    assert b.counters["verdicts"] == 0 and b.counters["fast_events"] == 0
    assert (a.rows, a.seqs, a.tallies) == (b.rows, b.seqs, b.tallies)


@pytest.mark.parametrize("seed", range(40))
def test_polish_fuzz_matches_oracle(tmp_path, seed, monkeypatch):
    monkeypatch.setenv("JB_VERIFY_RECOUNT", "1")
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    recs, path, thr = helpers.write_fuzz_db(tmp_path, seed, k)
    odb = helpers.OracleDB(path)
    want, _ = helpers.oracle_polish(odb, recs, k, thr, 2)
    qf = capi.QueryMerFile(path)
    try:
        if want is None:
            res = qf.polish(recs, k, thr, 2)
            assert res.status == -7
            return
        res = qf.polish(recs, k, thr, 2)
        names = [n for n, _ in recs]
        assert res.status == 0
        assert helpers.product_rows_text(res.rows, names) == helpers.rows_text(want.passes, names)
        assert [s for _, s in res.seqs] == [s for _, s in want.seqs]
        assert res.tallies == [(p.wrong, p.total) for p in want.passes]
    finally:
        qf.close()


def _mid_case(tmp_path, n, k, seed, err=0.001):
    rng = np.random.default_rng(seed)
    g = synth.random_genome(n, rng, homopolymers=n // 2000)
    a = synth.inject_errors(g, err, rng)
    lo, hi = synth.canonical_keys(g, k)
    lo, hi, mult = synth.unique_keys(lo, hi)
    cnt = np.maximum(1, rng.poisson(22.8, len(lo))) * mult
    # singleton error k-mers: random keys (absent from the genome with overwhelming probability)
    ne = len(lo) // 2
    elo = rng.integers(0, 1 << 62, ne, dtype=np.uint64)
    ehi = rng.integers(0, 1 << max(1, 2 * k - 64), ne, dtype=np.uint64) if 2 * k > 64 else np.zeros(ne, np.uint64)
    if 2 * k < 64:
        elo &= np.uint64((1 << (2 * k)) - 1)
    lo2, hi2, c2 = synth.unique_keys(np.concatenate([lo, elo]), np.concatenate([hi, ehi]), np.concatenate([cnt, np.ones(ne, np.int64)]))
    path = str(tmp_path / ("mid_%d_%d.jf" % (n, k)))
    synth.write_jf(path, lo2, hi2, c2, k, seed=seed)
    return synth.codes_to_bytes(a).decode(), path


def test_polish_mid_size_matches_oracle(tmp_path):
    """1 Mb contig + small ones, k=37, 3 passes: sequences, rows and tallies equal the oracle's."""
    k = 37
    s, path = _mid_case(tmp_path, 1_000_000, k, 77, err=0.002)
    recs = [("big", s), ("part", s[5000:90000]), ("short", s[:30]), ("empty", "")]
    odb = helpers.OracleDB(path)
    want, _ = helpers.oracle_polish(odb, recs, k, 3, 3)
    qf = capi.QueryMerFile(path)
    res = qf.polish(recs, k, 3, 3)
    names = [n for n, _ in recs]
    assert res.status == 0 and want is not None
    assert helpers.product_rows_text(res.rows, names) == helpers.rows_text(want.passes, names)
    assert [x for _, x in res.seqs] == [x for _, x in want.seqs]
    assert res.tallies == [(p.wrong, p.total) for p in want.passes]
    # size-independent properties: a polished assembly is a fixed point of the count-only pass, and
    # polishing again with fix=False changes nothing
    again = qf.polish(res.seqs, k, 3, 0)
    assert [x for _, x in again.seqs] == [x for _, x in res.seqs]
    assert again.tallies[0] == res.tallies[-1]
    qf.close()


def test_count_positions_large_properties(tmp_path):
    """8 Mb: every genome window is found with its exact count; reverse-complement symmetry."""
    k = 37
    rng = np.random.default_rng(5)
    g = synth.random_genome(8_000_000, rng)
    lo, hi = synth.canonical_keys(g, k)
    ulo, uhi, mult = synth.unique_keys(lo, hi)
    path = str(tmp_path / "big.jf")
    synth.write_jf(path, ulo, uhi, mult * 3, k, seed=1)
    qf = capi.QueryMerFile(path)
    s = synth.codes_to_bytes(g).decode()
    c = qf.count_positions(s)
    assert c.shape[0] == len(s) + 1
    assert np.all(c[:len(s) - k + 1] >= 3) and np.all(c[:len(s) - k + 1] % 3 == 0)
    rc = synth.codes_to_bytes((3 - g)[::-1]).decode()
    c2 = qf.count_positions(rc)
    n = len(s) - k + 1
    assert np.array_equal(c[:n], c2[:n][::-1])
    assert qf.info["max_displacement"] < 4096
    qf.close()


@pytest.mark.parametrize("cval_bits", [None, "3"])
def test_compact_table_stash_tombstones_and_escapes(tmp_path, monkeypatch, cval_bits):
    """Records a compact probe cannot reach (displaced beyond kMaxHops sectors by a pile-up on one hash position) and counts
    beyond the slot's count field live in the stash: lookups, dump and histogram still equal the file."""
    if cval_bits:
        monkeypatch.setenv("JB_COMPACT_CVAL_BITS", cval_bits)
        monkeypatch.setenv("JB_STASH_FRAC", "1")             # (almost every record escapes: the default stash, records / 64, would send the loader to the wide layout)
    path = str(tmp_path / "collide.jf")
    helpers.colliding_db(path, big_counts=50)
    odb = helpers.OracleDB(path)
    lo, hi, cnt = odb.records()
    qf = capi.QueryMerFile(path)
    try:
        assert qf.info["layout"] == capi.LAYOUT_COMPACT and qf.info["n_stash"] >= 50
        assert qf.info["max_displacement"] > 4 * 8            # some record sits more than kMaxHops sectors from home
        glo, ghi, gcnt = qf.dump()
        assert np.array_equal(glo, lo) and np.array_equal(ghi, hi) and np.array_equal(gcnt.astype(np.uint64), cnt & 0xFFFFFFFF)
        k = qf.k
        strings = ["".join("ACGT"[(int(x) >> (2 * (k - 1 - i))) & 3] for i in range(k)) for x in lo[::7]]
        assert qf.lookup(strings) == [odb.count(s) for s in strings]
        rng = np.random.default_rng(9)
        absent = ["".join("ACGT"[int(b)] for b in rng.integers(0, 4, k)) for _ in range(3000)]
        assert qf.lookup(absent) == [odb.count(s) for s in absent]
        seq = "".join(strings[:400])
        assert np.array_equal(qf.count_positions(seq), odb.count_positions(seq))
        hist = dict(qf.histogram(high=250))
        want = {}
        for c in (cnt & 0xFFFFFFFF):
            c = min(int(c), 251)
            want[c] = want.get(c, 0) + 1
        assert hist == want
    finally:
        qf.close()


def test_compact_all_zero_pattern_is_not_shadowed_by_empty_slots(tmp_path):
    """k = 8: the remainder is (almost) empty, so a key whose remainder and tag are 0 has the all-zero pattern that every EMPTY slot
    also shows (r2b: empty slots behind the record used to win).  Every one of the 4^8 k-mers against the CPU oracle."""
    k = 8
    rng = np.random.default_rng(21)
    lo = np.unique(rng.integers(0, 1 << (2 * k), 26000, dtype=np.uint64))
    cnt = rng.integers(1, 300, len(lo)).astype(np.uint64)
    path = str(tmp_path / "k8.jf")
    synth.write_jf(path, lo, np.zeros(len(lo), dtype=np.uint64), cnt, k, seed=5)
    odb = helpers.OracleDB(path)
    qf = capi.QueryMerFile(path)
    try:
        assert qf.info["layout"] == capi.LAYOUT_COMPACT and qf.info["remainder_bits"] <= 1
        strings = ["".join("ACGT"[(x >> (2 * (k - 1 - i))) & 3] for i in range(k)) for x in range(1 << (2 * k))]
        got = qf.lookup(strings)
        want = [odb.count(s) for s in strings]
        assert got == want
        seq = "".join(strings[:3000])
        assert np.array_equal(qf.count_positions(seq), odb.count_positions(seq))
    finally:
        qf.close()


def test_singular_matrix_falls_back_to_the_wide_layout(tmp_path):
    rng = np.random.default_rng(4)
    lo = np.unique(rng.integers(0, 1 << 50, 5000, dtype=np.uint64))
    path = str(tmp_path / "singular.jf")
    synth.write_jf(path, lo, np.zeros(len(lo), dtype=np.uint64), np.arange(1, len(lo) + 1, dtype=np.uint64), 25, jellyfish_like=False)
    qf = capi.QueryMerFile(path)
    try:
        assert qf.info["layout"] == capi.LAYOUT_WIDE
        glo, _, gcnt = qf.dump()
        odb = helpers.OracleDB(path)
        rlo, _, rcnt = odb.records()
        assert np.array_equal(glo, rlo) and np.array_equal(gcnt.astype(np.uint64), rcnt)
    finally:
        qf.close()


def test_compact_k37_table_at_bench_density(tmp_path):
    """k = 37 needs >= 2^21 lines before the 74-bit key quotients into an 8-byte slot: a 26 Mb genome's database
    (what bench.py loads, smaller) must come up compact, ~16 bytes per record at the default load 0.5, and count every window like the CPU oracle."""
    import torch
    from jasper_b200 import synth_gpu
    w = synth_gpu.make_workload(str(tmp_path), 26_000_000, 37, seed=3, err_kmers_per_bp=0.2)
    qf = capi.QueryMerFile(w["db"])
    try:
        info = qf.info
        assert info["layout"] == capi.LAYOUT_COMPACT and info["slot_bytes"] == 8 and info["n_stash"] == 0
        assert 0.45 < info["load_factor"] < 0.55 and info["bytes_per_record"] < 16.5
        assert info["max_displacement"] <= 4 * 16
        g = w["genome"]
        odb = helpers.OracleDB(w["db"])
        rng = np.random.default_rng(2)
        for at in rng.integers(0, len(g) - 300_000, 3):
            s = synth.codes_to_bytes(g[int(at):int(at) + 200_000]).decode()
            s = s[:1000] + "N" + s[1001:5000] + "acgtn" + s[5005:]               # invalid windows, lowercase
            got = qf.count_positions(s)
            assert np.array_equal(got, odb.count_positions(s))
        # mutated windows: mostly absent keys
        s = bytearray(synth.codes_to_bytes(g[1000:101000]))
        for i in range(0, len(s), 41):
            s[i] = ord("ACGT"[("ACGT".index(chr(s[i])) + 1) % 4])
        s = s.decode()
        assert np.array_equal(qf.count_positions(s), odb.count_positions(s))
        sectors = capi.counter("probe_steps")
    finally:
        qf.close()
    del torch


def test_errors_mirror_the_binding(tmp_path, golden_cases):
    with pytest.raises(capi.JasperError) as e:
        capi.QueryMerFile(str(tmp_path / "missing.jf"))
    assert "Can't open file" in str(e.value)
    bad = tmp_path / "bad.jf"
    bad.write_bytes(b"000000031{\"format\":\"bloomcounter\"}\0\0\0\0\0\0")
    with pytest.raises(capi.JasperError) as e:
        capi.QueryMerFile(str(bad))
    assert "Unsupported format" in str(e.value) or "format" in str(e.value)
    case = golden_cases["k25_mix"]
    qf = capi.QueryMerFile(case["db"])
    with pytest.raises(capi.JasperError) as e:
        qf.polish([("a", "ACGT" * 50)], 31, 3, 1)      # --ksize differs from the DB's k
    assert e.value.code == -6
    qf.close()


def test_gather_roofline_probe(dbs):
    gbps, gps = dbs["k37_mix"].gather_bench(n_probes=1 << 24, sector_bytes=32, iters=2)
    assert gbps > 0 and gps > 0


@pytest.mark.parametrize("name", ["k37_mix", "k31_dirty"])
def test_jasper_py_cli_is_drop_in(golden_cases, name, tmp_path):
    """jasper_b200/jasper.py run exactly as src/jasper.sh:207 runs src/jasper.py writes the reference's bytes."""
    import shutil
    import subprocess
    import sys
    case = golden_cases[name]
    shutil.copy(case["fasta"], tmp_path / "asm.fa")
    script = os.path.join(helpers.ROOT, "jasper_b200", "jasper.py")
    cmd = [sys.executable, script, "--db", case["db"], "--query", "asm.fa", "--ksize", str(case["k"]), "-p", str(case["passes"]),
           "--fix", "--fout", "asm.fa.fix.csv", "-ff", "asm.fa.fixed.fa.tmp", "--test", "-thre", str(case["threshold"])]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for fn in os.listdir(os.path.join(case["dir"], "expected")):
        assert (tmp_path / fn).read_bytes() == helpers.expected_file(case, fn), fn
    # a wrong k is refused with exit 1 (the reference would silently mis-read k-mers)
    bad = subprocess.run([c if c != str(case["k"]) else str(case["k"] - 2) for c in cmd], cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode == 1


def _n_gpus():
    try:
        return capi.load_library().jb_device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("name", ["k37_mix", "k25_mix"])
def test_partitioned_table_equals_the_whole_one(golden_cases, dbs, name, monkeypatch):
    """SURVEY.md section 8e fallback: the slots of one database spread over two GPUs, foreign parts probed by peer
    loads.  Every handle must answer exactly like the single-GPU table: edge-case lookups, per-position counts,
    histogram, and the whole polishing loop (run from both devices)."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs with peer access")
    case = golden_cases[name]
    whole = dbs[name]
    parts = capi.QueryMerFile.open_partitioned(case["db"], [0, 1])
    try:
        strings = [s for s, _, _ in case["lookups"]]
        seqs = jo.read_fasta(case["fasta"])
        want = whole.polish(seqs, case["k"], case["threshold"], case["passes"])
        for qf in parts:
            assert qf.lookup(strings) == [c for _, _, c in case["lookups"]]
            assert (qf.count_positions(seqs[0][1]) == whole.count_positions(seqs[0][1])).all()
            assert list(qf.histogram(1000)) == list(whole.histogram(1000))
            got = qf.polish(seqs, case["k"], case["threshold"], case["passes"])
            assert (got.rows, got.seqs, got.tallies) == (want.rows, want.seqs, want.tallies)
        with pytest.raises(capi.JasperError):
            parts[0].dump()
    finally:
        for qf in parts:
            qf.close()


@pytest.mark.parametrize("name", ["k37_mix", "k25_mix"])
@pytest.mark.parametrize("layout", ["compact", "wide"])
def test_partitioned_table_arithmetic_on_one_gpu(golden_cases, name, layout, tmp_path, monkeypatch):
    """The same check on a box with ONE GPU (the driver's): three parts of one table as three allocations on device 0
    (JB_PARTS_SAME_DEVICE).  What it covers: the partition boundaries and overlaps, the routing of every probe to the part that
    owns its sector, the loader's stores into the owning part; what it cannot cover: that the loads really cross NVLink."""
    monkeypatch.setenv("JB_PARTS_SAME_DEVICE", "1")
    monkeypatch.setenv("JB_TABLE_LAYOUT", layout)
    case = helpers.load_golden(name, str(tmp_path))
    whole = capi.QueryMerFile(case["db"])
    parts = capi.QueryMerFile.open_partitioned(case["db"], [0, 0, 0])
    try:
        strings = [s for s, _, _ in case["lookups"]]
        seqs = jo.read_fasta(case["fasta"])
        want = whole.polish(seqs, case["k"], case["threshold"], case["passes"])
        for qf in parts:
            assert qf.info["n_parts"] == 3
            assert qf.lookup(strings) == [c for _, _, c in case["lookups"]]
            assert (qf.count_positions(seqs[0][1]) == whole.count_positions(seqs[0][1])).all()
            assert list(qf.histogram(1000)) == list(whole.histogram(1000))
            got = qf.polish(seqs, case["k"], case["threshold"], case["passes"])
            assert (got.rows, got.seqs, got.tallies) == (want.rows, want.seqs, want.tallies)
    finally:
        for qf in parts:
            qf.close()
        whole.close()


def test_bench_contract_on_a_small_genome(tmp_path):
    """bench.py end to end on a 4 Mb genome: one JSON line with the contract's keys, and its built-in parity check
    (the CPU-baseline slices also go through the GPU path and must agree) passes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JB_BENCH_DIR=str(tmp_path / "bench"))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--genome-mb", "4", "--steps", "1", "--warmup", "1", "--cpu-sample-kb", "400"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and "identical to the GPU path" in line["cpu_baseline"]["parity"]


def _part_rank(rank, world, case, tmp, q):
    """One rank of the per-process partitioned table: build the own part, swap handles through files, attach, answer."""
    import time
    try:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        qf = capi.QueryMerFile.open_part(case["db"], rank, rank, world)
        try:
            qf.lookup(["ACGT"])
            q.put((rank, "lookup before attach did not fail"))
            return
        except capi.JasperError:
            pass
        with open(os.path.join(tmp, "h%d.tmp" % rank), "wb") as f:
            f.write(qf.part_handle())
        os.rename(os.path.join(tmp, "h%d.tmp" % rank), os.path.join(tmp, "h%d.bin" % rank))
        handles = []
        for r in range(world):
            path = os.path.join(tmp, "h%d.bin" % r)
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 120:
                    raise RuntimeError("rank %d never published its handle" % r)
                time.sleep(0.01)
            handles.append(open(path, "rb").read())
        qf.attach_parts(handles)
        strings = [s for s, _, _ in case["lookups"]]
        got = qf.lookup(strings)
        seqs = jo.read_fasta(case["fasta"])
        res = qf.polish(seqs, case["k"], case["threshold"], case["passes"])
        fasta = jo.format_fasta(res.seqs)
        # keep the mapping alive until every rank is done with its peers' memory
        open(os.path.join(tmp, "done%d" % rank), "w").close()
        t0 = time.time()
        while not all(os.path.exists(os.path.join(tmp, "done%d" % r)) for r in range(world)) and time.time() - t0 < 120:
            time.sleep(0.01)
        qf.close()
        q.put((rank, got, res.tallies, fasta))
    except BaseException as e:      # noqa: BLE001
        q.put((rank, "error: %r" % (e,)))


def test_partitioned_table_one_process_per_gpu(golden_cases, tmp_path):
    """jb_db_open_part / jb_db_part_handle / jb_db_attach_parts: two processes, one GPU each, every rank holds half of the
    slots and reads the other half through CUDA IPC; both must reproduce the golden lookups and polished FASTA."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs with peer access")
    import multiprocessing as mp
    case = golden_cases["k37_mix"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_part_rank, args=(r, 2, case, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want_fasta = helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    for out in outs:
        assert len(out) == 4, out
        assert out[1] == [c for _, _, c in case["lookups"]]
        assert out[3] == want_fasta

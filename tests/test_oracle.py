"""The oracle against the reference: golden runs of the unmodified jasper.py, Jellyfish's own
known answers, and -- when oracle/_ref is present -- the reference's compiled binding itself."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import helpers
import jasper_oracle as jo
from jasper_b200 import synth


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_polish_oracle_reproduces_reference_run(golden_cases, name):
    """fix.csv bytes per pass, polished FASTA bytes and both tally lines of the unmodified
    /root/reference/src/jasper.py (tests/golden/make_golden.py)."""
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    assert odb.k == case["k"]
    seqs = jo.read_fasta(case["fasta"])
    oracle = jo.PolishOracle(odb.count, case["k"], case["threshold"])
    res = oracle.polish(seqs, case["passes"])
    for ite in range(case["passes"]):
        assert jo.format_fix_csv(res.passes[ite].rows) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    assert jo.format_tally(res.passes[0]) == helpers.expected_file(case, "0qValCalcHelper.csv")
    assert jo.format_tally(res.passes[-1]) == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])


def _check_against_expected(case, res):
    for ite in range(case["passes"]):
        assert jo.format_fix_csv(res.passes[ite].rows) == helpers.expected_file(case, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(res.seqs) == helpers.expected_file(case, "_iter%d_asm.fa.fixed.fa.tmp" % (case["passes"] - 1))
    assert jo.format_tally(res.passes[0]) == helpers.expected_file(case, "0qValCalcHelper.csv")
    assert jo.format_tally(res.passes[-1]) == helpers.expected_file(case, "%dqValCalcHelper.csv" % case["passes"])


def test_rare_branches_are_pinned_by_reference_runs(tmp_path):
    """tests/golden/branches/*: outputs of the unmodified src/jasper.py on fuzz cases picked for the branches the big goldens
    miss -- fixdiploid's "s" side, a single inserted base (grow_single), substitution and deletion both failing (unfixed_k), the
    rolling threshold.  The restatement reproduces every file, and together the cases take EVERY branch at least five times."""
    total = {}
    for name in helpers.BRANCH_GOLDENS:
        case = helpers.load_golden(name, str(tmp_path))
        odb = helpers.OracleDB(case["db"])
        oracle = jo.PolishOracle(odb.count, case["k"], case["threshold"])
        res = oracle.polish(jo.read_fasta(case["fasta"]), case["passes"])
        _check_against_expected(case, res)
        assert oracle.branch_report() == case["branches"]
        for b, n in oracle.branch_report().items():
            total[b] = total.get(b, 0) + n
    assert sorted(total) == sorted(helpers.ALL_BRANCHES)
    assert min(total.values()) >= 5, total


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_lookup_oracle_matches_binding_vectors(golden_cases, name):
    """C(s) and the canonical k-mer text for edge-case strings recorded from the real binding."""
    case = golden_cases[name]
    odb = helpers.OracleDB(case["db"])
    lib = odb.lib
    for s, canon, count in case["lookups"]:
        assert odb.count(s) == count, s
        b = s.encode()
        z = b.find(b"\0")
        m, c = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
        lib.jfo_mer_from_chars(b, len(b), odb.k, m)
        lib.jfo_canonical(m, odb.k, c)
        out = C.create_string_buffer(odb.k + 1)
        lib.jfo_mer_to_string(c, odb.k, out)
        assert out.value.decode() == canon


def test_mer_dna_known_answers():
    """JF!unit_tests/test_mer_dna.cc semantics at k=5 (verified on the reference)."""
    lib = helpers.oracle_lib()

    def mer(s, k=5, canon=False):
        b = s.encode()
        m, c = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
        lib.jfo_mer_from_chars(b, len(b), k, m)
        if canon:
            lib.jfo_canonical(m, k, c)
            m = c
        out = C.create_string_buffer(k + 1)
        lib.jfo_mer_to_string(m, k, out)
        return out.value.decode()

    assert mer("") == "AAAAA"
    assert mer("ACG") == "ACGAA"
    assert mer("ACNGT") == "ACAAA"
    assert mer("ACGTTGG") == "ACGTT"
    assert mer("ACGTTGG", canon=True) == "AACGT"
    assert mer("TTTTT", canon=True) == "AAAAA"
    assert mer("acgtn") == "ACGTA"


def test_oracle_file_order_is_pos_then_key(golden_cases):
    """Records are strictly increasing in (M*key & mask, key) (JF!include/jellyfish/mer_heap.hpp:26-30)."""
    case = golden_cases["k37_mix"]
    odb = helpers.OracleDB(case["db"])
    lo, hi, _ = odb.records()
    key = (C.c_uint64 * 2)()
    pos = np.empty(len(lo), dtype=np.uint64)
    for i in range(len(lo)):
        key[0], key[1] = int(lo[i]), int(hi[i])
        pos[i] = odb.lib.jfo_hash_pos(odb.h, key)
    order = np.lexsort((lo, hi, pos))
    assert np.array_equal(order, np.arange(len(lo)))


def test_every_record_is_found(golden_cases):
    """JF!swig/python/test_mer_file.py test_query: qf[mer] == count for every record."""
    case = golden_cases["k25_mix"]
    odb = helpers.OracleDB(case["db"])
    lo, hi, cnt = odb.records()
    key = (C.c_uint64 * 2)()
    for i in range(0, len(lo), 7):
        key[0], key[1] = int(lo[i]), int(hi[i])
        assert odb.lib.jfo_lookup_key(odb.h, key) == cnt[i]


def test_synthetic_jf_writer_roundtrip(tmp_path):
    """A .jf written by jasper_b200.synth is read back record for record by the oracle and, when the
    reference build is present, by the reference's own QueryMerFile / `jellyfish dump`."""
    rng = np.random.default_rng(5)
    g = synth.random_genome(3000, rng)
    for k in (25, 37, 48):
        lo, hi = synth.canonical_keys(g, k)
        lo, hi, cnt = synth.unique_keys(lo, hi)
        cnt = cnt + rng.integers(0, 70000, len(cnt))
        path = str(tmp_path / ("w%d.jf" % k))
        synth.write_jf(path, lo, hi, cnt, k, seed=k)
        odb = helpers.OracleDB(path)
        assert odb.k == k and odb.n_records == len(lo)
        s = synth.codes_to_bytes(g).decode()
        want = {}
        for a, b, c in zip(lo, hi, cnt):
            want[(int(a), int(b))] = int(c)
        for i in range(0, len(s) - k, 13):
            m, c = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
            w = s[i:i + k].encode()
            odb.lib.jfo_mer_from_chars(w, k, k, m)
            odb.lib.jfo_canonical(m, k, c)
            assert odb.count(s[i:i + k]) == want[(c[0], c[1])] & 0xFFFFFFFF
        jf = helpers.reference_binding()
        if jf is not None and k <= 37:
            qf = jf.QueryMerFile(path)
            for i in range(0, len(s) - k, 29):
                assert qf[jf.MerDNA(s[i:i + k]).get_canonical()] == odb.count(s[i:i + k])


def test_oracle_matches_reference_binding_on_random_strings(golden_cases):
    jf = helpers.reference_binding()
    if jf is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    case = golden_cases["k31_dirty"]
    odb = helpers.OracleDB(case["db"])
    qf = jf.QueryMerFile(case["db"])
    rng = np.random.default_rng(9)
    seq = "".join(s for _, s in jo.read_fasta(case["fasta"]))
    alphabet = "ACGTacgtNnRY-"
    for _ in range(3000):
        if rng.random() < 0.5:
            at = int(rng.integers(0, len(seq) - 80)); s = seq[at:at + int(rng.integers(0, 70))]
        else:
            s = "".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), int(rng.integers(0, 60))))
        assert qf[jf.MerDNA(s).get_canonical()] == odb.count(s), s


def test_threshold_restatement():
    """synth.solid_threshold_from_histogram restates src/jellyfish.py (rows of the golden histograms)."""
    rows = [(1, 100), (2, 50), (3, 20), (4, 10), (5, 12), (6, 30)]
    assert synth.solid_threshold_from_histogram(rows) == 2
    assert synth.solid_threshold_from_histogram([(1, 100), (2, 50), (3, 60)]) is None       # threshold 1 < 2
    assert synth.solid_threshold_from_histogram([(1, 100), (2, 50), (3, 40)]) is None       # no local minimum
    assert synth.solid_threshold_from_histogram([(1, 10), (2, 10), (3, 10), (4, 10), (5, 11)]) == 2   # ties continue


def test_qv_bc_emulation():
    """5021/199962 at k=25 prints 29.92812; 1200/3.1e9 at k=37 prints 79.80504."""
    for fn in (jo.qv_like_bc, jo.qv_exact_math):
        assert fn(5021, 199962, 25) == "29.92812"
        assert fn(1200, 3100000000, 37) == "79.80504"
        assert fn(0, 1000, 25) == "Inf"


def test_pairwise_shim_basic():
    from Bio import pairwise2
    a = pairwise2.align.globalms("ACGTTGCA", "ACGTAGCA", 0, -1, -1, -1)[0]
    assert (a[0], a[1]) == ("ACGTTGCA", "ACGTAGCA")
    a = pairwise2.align.globalms("ACGTGCA", "ACGTTGCA", 0, -1, -1, -1)[0]
    assert a[0].replace("-", "") == "ACGTGCA" and a[1] == "ACGTTGCA" and a[0].count("-") == 1
    assert pairwise2.align.globalms("", "ACG", 0, -1, -1, -1) == []
    assert pairwise2.first_global_alignment("AAAC", "AAC")[0][:2] == pairwise2.align.globalms("AAAC", "AAC", 0, -1, -1, -1)[0][:2]

"""jasper_b200/pipeline.py against the UNMODIFIED src/jasper.sh (tests/golden/*/expected_sh, made by
tests/golden/make_golden_sh.py): threshold, histogram file, polished FASTA and the sorted, awk-mangled fixes.csv.
On CPUs the oracle stands in for the GPU tables (this file is test infrastructure); the GPU test runs the product."""
import json
import os

import pytest

import helpers
import jasper_oracle as jo
from jasper_b200 import capi, pipeline

SH_CASES = ["k31_dirty", "k21_even"]


class _OracleTable(object):
    def __init__(self, db_path, histo_rows):
        self.odb = helpers.OracleDB(db_path)
        self.rows = histo_rows

    def histogram(self, high):
        return self.rows

    def solid_threshold(self, rows):
        return capi.solid_threshold(rows)          # host arithmetic of the library (src/jellyfish.py), no GPU involved

    def polish(self, seqs, k, thre, passes):
        res, _ = helpers.oracle_polish(self.odb, seqs, k, thre, passes)

        class R(object):
            pass
        r = R()
        r.status = 0
        r.tallies = [(p.wrong, p.total) for p in res.passes]
        r.rows = [[list(x) for x in p.rows] for p in res.passes]
        r.seqs = res.seqs
        return r

    def close(self):
        pass


def _records(path):
    out, name = {}, None
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].rstrip("\n")
            assert name not in out
            out[name] = ""
        else:
            out[name] += line.rstrip("\n")
    return out


def _check_outputs(case, work, k):
    exp = os.path.join(case["dir"], "expected_sh")
    assert open(os.path.join(work, "threshold.txt")).read() == open(os.path.join(exp, "threshold.txt")).read()
    assert open(os.path.join(work, "jfhisto%d.csv" % k)).read() == open(os.path.join(exp, "jfhisto%d.csv" % k)).read()
    assert open(os.path.join(work, "asm.fa.fixes.csv"), "rb").read() == open(os.path.join(exp, "asm.fa.fixes.csv"), "rb").read()
    # the reference prints contigs in perl's hash order (random per run): compare the records, not the order
    assert _records(os.path.join(work, "asm.fa.polished.fasta")) == _records(os.path.join(exp, "asm.fa.polished.fasta"))


@pytest.mark.parametrize("name", SH_CASES)
def test_pipeline_reproduces_jasper_sh_with_the_oracle(golden_cases, name, tmp_path):
    case = golden_cases[name]
    meta = json.load(open(os.path.join(case["dir"], "expected_sh", "meta.json")))
    histo = [tuple(int(x) for x in l.split()) for l in open(os.path.join(case["dir"], "histo.csv"))]
    out = pipeline.run(case["fasta"], case["db"], k=meta["k"], passes=meta["passes"], threads=meta["threads"], workdir=str(tmp_path),
                       open_tables=lambda path, devices: [_OracleTable(path, histo)])
    assert out["batch_size"] == meta["batch_size"]
    assert out["threshold"] == case["threshold"]
    _check_outputs(case, str(tmp_path), meta["k"])


def test_fix_row_merge_is_the_awk_sort_chain():
    """Header mangling, ':' splitting (only the first two ':' fields survive), numeric keys, CRLF kept in the last field."""
    rows0 = [["c:0", 10, "A", "sT"], ["c:0", 9, "-", "iG"], ["b:100", 5, "AA", "d-"], ["odd:name:0", 3, "C", "sG"]]
    rows1 = [["c:0", 10, "C", "sA"]]
    got = pipeline.merge_fix_rows([rows0, rows1]).decode()
    assert got.split("\n") == [
        "Contig:Base_coord Original Mutation\r ",       # the header goes through the same awk chain
        "b:100 5 AA d-\r",
        "c:0 9 - iG\r",
        "c:0 10 A sT\r",                                # equal keys: the whole line decides
        "c:0 10 C sA\r",
        "odd:name   ",                                  # awk -F ':' keeps two fields: a name with ':' loses the rest of its row
        "",
    ]


def test_q_value_matches_the_oracle_restatement_of_bc():
    """The Q text of src/jasper.sh:239-257 goes through GNU bc; bc is absent here, so both sides restate its published rules
    (POSIX scale rules, the series of libmath.b) -- the product on integers (jasper_b200/bcmath.py), the oracle on `decimal`
    numbers (oracle/jasper_oracle.py).  Same text on the tallies of the bench workloads, the survey's two constants and 300
    seeded random tallies; the exact-math emulation of round 1 may differ in the last digit only."""
    import random
    from jasper_b200 import bcmath
    cases = [(2270316, 63999135, 37), (1627, 63999312, 37), (0, 1000, 25), (3284, 29980, 37), (5021, 199962, 25), (1200, 3100000000, 37),
             (110259410, 3099993862, 37), (49274, 3099993674, 37)]
    rng = random.Random(20)
    for _ in range(300):
        total = rng.randint(10 ** 3, 4 * 10 ** 9)
        cases.append((int(total * 10 ** rng.uniform(-8, -0.5)), total, rng.choice([21, 25, 31, 37, 48])))
    for wrong, total, k in cases:
        got = pipeline.q_value(wrong, total, k)
        assert got == jo.qv_like_bc(wrong, total, k), (wrong, total, k)
        exact = jo.qv_exact_math(wrong, total, k)
        if got != exact:
            assert got != "Inf" and exact != "Inf" and abs(float(got) - float(exact)) < 2.5e-5, (wrong, total, k, got, exact)
    assert pipeline.q_value(5021, 199962, 25) == "29.92812" and pipeline.q_value(1200, 3100000000, 37) == "79.80504"
    # what `bc -l` prints at its default scale of 20 (well-known outputs of the real program)
    B = bcmath.Bc
    assert bcmath.e(B(1), 20).text() == "2.71828182845904523536"
    assert bcmath.e(B(2), 20).text() == "7.38905609893065022723"
    assert bcmath.l(B(2), 20).text() == ".69314718055994530941"
    assert bcmath.l(B(10), 20).text() == "2.30258509299404568401"
    assert bcmath.sqrt(B(2), 20).text() == "1.41421356237309504880"
    assert bcmath.div(B(1), B(3), 20).text() == ".33333333333333333333" and bcmath.div(B(-1), B(3), 5).text() == "-.33333"
    assert bcmath.mul(B.parse("1.5"), B.parse("2.25"), 0).text() == "3.37"        # scale = min(1 + 2, max(0, 1, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SH_CASES)
def test_pipeline_reproduces_jasper_sh_on_the_gpu(golden_cases, name, tmp_path):
    case = golden_cases[name]
    meta = json.load(open(os.path.join(case["dir"], "expected_sh", "meta.json")))
    out = pipeline.run(case["fasta"], case["db"], k=meta["k"], passes=meta["passes"], threads=meta["threads"], workdir=str(tmp_path))
    assert out["batch_size"] == meta["batch_size"] and out["threshold"] == case["threshold"]
    _check_outputs(case, str(tmp_path), meta["k"])


def test_pipeline_shards_pieces_over_several_tables(golden_cases, tmp_path):
    """Two table stand-ins = two GPUs: the pieces are sharded by length, polished on two threads, and the outputs are the
    same bytes as with one."""
    case = golden_cases["k31_dirty"]
    meta = json.load(open(os.path.join(case["dir"], "expected_sh", "meta.json")))
    histo = [tuple(int(x) for x in l.split()) for l in open(os.path.join(case["dir"], "histo.csv"))]
    tables = []

    def two(path, devices):
        tables.extend([_OracleTable(path, histo), _OracleTable(path, histo)])
        calls = []
        for t in tables:
            orig = t.polish
            t.polish = (lambda o: (lambda seqs, k, thre, passes: (calls.append(len(seqs)), o(seqs, k, thre, passes))[1]))(orig)
        two.calls = calls
        return tables
    pipeline.run(case["fasta"], case["db"], k=meta["k"], passes=meta["passes"], threads=meta["threads"], workdir=str(tmp_path), open_tables=two)
    assert len(two.calls) == 2 and all(n > 0 for n in two.calls)          # both tables got work
    _check_outputs(case, str(tmp_path), meta["k"])


def test_pipeline_restart_markers_and_capacity_groups(golden_cases, tmp_path):
    """src/jasper.sh's restart protocol: a finished stage (jasper.<stage>.success) is skipped on the next run, a stage that
    runs invalidates the ones behind it; and a table that reports a small capacity gets its shard in several calls, each
    leaving the files of one reference worker -- the joined outputs do not change."""
    case = golden_cases["k31_dirty"]
    meta = json.load(open(os.path.join(case["dir"], "expected_sh", "meta.json")))
    histo = [tuple(int(x) for x in l.split()) for l in open(os.path.join(case["dir"], "histo.csv"))]
    work = str(tmp_path)
    calls = []

    def tables(path, devices):
        t = _OracleTable(path, histo)
        orig = t.polish
        t.polish = lambda seqs, k, thre, passes: (calls.append(sum(len(s) for _, s in seqs)), orig(seqs, k, thre, passes))[1]
        t.capacity = lambda: 9000
        return [t]

    kw = dict(k=meta["k"], passes=meta["passes"], threads=meta["threads"], workdir=work)
    out = pipeline.run(case["fasta"], case["db"], open_tables=tables, debug=True, **kw)
    assert len(calls) >= 3 and max(calls) <= 9000 + meta["batch_size"]            # several calls, none far above the capacity
    _check_outputs(case, work, meta["k"])
    for m in ("split", "histo", "correct", "join"):
        assert os.path.exists(os.path.join(work, "jasper.%s.success" % m))
    kept = sorted(f for f in os.listdir(work) if f.endswith(".fix.csv"))
    assert len(kept) == len(calls) * meta["passes"]                                # -d keeps the per-batch fix.csv files
    assert not [f for f in os.listdir(work) if f.endswith(".fixed.fa") or f.endswith("qValCalcHelper.csv")]

    # everything done: a rerun touches neither the table nor the outputs
    def no_tables(path, devices):
        raise AssertionError("a finished run must not open the database again")
    before = {f: os.path.getmtime(os.path.join(work, f)) for f in ("asm.fa.polished.fasta", "asm.fa.fixes.csv")}
    pipeline.run(case["fasta"], case["db"], open_tables=no_tables, debug=True, **kw)
    assert before == {f: os.path.getmtime(os.path.join(work, f)) for f in before}
    # the join stage alone (its inputs: the per-batch fix.csv kept by -d; the polished pieces are gone, like in the reference,
    # so only the fixes are rebuilt here) -- and a lost polish marker re-polishes and re-joins
    os.remove(os.path.join(work, "jasper.correct.success"))
    os.remove(os.path.join(work, "asm.fa.polished.fasta"))
    n0 = len(calls)
    out2 = pipeline.run(case["fasta"], case["db"], open_tables=tables, **kw)
    assert len(calls) > n0 and out2["tallies"] == out["tallies"]
    _check_outputs(case, work, meta["k"])
    assert not [f for f in os.listdir(work) if f.endswith(".fix.csv")]             # without -d the per-batch files are removed


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` on a CPU-only box: the reference arm's JSON line (the unmodified src/jasper.py under xargs -P when
    oracle/_ref is built, else the oracle port), with the keys the contract asks for and nothing else on stdout."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--genome-mb", "0.4", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "assembly bp polished/sec" and d["unit"] == "bp polished/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]

"""BASELINE configs[0]: the 4.6 Mb E. coli-size case of SURVEY.md Appendix I, database built by the reference's own `jellyfish count`
(oracle/_ref/bin/jellyfish), polished through jasper_b200/pipeline.py and compared with what the UNMODIFIED src/jasper.sh produced
on the same inputs in the build container (tests/golden/config1/expected.json, written by tests/golden/make_config1.py): the md5
of the polished records and of fixes.csv, 4,557 fix rows, threshold 2, batch size 517,504."""
import hashlib
import importlib.util
import json
import os

import pytest

import helpers
from jasper_b200 import pipeline

GOLD = os.path.join(helpers.GOLDEN_DIR, "config1", "expected.json")
JELLYFISH = os.path.join(helpers.ROOT, "oracle", "_ref", "bin", "jellyfish")
pytestmark = pytest.mark.skipif(not os.path.exists(JELLYFISH), reason="needs the reference's jellyfish binary (oracle/_ref, built by `make -C oracle ref`)")


def _maker():
    spec = importlib.util.spec_from_file_location("make_config1", os.path.join(helpers.GOLDEN_DIR, "make_config1.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def config1(tmp_path_factory):
    mk = _maker()
    d = str(tmp_path_factory.mktemp("config1"))
    reads, asm, asm_len, n_err = mk.inputs(d)
    want = json.load(open(GOLD))
    assert hashlib.md5(open(reads, "rb").read()).hexdigest() == want["reads_md5"]
    assert hashlib.md5(open(asm, "rb").read()).hexdigest() == want["asm_md5"] and asm_len == want["assembly_bp"] and n_err == want["error_sites"]
    db = os.path.join(d, "mer_counts25.jf")
    mk.count_db(JELLYFISH, reads, db)                     # the reference binary: building the database is out of scope
    os.remove(reads)
    return {"mk": mk, "dir": d, "asm": asm, "db": db, "want": want}


def _check(c, work, out):
    mk, want = c["mk"], c["want"]
    assert out["batch_size"] == want["batch_size"] and str(out["threshold"]) == want["threshold"]
    md5_fa, bp = mk.fasta_digest(os.path.join(work, "asm.fa.polished.fasta"))
    fixes = open(os.path.join(work, "asm.fa.fixes.csv"), "rb").read()
    n_rows, kinds = mk.fixes_census(os.path.join(work, "asm.fa.fixes.csv"))
    assert (n_rows, kinds) == (want["fix_rows"], want["fix_kinds"]) and n_rows == 4557
    assert bp == want["polished_bp"] and md5_fa == want["polished_fasta_md5_sorted_records"]
    assert hashlib.md5(fixes).hexdigest() == want["fixes_csv_md5"] and len(fixes) == want["fixes_csv_bytes"]


@pytest.mark.gpu
def test_config1_through_the_gpu_pipeline_equals_jasper_sh(config1, tmp_path):
    work = str(tmp_path)
    out = pipeline.run(config1["asm"], config1["db"], k=25, passes=2, threads=8, workdir=work)
    _check(config1, work, out)
    assert out["pieces"] == 9


def test_config1_restatement_equals_jasper_sh(config1, tmp_path):
    """The same comparison with the CPU restatement standing in for the GPU tables (pins oracle + pipeline at this scale)."""
    from test_pipeline import _OracleTable
    from jasper_b200 import capi
    odb = helpers.OracleDB(config1["db"])
    lo, hi, cnt = odb.records()
    import numpy as np
    c, n = np.unique(np.minimum(cnt & 0xFFFFFFFF, 10001), return_counts=True)
    histo = [(int(a), int(b)) for a, b in zip(c, n)]
    work = str(tmp_path)
    out = pipeline.run(config1["asm"], config1["db"], k=25, passes=2, threads=8, workdir=work, open_tables=lambda p, d: [_OracleTable(p, histo)])
    _check(config1, work, out)

"""The restatement (oracle/jasper_oracle.py) against the UNMODIFIED /root/reference/src/jasper.py on the fuzz inputs the CUDA path is
compared with everywhere else.  Needs the reference checkout and its compiled Jellyfish binding (oracle/_ref), i.e. it runs in
the build container and is skipped on the GPU box (nothing under -m gpu reads /root/reference).  This is what pins the oracle on
inputs the committed goldens do not hold: every fuzz seed used by tests/test_host.py and tests/test_gpu.py goes through the real
thing here, files compared byte for byte."""
import glob
import os
import subprocess
import sys

import pytest

import helpers
import jasper_oracle as jo

REF_JASPER = "/root/reference/src/jasper.py"
HAVE_REF = os.path.exists(REF_JASPER) and bool(glob.glob(os.path.join(helpers.ROOT, "oracle", "_ref", "_dna_jellyfish*.so")))
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference and oracle/_ref (build container only)")


def _run_reference(tmp, db, recs, k, thr, passes):
    with open(os.path.join(tmp, "asm.fa"), "w") as f:
        for name, s in recs:
            f.write(">%s\n%s\n" % (name, s))
    env = dict(os.environ, PYTHONPATH=os.path.join(helpers.ROOT, "oracle"))
    cmd = [sys.executable, REF_JASPER, "--db", db, "--query", "asm.fa", "--ksize", str(k), "-p", str(passes), "--fix", "--fout", "asm.fa.fix.csv",
           "-ff", "asm.fa.fixed.fa.tmp", "--test", "-thre", str(thr)]
    return subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)


def _read(tmp, fn):
    with open(os.path.join(tmp, fn), "rb") as f:
        return f.read()


@pytest.mark.parametrize("seed", list(range(40)) + [100, 101, 102, 103])
def test_restatement_equals_the_unmodified_reference_on_fuzz_inputs(tmp_path, seed):
    big = seed >= 100
    k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
    recs, db, thr = helpers.write_fuzz_db(tmp_path, seed, k, big=big)
    # the reference keys its sequences by name in a dict (src/jasper.py:615-631): feed both sides what parse_fasta makes of the file
    passes = 2
    tmp = str(tmp_path)
    r = _run_reference(tmp, db, recs, k, thr, passes)
    odb = helpers.OracleDB(db)
    want, oracle = helpers.oracle_polish(odb, jo.read_fasta(os.path.join(tmp, "asm.fa")), k, thr, passes)
    if want is None:                                   # the restatement says the reference raises here: it must exit 1
        assert r.returncode == 1
        return
    assert r.returncode == 0, r.stdout + r.stderr
    for ite in range(passes):
        assert jo.format_fix_csv(want.passes[ite].rows) == _read(tmp, "_iter%d_asm.fa.fix.csv" % ite)
    assert jo.format_fasta(want.seqs) == _read(tmp, "_iter%d_asm.fa.fixed.fa.tmp" % (passes - 1))
    assert jo.format_tally(want.passes[0]) == _read(tmp, "0qValCalcHelper.csv")
    assert jo.format_tally(want.passes[-1]) == _read(tmp, "%dqValCalcHelper.csv" % passes)

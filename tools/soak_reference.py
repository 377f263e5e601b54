#!/usr/bin/env python
"""Soak (build container only): the restatement oracle/jasper_oracle.py against the UNMODIFIED /root/reference/src/jasper.py on fuzz
seeds beyond those tests/test_reference_container.py runs.  Files compared byte for byte, as in that test.

    python tools/soak_reference.py FIRST_SEED N_SEEDS
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
import jasper_oracle as jo  # noqa: E402
import test_reference_container as trc  # noqa: E402


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    if not trc.HAVE_REF:
        print("no /root/reference or oracle/_ref here")
        return 2
    bad, crashes, rows = [], 0, 0
    for seed in range(first, first + n):
        with tempfile.TemporaryDirectory() as tmp:
            k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
            recs, db, thr = helpers.write_fuzz_db(tmp, seed, k)
            r = trc._run_reference(tmp, db, recs, k, thr, 2)
            odb = helpers.OracleDB(db)
            want, _ = helpers.oracle_polish(odb, jo.read_fasta(os.path.join(tmp, "asm.fa")), k, thr, 2)
            if want is None:
                crashes += 1
                if r.returncode != 1:
                    bad.append((seed, "the restatement raises, the reference exits %d" % r.returncode))
                continue
            ok = r.returncode == 0
            for ite in range(2):
                ok = ok and jo.format_fix_csv(want.passes[ite].rows) == trc._read(tmp, "_iter%d_asm.fa.fix.csv" % ite)
                rows += len(want.passes[ite].rows)
            ok = ok and jo.format_fasta(want.seqs) == trc._read(tmp, "_iter1_asm.fa.fixed.fa.tmp")
            ok = ok and jo.format_tally(want.passes[0]) == trc._read(tmp, "0qValCalcHelper.csv")
            ok = ok and jo.format_tally(want.passes[-1]) == trc._read(tmp, "2qValCalcHelper.csv")
            if not ok:
                bad.append((seed, "files differ (exit %d)" % r.returncode))
    print("seeds %d..%d through the unmodified src/jasper.py: %d divergences %s; %d reference crashes, %d fix rows compared"
          % (first, first + n - 1, len(bad), bad[:5], crashes, rows))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

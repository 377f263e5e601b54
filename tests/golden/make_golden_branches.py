#!/usr/bin/env python
"""Golden fixtures for the RARE branches of src/jasper.py (tests/golden/branches/s<seed>/), made by running the UNMODIFIED
reference on fuzz cases chosen because the restatement's branch_report() shows them hitting branches the four big goldens
miss or barely touch: fixdiploid's "s" side (src/jasper.py:374-375), a single inserted base in front of the good k-mer
(:471-476), substitution and deletion both failing (:232-245, "unfixed_k"), the rolling-threshold branch (:86-97).

Runs only in the build container (needs /root/reference and `make -C oracle ref`).  The inputs are helpers.fuzz_case(seed)
(a synthetic .jf written by jasper_b200.synth, readable by the reference's own binding); the expected files are whatever
/root/reference/src/jasper.py wrote.  case.json records the branch counts so that tests can assert the coverage.

    python tests/golden/make_golden_branches.py
"""
import gzip, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
import jasper_oracle as jo  # noqa: E402

REF_SRC = "/root/reference/src"
SEEDS = [32, 99, 278, 377, 5, 9, 39, 66]
PASSES = 2


def run_reference(tmp, db, k, thr, passes):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle"))
    cmd = [sys.executable, os.path.join(REF_SRC, "jasper.py"), "--db", db, "--query", "asm.fa", "--ksize", str(k), "-p", str(passes),
           "--fix", "--fout", "asm.fa.fix.csv", "-ff", "asm.fa.fixed.fa.tmp", "--test", "-thre", str(thr)]
    return subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)


def main():
    total = {}
    for seed in SEEDS:
        k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
        tmp = tempfile.mkdtemp(prefix="golden_branch_%d" % seed)
        recs, db, thr = helpers.write_fuzz_db(tmp, seed, k)
        with open(os.path.join(tmp, "asm.fa"), "w") as f:
            for name, s in recs:
                f.write(">%s\n" % name)
                for i in range(0, len(s), 70):
                    f.write(s[i:i + 70] + "\n")
        r = run_reference(tmp, db, k, thr, PASSES)
        if r.returncode != 0:
            raise SystemExit("reference jasper.py failed on seed %d:\n%s\n%s" % (seed, r.stdout, r.stderr))
        odb = helpers.OracleDB(db)
        res, o = helpers.oracle_polish(odb, jo.read_fasta(os.path.join(tmp, "asm.fa")), k, thr, PASSES)
        out = os.path.join(HERE, "branches", "s%d" % seed)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(os.path.join(out, "expected"))
        shutil.copy(os.path.join(tmp, "asm.fa"), out)
        with open(db, "rb") as f, gzip.GzipFile(os.path.join(out, "db.jf.gz"), "wb", mtime=0) as gz:
            shutil.copyfileobj(f, gz)
        for fn in ["_iter%d_asm.fa.fix.csv" % i for i in range(PASSES)] + ["_iter%d_asm.fa.fixed.fa.tmp" % (PASSES - 1), "0qValCalcHelper.csv",
                                                                           "%dqValCalcHelper.csv" % PASSES]:
            shutil.copy(os.path.join(tmp, fn), os.path.join(out, "expected", fn))
        br = o.branch_report()
        json.dump(dict(name="branches/s%d" % seed, k=k, passes=PASSES, threshold=thr, seed=seed, branches=br,
                       source="helpers.fuzz_case(%d, %d) through the unmodified /root/reference/src/jasper.py" % (seed, k)),
                  open(os.path.join(out, "case.json"), "w"), indent=1)
        for b, n in br.items():
            total[b] = total.get(b, 0) + n
        print("s%d k=%d thr=%d" % (seed, k, thr), br)
        shutil.rmtree(tmp)
    print("all cases:", dict(sorted(total.items())))


if __name__ == "__main__":
    main()

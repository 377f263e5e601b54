#!/usr/bin/env python
"""Golden fixture tests/golden/branches/deep_bridge/: a multi-error stretch whose de-Bruijn search (src/jasper.py:527-583,
base_extension) keeps thousands of paths alive for hundreds of rounds -- more than a million path extensions, i.e. more trie
nodes than one standard bridging slot of the CUDA walker holds (kTrieCap, jasper_b200/csrc/jb_polish.cu), so the library has
to re-walk the contig with the grown pool.  Expected files = what the UNMODIFIED /root/reference/src/jasper.py wrote.

The k-mer graph: two haplotypes that differ by a SNP every `gap` bases over the first part of the stretch (every SNP doubles
the live paths: the k-mers of both alleles are solid and, gap > k, never see two SNPs at once), then a long shared tail the
paths walk in lockstep.  The assembly carries an error every `k - 5` bases over the whole stretch, so every window there is
weak and the stretch is one multi-error run.

Runs only in the build container (needs /root/reference and `make -C oracle ref`).

    python tests/golden/make_golden_deep_bridge.py
"""
import gzip, json, os, shutil, subprocess, sys, tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
import jasper_oracle as jo  # noqa: E402
from jasper_b200 import synth  # noqa: E402
from make_golden_branches import run_reference  # noqa: E402

K, THR, PASSES = 25, 3, 2
N_SNP, GAP, TAIL, FLANK = 11, 30, 640, 1500


def build():
    rng = np.random.default_rng(4242)
    stretch = N_SNP * GAP + TAIL
    g = synth.random_genome(2 * FLANK + stretch, rng)
    h2 = g.copy()
    for s in range(N_SNP):
        at = FLANK + 10 + s * GAP
        h2[at] = (h2[at] + 1 + int(rng.integers(0, 3))) % 4
    lo, hi = synth.canonical_keys(np.concatenate([g, h2]), K)      # (the k-1 windows across the joint are harmless extras)
    lo, hi, cnt = synth.unique_keys(lo, hi, np.full(len(lo), 12, dtype=np.int64))
    a = g.copy()
    for at in range(FLANK + 3, FLANK + stretch - 3, K - 5):
        a[at] = (a[at] + 1 + int(rng.integers(0, 3))) % 4
    return [("bubbly", synth.codes_to_bytes(a).decode()), ("plain", synth.codes_to_bytes(g[:900]).decode())], (lo, hi, cnt)


def main():
    tmp = tempfile.mkdtemp(prefix="golden_deep_bridge")
    recs, (lo, hi, cnt) = build()
    db = os.path.join(tmp, "db.jf")
    synth.write_jf(db, lo, hi, cnt, K, seed=11)
    with open(os.path.join(tmp, "asm.fa"), "w") as f:
        for name, s in recs:
            f.write(">%s\n" % name)
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    r = run_reference(tmp, db, K, THR, PASSES)
    if r.returncode != 0:
        raise SystemExit("reference jasper.py failed:\n%s\n%s" % (r.stdout, r.stderr))
    odb = helpers.OracleDB(db)
    res, o = helpers.oracle_polish(odb, jo.read_fasta(os.path.join(tmp, "asm.fa")), K, THR, PASSES)
    out = os.path.join(HERE, "branches", "deep_bridge")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(os.path.join(out, "expected"))
    shutil.copy(os.path.join(tmp, "asm.fa"), out)
    with open(db, "rb") as f, gzip.GzipFile(os.path.join(out, "db.jf.gz"), "wb", mtime=0) as gz:
        shutil.copyfileobj(f, gz)
    for fn in ["_iter%d_asm.fa.fix.csv" % i for i in range(PASSES)] + ["_iter%d_asm.fa.fixed.fa.tmp" % (PASSES - 1), "0qValCalcHelper.csv",
                                                                       "%dqValCalcHelper.csv" % PASSES]:
        shutil.copy(os.path.join(tmp, fn), os.path.join(out, "expected", fn))
    br = o.branch_report()
    json.dump(dict(name="branches/deep_bridge", k=K, passes=PASSES, threshold=THR, branches=br, oracle_lookups=res.lookups if res else None,
                   source="make_golden_deep_bridge.py (%d SNP bubbles %d apart + %d shared bases, an error every %d bases) through the unmodified "
                          "/root/reference/src/jasper.py" % (N_SNP, GAP, TAIL, K - 5)),
              open(os.path.join(out, "case.json"), "w"), indent=1)
    print("deep_bridge", br, "oracle lookups", res.lookups if res else None)
    print(open(os.path.join(out, "expected", "_iter0_asm.fa.fix.csv")).read()[:600])
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and `make -C oracle ref`):
  * `oracle/_ref/bin/jellyfish count|histo` (reference binary) builds each case's .jf
  * /root/reference/src/jellyfish.py picks the threshold (falls back to a stated value)
  * /root/reference/src/jasper.py polishes, importing the reference's own Jellyfish binding
    through oracle/dna_jellyfish.py and the Bio.pairwise2 stand-in oracle/Bio/pairwise2.py
Outputs per case: case.json, asm.fa, db.jf.gz, and expected/ (byte-exact reference outputs).
Also writes lookups.json: C(s) for edge-case strings straight from the binding.

    python tests/golden/make_golden.py [case ...]
"""
import gzip, json, os, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from jasper_b200 import synth  # noqa: E402

REF_SRC = "/root/reference/src"
JELLYFISH = os.path.join(ROOT, "oracle", "_ref", "bin", "jellyfish")


def decorate(seq: bytearray, rng, n_runs=0, lower=0, iupac=0):
    """Sprinkle N runs, lowercase stretches and IUPAC codes."""
    L = len(seq)
    for _ in range(n_runs):
        ln = int(rng.integers(1, 60)); at = int(rng.integers(0, L - ln))
        seq[at:at + ln] = (b"N" if rng.random() < 0.7 else b"n") * ln
    for _ in range(lower):
        ln = int(rng.integers(5, 300)); at = int(rng.integers(0, L - ln))
        seq[at:at + ln] = bytes(seq[at:at + ln]).lower()
    for _ in range(iupac):
        at = int(rng.integers(0, L)); seq[at] = rng.choice(list(b"RYKMSWryBDHV-*"))
    return seq


def clustered_errors(a: np.ndarray, rng, n_sites, span):
    """Pairs/triples of substitutions within `span` bases (exercises the nbad > k branch)."""
    a = a.copy()
    for _ in range(n_sites):
        at = int(rng.integers(100, len(a) - 100))
        for off in sorted(set(int(x) for x in rng.integers(0, span, int(rng.integers(2, 4))))):
            a[at + off] = (a[at + off] + rng.integers(1, 4)) % 4
    return a


def homopolymer_errors(a: np.ndarray, rng, n_sites):
    """Lengthen/shorten existing homopolymer runs by 1-2 bases."""
    a = list(a)
    for _ in range(n_sites):
        at = int(rng.integers(100, len(a) - 100))
        j = at
        while j + 1 < len(a) - 100 and not (a[j] == a[j + 1] == a[j + 2]):
            j += 1
        if j + 1 >= len(a) - 100:
            continue
        n = int(rng.integers(1, 3))
        if rng.random() < 0.5:
            for _ in range(n):
                a.insert(j, a[j])
        else:
            del a[j:j + n]
    return np.array(a, dtype=np.uint8)


CASES = {
    "k25_mix": dict(k=25, passes=2, seed=101, contigs=[30000], err=0.004, cov=26, rerr=0.0015, hp=40, hp_err=12, cluster=6),
    "k37_mix": dict(k=37, passes=2, seed=202, contigs=[30000], err=0.003, cov=26, rerr=0.0015, hp=40, hp_err=12, cluster=6),
    "k31_dirty": dict(k=31, passes=3, seed=303, contigs=[12000, 7000, 20, 0, 45, 9000], err=0.004, cov=26, rerr=0.0015, hp=30,
                      hp_err=8, cluster=4, n_runs=5, lower=6, iupac=6, diploid=0.004, repeats=(40, 300)),
    "k21_even": dict(k=20, passes=2, seed=404, contigs=[15000, 3000], err=0.004, cov=26, rerr=0.001, hp=20, hp_err=6, cluster=3,
                     lower=2, repeats=(25, 200)),
}


def build_case(name, spec, out_dir):
    rng = np.random.default_rng(spec["seed"])
    k = spec["k"]
    tmp = tempfile.mkdtemp(prefix="golden_" + name)
    haps, records = [], []
    rep = spec.get("repeats")
    for ci, n in enumerate(spec["contigs"]):
        g = synth.random_genome(n, rng, homopolymers=spec.get("hp", 0) if n > 1000 else 0,
                                repeat_copies=rep[0] if rep and n > 5000 else 0, repeat_len=rep[1] if rep else 0) if n else np.zeros(0, np.uint8)
        if n > 200:
            haps.append(g)
            if spec.get("diploid"):
                h2 = g.copy(); m = rng.random(n) < spec["diploid"]
                h2[m] = (h2[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
                haps.append(h2)
        a = g
        if n > 1000:
            a = synth.inject_errors(g, spec["err"], rng)
            a = clustered_errors(a, rng, spec.get("cluster", 0), k)
            a = homopolymer_errors(a, rng, spec.get("hp_err", 0))
        seq = bytearray(synth.codes_to_bytes(a))
        if n > 5000 and ci % 2 == 0:
            seq = decorate(seq, rng, spec.get("n_runs", 0), spec.get("lower", 0), spec.get("iupac", 0))
        records.append(("ctg%d" % ci + (" some description" if ci == 1 else ""), bytes(seq)))
    # haplotypes get coverage proportional to their number (diploid: half each)
    reads = np.concatenate([synth.simulate_reads([h], spec["cov"] / (2 if spec.get("diploid") else 1), 150, spec["rerr"], rng) for h in haps])
    synth.write_reads_fasta(os.path.join(tmp, "reads.fa"), reads)
    synth.write_fasta(os.path.join(tmp, "asm.fa"), records, 80)
    db = os.path.join(tmp, "db.jf")
    subprocess.check_call([JELLYFISH, "count", "-C", "-m", str(k), "-s", str(max(1000, reads.size // 10)), "-t", "4", "-o", db, os.path.join(tmp, "reads.fa")])
    histo = subprocess.check_output([JELLYFISH, "histo", "-t", "2", db]).decode()
    open(os.path.join(tmp, "histo.csv"), "w").write(histo)
    r = subprocess.run([sys.executable, os.path.join(REF_SRC, "jellyfish.py"), os.path.join(tmp, "histo.csv")], capture_output=True, text=True)
    thr_src = "reference jellyfish.py"
    if r.returncode != 0 or not r.stdout.strip():
        thr, thr_src = 3, "fallback (reference jellyfish.py exited %d)" % r.returncode
    else:
        thr = int(r.stdout.strip())
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle"))
    cmd = [sys.executable, os.path.join(REF_SRC, "jasper.py"), "--db", "db.jf", "--query", "asm.fa", "--ksize", str(k), "-p", str(spec["passes"]),
           "--fix", "--fout", "asm.fa.fix.csv", "-ff", "asm.fa.fixed.fa.tmp", "--test", "-thre", str(thr)]
    r = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("reference jasper.py failed on %s:\n%s\n%s" % (name, r.stdout, r.stderr))
    shutil.rmtree(out_dir, ignore_errors=True)
    os.makedirs(os.path.join(out_dir, "expected"))
    shutil.copy(os.path.join(tmp, "asm.fa"), out_dir)
    with open(db, "rb") as f, gzip.GzipFile(os.path.join(out_dir, "db.jf.gz"), "wb", mtime=0) as gz:
        shutil.copyfileobj(f, gz)
    keep = ["_iter%d_asm.fa.fix.csv" % i for i in range(spec["passes"])] + ["_iter%d_asm.fa.fixed.fa.tmp" % (spec["passes"] - 1),
            "0qValCalcHelper.csv", "%dqValCalcHelper.csv" % spec["passes"]]
    for fn in keep:
        shutil.copy(os.path.join(tmp, fn), os.path.join(out_dir, "expected", fn))
    open(os.path.join(out_dir, "histo.csv"), "w").write(histo)
    meta = dict(name=name, k=k, passes=spec["passes"], threshold=thr, threshold_source=thr_src, seed=spec["seed"],
                cmd=" ".join(cmd[1:]), n_records=(os.path.getsize(db)), stdout=r.stdout[-500:])
    json.dump(meta, open(os.path.join(out_dir, "case.json"), "w"), indent=1)
    nrows = [sum(1 for _ in open(os.path.join(tmp, "_iter%d_asm.fa.fix.csv" % i))) - 1 for i in range(spec["passes"])]
    print(name, "threshold", thr, "(%s)" % thr_src, "fix rows per pass", nrows,
          "tallies", open(os.path.join(tmp, "0qValCalcHelper.csv")).read().strip(), "->",
          open(os.path.join(tmp, "%dqValCalcHelper.csv" % spec["passes"])).read().strip())
    # edge-case lookups straight from the binding for this DB
    import dna_jellyfish as jf
    qf = jf.QueryMerFile(db)
    probes = ["", "A", "ACG", "ACNGT" * 10, "acgt" * 12, "N" * 50, "T" * 60, "A" * 60, "ACGT" * 4 + "\n" + "ACGT" * 8]
    asm = b"".join(s for _, s in records).decode()
    for at in list(rng.integers(0, max(1, len(asm) - 2 * k), 60)):
        ln = int(rng.integers(0, 2 * k)); probes.append(asm[int(at):int(at) + ln])
    rows = [[s, str(jf.MerDNA(s).get_canonical()), int(qf[jf.MerDNA(s).get_canonical()])] for s in probes]
    json.dump(rows, open(os.path.join(out_dir, "lookups.json"), "w"))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    if not os.path.exists(JELLYFISH) or not os.path.isdir(REF_SRC):
        raise SystemExit("needs /root/reference and `make -C oracle ref`")
    for name in (sys.argv[1:] or CASES):
        build_case(name, CASES[name], os.path.join(HERE, name))

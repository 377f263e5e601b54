#!/usr/bin/env python
"""BASELINE configs[0] -- the reference's own CPU-runnable case: a 4.6 Mb E. coli-size random genome,
30x simulated 150-bp reads with 0.2 % substitutions, an assembly with 0.1 % injected errors, numpy default_rng(12345) with the
survey's draw order; the database is built by the reference's own `jellyfish count -C -s 30000000 -m 25 -t 8`.

`inputs(dir)` writes reads.fa and asm.fa and is what tests/test_config1.py calls on the GPU box (nothing there reads
/root/reference; oracle/_ref/bin/jellyfish travels with the repo).  Run as a script (build container only: needs
/root/reference, perl, awk, sort), it also runs the UNMODIFIED src/jasper.sh -k 25 -p 2 -t 8 and records what it produced in
tests/golden/config1/expected.json: md5 and size of asm.fa.polished.fasta (records sorted by name: the reference prints perl's
hash order) and of asm.fa.fixes.csv, the threshold, the batch size and the fix-row census (SURVEY: 4,557 rows = 2,347 s /
1,124 i / 1,086 d).

    python tests/golden/make_config1.py
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
K, PASSES, THREADS, JF_SIZE = 25, 2, 8, 30000000


def inputs(out_dir):
    """SURVEY.md Appendix I, same draw order -> (reads.fa path, asm.fa path, assembly length, number of error sites)"""
    rng = np.random.default_rng(12345)
    G, cov, RL, rerr, aerr = 4_600_000, 30, 150, 0.002, 0.001
    g = rng.integers(0, 4, G, dtype=np.uint8)
    n = G * cov // RL
    starts = rng.integers(0, G - RL, n)
    r = g[starts[:, None] + np.arange(RL)]
    e = rng.random(r.shape) < rerr
    r = np.where(e, (r + rng.integers(1, 4, r.shape, dtype=np.uint8)) % 4, r)
    rev = rng.random(n) < 0.5
    r[rev] = (3 - r[rev])[:, ::-1]
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    rows = np.empty((n, 3 + RL + 1), dtype=np.uint8)
    rows[:, :3] = np.frombuffer(b">r\n", dtype=np.uint8)
    rows[:, 3:3 + RL] = lut[r]
    rows[:, 3 + RL] = 10
    reads = os.path.join(out_dir, "reads.fa")
    rows.tofile(reads)
    x = rng.random(G)
    sub = x < aerr * 0.5
    dele = (aerr * 0.5 <= x) & (x < aerr * 0.75)
    ins = (aerr * 0.75 <= x) & (x < aerr)
    a = g.copy()
    a[sub] = (a[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) % 4
    insb = rng.integers(0, 4, G, dtype=np.uint8)
    rep = np.ones(G, dtype=np.int64)
    rep[dele] = 0
    rep[ins] = 2
    a2 = np.repeat(a, rep)
    a2[(np.cumsum(rep) - 1)[ins]] = insb[ins]
    s = lut[a2].tobytes()
    asm = os.path.join(out_dir, "asm.fa")
    with open(asm, "wb") as f:
        f.write(b">ecoli_like\n")
        for i in range(0, len(s), 80):
            f.write(s[i:i + 80] + b"\n")
    return reads, asm, len(s), int(sub.sum() + dele.sum() + ins.sum())


def count_db(jellyfish, reads, db):
    subprocess.check_call([jellyfish, "count", "-C", "-s", str(JF_SIZE), "-m", str(K), "-t", str(THREADS), "-o", db, reads])


def fasta_digest(path):
    """md5 over the records sorted by name, each as 'name\\nsequence\\n' (independent of record order and line width)."""
    recs, name, parts = {}, None, []
    for line in open(path):
        if line.startswith(">"):
            if name is not None:
                recs[name] = "".join(parts)
            name, parts = line[1:].rstrip("\n"), []
        else:
            parts.append(line.rstrip("\n"))
    if name is not None:
        recs[name] = "".join(parts)
    h = hashlib.md5()
    for nm in sorted(recs):
        h.update(("%s\n%s\n" % (nm, recs[nm])).encode())
    return h.hexdigest(), sum(len(v) for v in recs.values())


def fixes_census(path):
    rows = [l.split() for l in open(path, "rb").read().decode("latin-1").split("\n") if l.strip()]
    body = [r for r in rows if not r[0].startswith("Contig")]
    kinds = {"s": 0, "i": 0, "d": 0, "other": 0}
    for r in body:
        f = r[-1].rstrip("\r") if len(r) >= 4 else ""
        kinds[f[0] if f[:1] in ("s", "i", "d") else "other"] += 1
    return len(body), kinds


def main():
    top = tempfile.mkdtemp(prefix="config1_")
    bindir, pydir, run = os.path.join(top, "bin"), os.path.join(top, "py"), os.path.join(top, "run")
    for d in (bindir, pydir, run):
        os.makedirs(d)
    for f in ("jasper.sh", "jasper.py", "jellyfish.py"):
        shutil.copy(os.path.join(REF_SRC, f), os.path.join(bindir, f))
        os.chmod(os.path.join(bindir, f), 0o755)
    shutil.copy(os.path.join(REF_SRC, "jellyfish.py"), os.path.join(pydir, "jellyfish.py"))
    shutil.copytree(os.path.join(ROOT, "oracle", "Bio"), os.path.join(pydir, "Bio"))
    with open(os.path.join(pydir, "dna_jellyfish.py"), "w") as f:
        f.write("import importlib.util\n"
                "spec = importlib.util.spec_from_file_location('_orc_dna_jellyfish', %r)\n"
                "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
                "globals().update({k: getattr(m, k) for k in dir(m) if not k.startswith('__')})\n" % os.path.join(ROOT, "oracle", "dna_jellyfish.py"))
    jf = os.path.join(ROOT, "oracle", "_ref", "bin", "jellyfish")
    env = dict(os.environ, PATH=bindir + ":" + os.path.dirname(jf) + ":" + os.environ["PATH"], PYTHONPATH=pydir, LC_ALL="POSIX")
    reads, asm, asm_len, n_err = inputs(run)
    db = os.path.join(run, "mer_counts25.jf")
    count_db(jf, reads, db)
    import time
    t0 = time.time()
    cmd = ["bash", os.path.join(bindir, "jasper.sh"), "-a", "asm.fa", "-j", "mer_counts25.jf", "-k", str(K), "-t", str(THREADS), "-p", str(PASSES)]
    r = subprocess.run(cmd, cwd=run, env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if r.returncode != 0 or not os.path.exists(os.path.join(run, "asm.fa.polished.fasta")):
        sys.exit("jasper.sh failed:\n%s\n%s" % (r.stdout, r.stderr))
    md5_fa, polished_bp = fasta_digest(os.path.join(run, "asm.fa.polished.fasta"))
    fixes = open(os.path.join(run, "asm.fa.fixes.csv"), "rb").read()
    n_rows, kinds = fixes_census(os.path.join(run, "asm.fa.fixes.csv"))
    bs = [l for l in r.stdout.splitlines() if "Using BATCH SIZE" in l][0].split()[-1]
    out = {"k": K, "passes": PASSES, "threads": THREADS, "jf_size": JF_SIZE, "assembly_bp": asm_len, "error_sites": n_err,
           "jf_records": (os.path.getsize(db) - int(open(db, "rb").read(9))) // 11 if False else None,
           "jf_bytes": os.path.getsize(db), "threshold": open(os.path.join(run, "threshold.txt")).read(), "batch_size": int(bs),
           "polished_fasta_md5_sorted_records": md5_fa, "polished_bp": polished_bp,
           "fixes_csv_md5": hashlib.md5(fixes).hexdigest(), "fixes_csv_bytes": len(fixes), "fix_rows": n_rows, "fix_kinds": kinds,
           "reads_md5": hashlib.md5(open(reads, "rb").read()).hexdigest(), "asm_md5": hashlib.md5(open(asm, "rb").read()).hexdigest(),
           "jasper_sh_wall_s_in_build_container": round(wall, 1), "build_container_cores": len(os.sched_getaffinity(0)),
           "source": "unmodified /root/reference/src/jasper.sh -a asm.fa -j mer_counts25.jf -k 25 -t 8 -p 2 (bc absent: the Q lines are not part of the fixture)"}
    out.pop("jf_records")
    os.makedirs(os.path.join(HERE, "config1"), exist_ok=True)
    json.dump(out, open(os.path.join(HERE, "config1", "expected.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
    shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    main()

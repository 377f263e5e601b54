#!/usr/bin/env python
"""Golden outputs of the UNMODIFIED src/jasper.sh (the whole pipeline: split, histo, threshold, per-batch jasper.py,
join, fixes.csv) for tests of jasper_b200/pipeline.py.

Runs only in the build container: needs /root/reference, `make -C oracle ref` (the reference's jellyfish binary and
binding) and perl / awk / sort.  The reference scripts are copied to a scratch directory under /tmp (never into the
repo), executed with PYTHONPATH pointing at a directory that holds src/jellyfish.py, the binding loader
(oracle/dna_jellyfish.py) and the Bio.pairwise2 stand-in (oracle/Bio), LC_ALL=POSIX.  `bc` is absent here, so the two
"Q value" log lines of the reference read "Inf" and are not part of the fixture.

    python tests/golden/make_golden_sh.py        -> tests/golden/<case>/expected_sh/
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
CASES = {"k31_dirty": {"t": 3}, "k21_even": {"t": 2}}


def main():
    top = tempfile.mkdtemp(prefix="jsh_")
    bindir, pydir = os.path.join(top, "bin"), os.path.join(top, "py")
    os.makedirs(bindir); os.makedirs(pydir)
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
    env = dict(os.environ, PATH=bindir + ":" + os.path.join(ROOT, "oracle", "_ref", "bin") + ":" + os.environ["PATH"], PYTHONPATH=pydir, LC_ALL="POSIX")
    for name, opt in CASES.items():
        cdir = os.path.join(HERE, name)
        case = json.load(open(os.path.join(cdir, "case.json")))
        run = os.path.join(top, name)
        os.makedirs(run)
        shutil.copy(os.path.join(cdir, "asm.fa"), os.path.join(run, "asm.fa"))
        with gzip.open(os.path.join(cdir, "db.jf.gz"), "rb") as src, open(os.path.join(run, "db.jf"), "wb") as dst:
            shutil.copyfileobj(src, dst)
        cmd = ["bash", os.path.join(bindir, "jasper.sh"), "-a", "asm.fa", "-j", "db.jf", "-k", str(case["k"]), "-t", str(opt["t"]), "-p", str(case["passes"])]
        r = subprocess.run(cmd, cwd=run, env=env, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(os.path.join(run, "asm.fa.polished.fasta")):
            sys.exit("jasper.sh failed on %s:\n%s\n%s" % (name, r.stdout, r.stderr))
        out = os.path.join(cdir, "expected_sh")
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        for f in ("asm.fa.polished.fasta", "asm.fa.fixes.csv", "threshold.txt", "jfhisto%d.csv" % case["k"]):
            shutil.copy(os.path.join(run, f), os.path.join(out, f))
        bs = [l for l in r.stdout.splitlines() if "Using BATCH SIZE" in l][0].split()[-1]
        json.dump({"cmd": " ".join(cmd[1:]).replace(bindir + "/", ""), "k": case["k"], "passes": case["passes"], "threads": opt["t"], "batch_size": int(bs)},
                  open(os.path.join(out, "meta.json"), "w"), indent=1)
        print(name, "ok: batch size", bs, "threshold", open(os.path.join(run, "threshold.txt")).read())
    shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    main()

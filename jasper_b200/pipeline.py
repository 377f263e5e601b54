#!/usr/bin/env python
"""jasper.sh in one process: the stages the reference runs as bash / perl / awk / sort around
its per-batch workers (src/jasper.sh:120-262), with the workers replaced by the C-ABI library on the GPUs of the box.

    python -m jasper_b200.pipeline -a asm.fa -j mer_counts.jf -k 37 -p 2 -t 16 [-b BATCH] [-d]

Same options as src/jasper.sh:58-110 (-r / --reads builds the database with the reference's own `jellyfish count`, which
must then be on PATH: DB construction is out of scope here).  Same files written to the working directory:

  jfhisto<K>.csv           `jellyfish histo` of the database                       src/jasper.sh:187-193
  threshold.txt            src/jellyfish.py's threshold, no newline                src/jasper.sh:197-205
  <asm>.polished.fasta     one line per contig, pieces re-stitched                 src/jasper.sh:220
  <asm>.fixes.csv          every fix row of every pass, sorted, header mangled by the awk chain exactly as the reference
                           mangles it                                               src/jasper.sh:222-226
and the two "Q value" log lines (src/jasper.sh:235-257, GNU bc's fixed-point steps emulated with `decimal`).

The intermediate protocol is the reference's too: every polishing call leaves the files a reference worker leaves for its
batch file (_iter<i>_<asm>.batch.<g>.fa.fix.csv, _iter<last>_<asm>.batch.<g>.fa.fixed.fa, lines in *qValCalcHelper.csv),
the join reads them back, and the jasper.split/histo/correct/join.success markers make a rerun skip finished stages
exactly as src/jasper.sh does.  What differs, deliberately: the INPUT batch files are not written (pieces go to the GPUs
from memory, sharded by length over the devices in JASPER_B200_DEVICES, default all, and cut into calls that fit each
device, jb_polish_capacity; the table is read once and replicated over NVLink); contigs come out in input order (the
reference's order is perl's hash order, i.e. random per run).  There is no CPU fallback: without the library and a GPU
this exits 1.
"""
from __future__ import annotations

import argparse
import datetime
import os
import re
import subprocess
import sys
import threading

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from jasper_b200 import batching, bcmath
else:
    from . import batching, bcmath


def log(msg: str) -> None:
    print("[%s] %s" % (datetime.datetime.now().strftime("%a %b %d %H:%M:%S %Z %Y").replace("  ", " "), msg), flush=True)


# ----------------------------------------------------------------------------- the perl / awk stages, restated
def read_assembly(path: str):
    """What perl -ane sees (src/jasper.sh:155): the first whitespace-separated token of every line; a line whose
    token starts with '>' opens a contig named by that token (without '>'), any other token is sequence.
    Returns ([(name, seq)] in file order, total sequence characters as `grep -v '^>' | tr -d '\\n' | wc -c` counts them)."""
    contigs, name, parts, total = [], None, [], 0
    with open(path, "r", newline="") as f:
        for line in f:
            if not line.startswith(">"):
                total += len(line.rstrip("\n")) if line.endswith("\n") else len(line)
            tok = line.split()
            if not tok:
                continue
            if tok[0].startswith(">"):
                if name is not None:
                    contigs.append((name, "".join(parts)))
                name, parts = tok[0][1:], []
            elif name is not None:
                parts.append(tok[0])
            # (sequence in front of the first header is appended to an undefined contig by perl and printed under the
            #  name ":0"-less header; nobody writes such files)
    if name is not None:
        contigs.append((name, "".join(parts)))
    return contigs, total


def merge_fix_rows(rows_per_pass):
    """src/jasper.sh:222-226 on the rows of all passes: rows_per_pass[ite] = [[name, coord, mutation, original]].
    Returns the bytes of <asm>.fixes.csv."""
    import csv
    import io
    texts = []
    for rows in rows_per_pass:
        buf = io.StringIO()
        w = csv.writer(buf, delimiter=" ")                    # the dialect src/jasper.py:114-119 writes with (CRLF ends)
        w.writerow(["Contig", "Base_coord", "Original", "Mutation"])
        w.writerows(rows)
        texts.append(buf.getvalue())
    return merge_fix_csv_texts(texts)


def merge_fix_csv_texts(texts):
    """src/jasper.sh:222-226 on the per-batch fix.csv files (their contents, in `ls` order): `awk 'NR==1 || FNR>1'` keeps the
    first file's header and every file's rows; then the ':' split, the sort and the re-join.  -> bytes of <asm>.fixes.csv."""
    lines = []
    for fi, text in enumerate(texts):
        recs = text.split("\n")
        if recs and recs[-1] == "":
            recs.pop()
        for li, line in enumerate(recs):                      # awk reads \n-terminated records; the \r stays in the last field
            if li == 0 and fi > 0:
                continue
            f = line.split(":")                               # awk -F ':' '{print $1" "$2}'
            lines.append(f[0] + " " + (f[1] if len(f) > 1 else ""))

    def num(s: str) -> float:                                 # sort -n: leading blanks, optional '-', digits, optional fraction
        s = s.lstrip(" \t")
        j = 1 if s[:1] == "-" else 0
        k = j
        while k < len(s) and s[k].isdigit():
            k += 1
        if k < len(s) and s[k] == ".":
            k += 1
            while k < len(s) and s[k].isdigit():
                k += 1
        try:
            return float(s[:k]) if k > j else 0.0
        except ValueError:
            return 0.0

    def fields(line: str):                                    # sort's default fields: a field starts at the blank(s) that precede it
        out, i, n = [], 0, len(line)
        while i < n:
            j = i
            while j < n and line[j] in " \t":
                j += 1
            while j < n and line[j] not in " \t":
                j += 1
            out.append(line[i:j])
            i = j
        return out

    def key(line: str):
        f = fields(line) + ["", "", ""]
        return (f[0].encode("latin-1"), num(f[1]), num(f[2]), line.encode("latin-1"))   # -k1,1 -k2,2n -k3,3n, then the whole line (POSIX bytes)

    out = []
    for line in sorted(lines, key=key):
        f = [x for x in re.split("[ \t\n]+", line) if x != ""]    # awk default splitting ('\r' is not a blank)
        f += [""] * (5 - len(f))
        out.append("%s:%s %s %s %s\n" % (f[0], f[1], f[2], f[3], f[4]))
    return "".join(out).encode("latin-1")


def q_value(wrong: int, total: int, k: int) -> str:
    """src/jasper.sh:239-246: the Q text of the log, through GNU bc's fixed-point rules and math-library series (bcmath.py)."""
    return bcmath.q_value(wrong, total, k)


# ----------------------------------------------------------------------------- the run
def group_by_capacity(indices, lengths, capacity):
    """Consecutive groups of piece indices whose bases stay below `capacity` (None = one group).  What BATCH_SIZE does for the
    reference's workers: it bounds the working set of one polishing call, nothing else."""
    if not capacity:
        return [list(indices)] if indices else []
    groups, cur, used = [], [], 0
    for i in indices:
        if cur and used + lengths[i] > capacity:
            groups.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += lengths[i]
    if cur:
        groups.append(cur)
    return groups


def _batch_file(qfn, g):
    return "%s.batch.%d.fa" % (qfn, g)


def _spill_group(workdir, qfn, g, names, res, passes):
    """The files one reference worker leaves for its batch file (src/jasper.py:107-128 through run_jasper.sh,
    src/jasper.sh:207-208): _iter<i>_<batch>.fix.csv per fixing pass, _iter<last>_<batch>.fixed.fa, and one line appended
    to 0qValCalcHelper.csv / <p>qValCalcHelper.csv."""
    from . import jasper as dropin
    from . import capi
    bf = _batch_file(qfn, g)
    for ite in range(passes):
        rows = res.rows[ite]
        path = os.path.join(workdir, "_iter%d_%s.fix.csv" % (ite, bf))
        if isinstance(rows, list):
            dropin.write_fix_csv(path, rows)
        else:
            with open(path, "wb") as f:
                f.write(b"Contig Base_coord Original Mutation\r\n")
                f.write(capi.rows_csv_bytes(rows, names))
    with open(os.path.join(workdir, "_iter%d_%s.fixed.fa" % (passes - 1, bf)), "wb") as f:
        for name, seq in res.seqs:
            f.write(dropin.fasta_bytes(name, seq.encode("latin-1") if isinstance(seq, str) else seq))
    for ite in (0, passes):
        with open(os.path.join(workdir, "%dqValCalcHelper.csv" % ite), "a") as f:
            f.write("%d %d\n" % tuple(res.tallies[ite if ite == 0 else -1]))


def polish_pieces(pieces, k, thre, passes, qfs, workdir, qfn):
    """pieces [(name, seq)] over the open tables `qfs` (one per GPU): sharded by length over the tables, every shard cut into
    calls that fit the device (jb_polish_capacity), every call's outputs spilled like a reference worker's."""
    lengths = [len(s) for _, s in pieces]
    shards = batching.shard_units(lengths, len(qfs))
    plans, g = [], 0
    for r, qf in enumerate(qfs):
        cap = qf.capacity() if hasattr(qf, "capacity") else None
        groups = group_by_capacity(shards[r], lengths, cap)
        plans.append([(g + j, grp) for j, grp in enumerate(groups)])
        g += len(groups)
    errors = []
    lock = threading.Lock()

    def work(r):
        try:
            for gi, grp in plans[r]:
                seqs = [pieces[i] for i in grp]
                res = qfs[r].polish(seqs, k, thre, passes)
                if res.status != 0:
                    raise RuntimeError("the reference polisher raises on this input (exit 1)")
                with lock:                                    # (the qValCalcHelper files are shared, like the workers' appends)
                    _spill_group(workdir, qfn, gi, [n for n, _ in seqs], res, passes)
        except BaseException as e:                            # noqa: BLE001 -- reported below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(qfs)) if plans[r]]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return g


def _touch(workdir, name):
    open(os.path.join(workdir, name), "a").close()


def _rm(workdir, *names):
    for n in names:
        try:
            os.remove(os.path.join(workdir, n))
        except FileNotFoundError:
            pass


def _glob(workdir, pattern):
    import glob
    return sorted(glob.glob(os.path.join(workdir, pattern)))  # the shell's `ls` / glob order (C locale)


def run(assembly, jf_db, k=37, passes=2, threads=2, batch=0, debug=False, devices=None, open_tables=None, workdir="."):
    """The stages of src/jasper.sh, with its restart markers (jasper.split/histo/correct/join.success: a stage whose marker
    exists is skipped, a stage that runs invalidates the markers behind it, src/jasper.sh:152-158,187-195,214-218,227).
    open_tables(path, devices) -> [table-like objects with .histogram(), .solid_threshold(), .polish(), .close(), optionally
    .capacity()] (default: one jasper_b200.capi.QueryMerFile per device, read once, replicated); tests inject a stand-in."""
    qfn = os.path.basename(assembly)
    if not os.path.isfile(assembly) or os.path.getsize(assembly) == 0:
        raise SystemExit("The query file does not exist. Please supply a valid fasta file to be polished with -a option.")
    if passes < 1:
        raise SystemExit("The number of passes supplied by -p must be a positive integer")
    if k < 1:
        raise SystemExit("The k-mer size supplied by -k must be a positive integer")
    have = lambda name: os.path.exists(os.path.join(workdir, name))
    contigs, total = read_assembly(assembly)
    bs = batching.batch_size(total, threads, batch)
    log("Using BATCH SIZE %d" % bs)
    last = passes - 1
    # ---- split (src/jasper.sh:152-159).  Pieces go to the GPUs from memory: nothing to write, but the marker protocol stays.
    pieces = None
    if not have("jasper.split.success"):
        log("Splitting query into batches for parallel execution")
        _rm(workdir, "jasper.correct.success")
        _touch(workdir, "jasper.split.success")
    if open_tables is None:
        open_tables = open_gpu_tables
    qfs = []
    histo_path = os.path.join(workdir, "jfhisto%d.csv" % k)
    thre = None
    try:
        # ---- histogram (src/jasper.sh:187-193)
        if not have("jasper.histo.success") or not (os.path.isfile(histo_path) and os.path.getsize(histo_path)):
            log("Computing K-mer histogram")
            qfs = qfs or open_tables(jf_db, devices)
            rows = qfs[0].histogram(10000)
            with open(histo_path + ".tmp", "w") as f:
                f.write("".join("%d %d\n" % (c, n) for c, n in rows))
            os.replace(histo_path + ".tmp", histo_path)
            _rm(workdir, "jasper.correct.success")
            _touch(workdir, "jasper.histo.success")
        # ---- polish (src/jasper.sh:195-216)
        if not have("jasper.correct.success"):
            log("Polishing")
            rows = [tuple(int(x) for x in line.split()) for line in open(histo_path) if line.strip()]
            qfs = qfs or open_tables(jf_db, devices)
            try:
                thre = qfs[0].solid_threshold(rows)
            except Exception:
                thre = None
            with open(os.path.join(workdir, "threshold.txt"), "w") as f:
                f.write("" if thre is None else str(thre))
            if thre is None:
                raise SystemExit("Local min of kmer counts is smaller than 4. The input read data is not suitable for polishing.")
            log("Lower threshold for unreliable kmers is %d" % thre)
            for stale in _glob(workdir, "_iter*_%s.batch.*.fa.fix*" % qfn) + _glob(workdir, "*qValCalcHelper.csv"):
                os.remove(stale)                              # (a worker appends to the tallies: leftovers of a failed run would be summed twice)
            pieces = batching.split_contigs(contigs, bs)
            polish_pieces(pieces, k, thre, passes, qfs, workdir, qfn)
            _rm(workdir, "jasper.join.success")
            _touch(workdir, "jasper.correct.success")
    finally:
        for qf in qfs:
            qf.close()
    if thre is None:
        try:
            thre = int(open(os.path.join(workdir, "threshold.txt")).read().strip())
        except (OSError, ValueError):
            thre = None
    # ---- join (src/jasper.sh:218-233)
    if not have("jasper.join.success"):
        log("Joining")
        named = []
        for path in _glob(workdir, "_iter%d_%s.batch.*.fa.fixed.fa" % (last, qfn)):
            name, parts = None, []
            with open(path, "r") as f:
                for line in f:
                    tok = line.split()
                    if not tok:
                        continue
                    if tok[0].startswith(">"):
                        if name is not None:
                            named.append((name, "".join(parts)))
                        name, parts = tok[0][1:], []
                    else:
                        parts.append(tok[0])
            if name is not None:
                named.append((name, "".join(parts)))
        joined = batching.join_pieces(named, bs)
        with open(os.path.join(workdir, qfn + ".fixed.fasta.tmp"), "w") as f:
            seen = set()
            for name, _ in contigs:
                if name in joined and name not in seen:       # (the reference prints perl's hash order)
                    seen.add(name)
                    f.write(">%s\n%s\n" % (name, joined[name]))
        os.replace(os.path.join(workdir, qfn + ".fixed.fasta.tmp"), os.path.join(workdir, qfn + ".polished.fasta"))
        for path in _glob(workdir, "_iter*_%s.batch.*.fa.fixed.fa" % qfn) + _glob(workdir, "_iter*_%s.batch.*.fa.fixed.fa.tmp" % qfn):
            os.remove(path)
        csvs = _glob(workdir, "_iter*_%s.batch.*.fa.fix.csv" % qfn)
        with open(os.path.join(workdir, qfn + ".fixes.csv.tmp"), "wb") as f:
            f.write(merge_fix_csv_texts([open(p, "r", newline="").read() for p in csvs]))
        os.replace(os.path.join(workdir, qfn + ".fixes.csv.tmp"), os.path.join(workdir, qfn + ".fixes.csv"))
        _touch(workdir, "jasper.join.success")
        if not debug:
            for path in csvs:
                os.remove(path)
    # ---- Q values (src/jasper.sh:235-259)
    t4 = [0, 0, 0, 0]
    q0 = q1 = None
    try:
        for j, ite in enumerate((0, passes)):
            for line in open(os.path.join(workdir, "%dqValCalcHelper.csv" % ite)):
                a, b = line.split()
                t4[2 * j] += int(a); t4[2 * j + 1] += int(b)
        q0, q1 = q_value(t4[0], t4[1], k), q_value(t4[2], t4[3], k)
        log("Before Polishing: Q value = %s" % q0)
        log("After Polishing: Q value = %s" % q1)
    except OSError:
        log("Q values: the *qValCalcHelper.csv files of the polishing stage are gone (removed by an earlier complete run)")
    for path in _glob(workdir, "*qValCalcHelper.csv"):
        os.remove(path)
    log("Polished sequence is in %s.polished.fasta" % qfn)
    return {"batch_size": bs, "threshold": thre, "tallies": t4, "q_before": q0, "q_after": q1,
            "pieces": len(pieces) if pieces is not None else None}


class _GpuTable(object):
    def __init__(self, qf):
        self.qf = qf

    def histogram(self, high):
        return self.qf.histogram(high)

    def solid_threshold(self, rows):
        from . import capi
        return capi.solid_threshold(rows)

    def polish(self, seqs, k, thre, passes):
        return self.qf.polish(seqs, k, thre, passes, raw=True)

    def capacity(self):
        env = os.environ.get("JASPER_B200_MAX_CALL_BP")       # tests: force several calls per device
        return int(env) if env else self.qf.polish_capacity()

    def close(self):
        self.qf.close()


def open_gpu_tables(jf_db, devices=None):
    from . import capi
    lib = capi.load_library()
    if devices is None:
        env = os.environ.get("JASPER_B200_DEVICES")
        devices = [int(x) for x in env.split(",")] if env else list(range(max(1, lib.jb_device_count())))
    first = capi.QueryMerFile(jf_db, device=devices[0])
    return [_GpuTable(first)] + [_GpuTable(q) for q in first.replicate(devices[1:])]


def build_parser():
    ap = argparse.ArgumentParser(prog="jasper_b200.pipeline", description="jasper.sh in one process, polishing on the GPUs")
    ap.add_argument("-b", "--batch", type=int, default=0)
    ap.add_argument("-t", "--threads", type=int, default=2)
    ap.add_argument("-a", "--assembly", required=True)
    ap.add_argument("-j", "--jf", default=None)
    ap.add_argument("-r", "--reads", default=None)
    ap.add_argument("-k", "--kmer", type=int, default=37)
    ap.add_argument("-p", "--num_passes", type=int, default=2)
    ap.add_argument("-d", "--debug", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    return ap


def main(argv=None):
    a = build_parser().parse_args(argv)
    jf = a.jf
    if jf is None:
        if not a.reads:
            raise SystemExit("Either a jf database or files of polishing reads must be provided in the argument.")
        jf = "mer_counts%d.jf" % a.kmer                       # src/jasper.sh:170-182, with the reference binary
        if not (os.path.isfile(jf) and os.path.getsize(jf)):
            log("Creating jellyfish database %s" % jf)
            size = sum(os.path.getsize(p) for p in a.reads.split()) // 10
            cmd = "zcat -f %s | jellyfish count -C -s %d -m %d -o %s -t %d /dev/stdin" % (a.reads, max(size, 1), a.kmer, jf, a.threads)
            if subprocess.call(["bash", "-c", "set -o pipefail; " + cmd]) != 0:
                raise SystemExit("jellyfish count failed (building the database stays with the reference binary, which must be on PATH)")
        else:
            log("Using existing jellyfish database %s" % jf)
    try:
        run(a.assembly, jf, a.kmer, a.num_passes, a.threads, a.batch, a.debug)
    except SystemExit:
        raise
    except BaseException as e:                                # noqa: BLE001
        print("Polishing failed: %r" % (e,), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()

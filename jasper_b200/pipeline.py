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

What differs, deliberately: nothing is spilled to batch files (pieces go to the GPUs from memory, sharded by length
over the devices in JASPER_B200_DEVICES, default all; the table is read once and replicated over NVLink); contigs come
out in input order (the reference's order is perl's hash order, i.e. random per run); the jasper.*.success restart
markers are not written.  There is no CPU fallback: without the library and a GPU this exits 1.
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
    from jasper_b200 import batching
else:
    from . import batching


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
    lines = []
    buf = io.StringIO()
    w = csv.writer(buf, delimiter=" ")                        # the dialect src/jasper.py:114-119 writes with (CRLF ends)
    w.writerow(["Contig", "Base_coord", "Original", "Mutation"])
    for rows in rows_per_pass:
        w.writerows(rows)
    for line in buf.getvalue().split("\n")[:-1]:              # awk reads \n-terminated records; the \r stays in the last field
        f = line.split(":")                                   # awk -F ':' '{print $1" "$2}'
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
    """src/jasper.sh:239-246: GNU bc's scale-limited steps, digit for digit (truncation, no rounding)."""
    from decimal import ROUND_DOWN, Decimal, getcontext
    getcontext().prec = 120

    def trunc(x, scale):
        return x.quantize(Decimal(1).scaleb(-scale), rounding=ROUND_DOWN)

    pgood = Decimal(1) - trunc(Decimal(wrong) / Decimal(total), 10)
    inv_k = trunc(Decimal(1) / Decimal(k), 50)
    err = Decimal(1) - trunc(trunc(trunc(pgood.ln(), 50) * inv_k, 50).exp(), 50)
    if not err > 0:
        return "Inf"
    q = trunc(trunc(Decimal(-10) * trunc(err.ln(), 5), 5) / trunc(Decimal(10).ln(), 5), 5)
    s = format(q, "f")
    if s.startswith("0."):
        s = s[1:]
    elif s.startswith("-0."):
        s = "-" + s[2:]
    return s


# ----------------------------------------------------------------------------- the run
def polish_pieces(pieces, k, thre, passes, qfs):
    """pieces [(name, seq)] over the open tables `qfs` (one per GPU): (tallies first/last summed, rows per pass, seqs in piece order)."""
    shards = batching.shard_units([len(s) for _, s in pieces], len(qfs))
    results = [None] * len(qfs)
    errors = []

    def work(r):
        try:
            results[r] = qfs[r].polish([pieces[i] for i in shards[r]], k, thre, passes)
        except BaseException as e:                            # noqa: BLE001 -- reported below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(qfs)) if shards[r]]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    seqs = [None] * len(pieces)
    rows = [[] for _ in range(passes)]
    t4 = [0, 0, 0, 0]
    for r, res in enumerate(results):
        if res is None:
            continue
        if res.status != 0:
            raise RuntimeError("the reference polisher raises on this input (exit 1)")
        for j, i in enumerate(shards[r]):
            seqs[i] = res.seqs[j]
        for ite in range(passes):
            rows[ite].extend(res.rows[ite])
        t4[0] += res.tallies[0][0]; t4[1] += res.tallies[0][1]
        t4[2] += res.tallies[-1][0]; t4[3] += res.tallies[-1][1]
    return t4, rows, seqs


def run(assembly, jf_db, k=37, passes=2, threads=2, batch=0, debug=False, devices=None, open_tables=None, workdir="."):
    """The stages of src/jasper.sh.  open_tables(path, devices) -> [table-like objects with .histogram() and .polish()]
    (default: one jasper_b200.capi.QueryMerFile per device, read once, replicated); tests inject a stand-in."""
    qfn = os.path.basename(assembly)
    if not os.path.isfile(assembly) or os.path.getsize(assembly) == 0:
        raise SystemExit("The query file does not exist. Please supply a valid fasta file to be polished with -a option.")
    if passes < 1:
        raise SystemExit("The number of passes supplied by -p must be a positive integer")
    if k < 1:
        raise SystemExit("The k-mer size supplied by -k must be a positive integer")
    contigs, total = read_assembly(assembly)
    bs = batching.batch_size(total, threads, batch)
    log("Using BATCH SIZE %d" % bs)
    log("Splitting query into batches for parallel execution")
    pieces = batching.split_contigs(contigs, bs)
    if open_tables is None:
        open_tables = open_gpu_tables
    qfs = open_tables(jf_db, devices)
    try:
        log("Computing K-mer histogram")
        rows = qfs[0].histogram(10000)
        with open(os.path.join(workdir, "jfhisto%d.csv" % k), "w") as f:
            f.write("".join("%d %d\n" % (c, n) for c, n in rows))
        log("Polishing")
        try:
            thre = qfs[0].solid_threshold(rows)
        except Exception:
            thre = None
        with open(os.path.join(workdir, "threshold.txt"), "w") as f:
            f.write("" if thre is None else str(thre))
        if thre is None:
            raise SystemExit("Local min of kmer counts is smaller than 4. The input read data is not suitable for polishing.")
        log("Lower threshold for unreliable kmers is %d" % thre)
        t4, fix_rows, seqs = polish_pieces(pieces, k, thre, passes, qfs)
    finally:
        for qf in qfs:
            qf.close()
    log("Joining")
    joined = batching.join_pieces(seqs, bs)
    with open(os.path.join(workdir, qfn + ".polished.fasta"), "w") as f:
        seen = set()
        for name, _ in contigs:
            if name in joined and name not in seen:           # (the reference prints perl's hash order)
                seen.add(name)
                f.write(">%s\n%s\n" % (name, joined[name]))
    with open(os.path.join(workdir, qfn + ".fixes.csv"), "wb") as f:
        f.write(merge_fix_rows(fix_rows))
    q0, q1 = q_value(t4[0], t4[1], k), q_value(t4[2], t4[3], k)
    log("Before Polishing: Q value = %s" % q0)
    log("After Polishing: Q value = %s" % q1)
    log("Polished sequence is in %s.polished.fasta" % qfn)
    if not debug:
        pass                                                  # nothing was spilled: no batch files to remove
    return {"batch_size": bs, "threshold": thre, "tallies": t4, "q_before": q0, "q_after": q1, "pieces": len(pieces)}


class _GpuTable(object):
    def __init__(self, qf):
        self.qf = qf

    def histogram(self, high):
        return self.qf.histogram(high)

    def solid_threshold(self, rows):
        from . import capi
        return capi.solid_threshold(rows)

    def polish(self, seqs, k, thre, passes):
        return self.qf.polish(seqs, k, thre, passes)

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

#!/usr/bin/env python
"""Drop-in for the reference's src/jasper.py: same command line, same files written, same exit codes --
the per-batch loop itself (src/jasper.py:12-32, 35-137, 150-599) runs on the GPU behind one C-ABI call
(jb_polish, include/jasper_b200.h) instead of ~1e5 SWIG lookups per megabase.

    jasper.py --db DB.jf --query BATCH.fa --ksize K -p P --fix --fout BATCH.fix.csv \
              -ff BATCH.fixed.fa.tmp --test -thre T          (exactly what src/jasper.sh:207 runs)

Files (cwd, like the reference):
  _iter{0..P-1}_<fout>        fix rows, csv.writer(delimiter=' ') with CRLF ends      (src/jasper.py:114-119)
  _iter{P-1}_<fixedfasta>     polished FASTA, 60 columns                              (src/jasper.py:120-128)
  0qValCalcHelper.csv, {P}qValCalcHelper.csv   "wrong total\n" appended               (src/jasper.py:107-111)
Exit 1 where the reference's bare `except` exits 1 (bad DB, crash paths); there is no CPU fallback.
GPU selection comes from the environment (JASPER_B200_DEVICE, default 0): the CLI stays the reference's.
"""
from __future__ import annotations

import argparse
import csv
import os
import sys

if __package__ in (None, ""):                      # run as a script next to jasper.sh
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from jasper_b200 import capi
else:
    from . import capi


def parse_fasta(query_file: str):
    """src/jasper.py:615-631 -- dict semantics included: a repeated name keeps its first position and its
    last sequence; text before the first header is dropped; only '\\n' is stripped from sequence lines."""
    seq = {}
    name = None
    parts = []
    with open(query_file, "r") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    seq[name] = "".join(parts)
                name = line.split()[0][1:]
                parts = []
            else:
                parts.append(line.replace("\n", ""))
    if name is not None:
        seq[name] = "".join(parts)
    return seq


def _prefixed(path: str, ite: int) -> str:
    head, tail = os.path.split(path)               # the reference drops the separator: only bare names work
    return head + "_iter" + str(ite) + "_" + tail


def write_fix_csv(path: str, rows) -> None:
    with open(path, "w") as csvf:                  # same writer, same dialect as src/jasper.py:114-119
        w = csv.writer(csvf, delimiter=" ")
        w.writerow(["Contig", "Base_coord", "Original", "Mutation"])
        w.writerows(rows)


def write_fasta(path: str, seqs, width: int = 60) -> None:
    with open(path, "w") as of:
        for name, s in seqs:
            of.write(">{}\n".format(name))
            for i in range(0, len(s), width):
                of.write(s[i:i + width] + "\n")


def fasta_bytes(name: str, seq, width: int = 60) -> bytes:
    """One record as src/jasper.py:120-128 writes it (split_output: `width` bases per line), built in bulk.  seq: bytes-like."""
    head = b">" + name.encode("utf-8") + b"\n"
    full = len(seq) // width * width
    try:
        import numpy as np
    except ImportError:
        return head + b"".join(bytes(seq[i:i + width]) + b"\n" for i in range(0, len(seq), width))
    body = b""
    if full:
        a = np.empty((full // width, width + 1), dtype=np.uint8)
        a[:, :width] = np.frombuffer(seq, dtype=np.uint8, count=full).reshape(-1, width)
        a[:, width] = 10
        body = a.tobytes()
    tail = bytes(seq[full:]) + b"\n" if len(seq) > full else b""
    return head + body + tail


def write_outputs_raw(res, names, num_iter, test, fix, fout, fixedout) -> None:
    """The files of one src/jasper.py run from a raw PolishResult (capi._collect(raw=True)), byte for byte."""
    for ite in range(num_iter + 1):
        it_fix = fix and ite != num_iter
        if test and (ite == 0 or ite == num_iter):
            with open(str(ite) + "qValCalcHelper.csv", "a") as f:
                f.write("{} {}\n".format(*res.tallies[ite]))
        if it_fix:
            with open(_prefixed(fout, ite), "wb") as f:
                f.write(b"Contig Base_coord Original Mutation\r\n")
                f.write(capi.rows_csv_bytes(res.rows[ite], names))
        if ite == num_iter and fix:
            with open(_prefixed(fixedout, ite - 1), "wb") as of:
                for name, s in res.seqs:
                    of.write(fasta_bytes(name, s))


def main(contigs, query_path, k, test, fix, fout, fixedout, db, thre, num_iter):
    try:
        device = int(os.environ.get("JASPER_B200_DEVICE", "0"))
        qf = capi.QueryMerFile(db, device=device)
        seq_dict = parse_fasta(query_path)
        seqs = list(seq_dict.items())
        res = qf.polish(seqs, k, thre, num_iter, fix=bool(fix), raw=True)
        if res.status != 0:
            raise capi.JasperError(res.status, "the reference polisher raises on this input (exit 1)")
        write_outputs_raw(res, [n for n, _ in seqs], num_iter, test, fix, fout, fixedout)
        qf.close()
    except SystemExit:
        raise
    except BaseException:
        print(sys.exc_info())                      # the reference prints the line number and exc_info, then exits 1
        sys.exit(1)


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--db", default=None, help="The path to the .jf  database file. Not needed if --contigs is given.")
    parser.add_argument("--reads", nargs="+", default=None, help="The path to the .fasta file(s) containing the contigs to build the jellyfish database. Not needed if --db is provided")
    parser.add_argument("-q", "--query", help="The path to the .fasta query file")
    parser.add_argument("-thre", "--threshold", type=int, default=None, help="The threshold for an unreliable kmer.")
    parser.add_argument("-k", "--ksize", type=int, help="The kmer size")
    parser.add_argument("--test", action="store_true", help="Ouput the total num of bad kmers, and an estimation for Q value")
    parser.add_argument("--fix", action="store_true", help="Output the index of fixed bases and output the new sequence")
    parser.add_argument("--fout", default="fout.csv", help="The path to output the index of the fixed bases.")
    parser.add_argument("-ff", "--fixedfasta", default="fixed_seq.fasta", help="The path to output the fixed assembly sequences")
    parser.add_argument("-p", "--passes", type=int, default=2, help="The number of iterations of fixing.")
    return parser


if __name__ == "__main__":
    args = build_parser().parse_args()
    main(args.reads, args.query, args.ksize, args.test, args.fix, args.fout, args.fixedfasta, args.db, args.threshold, args.passes)

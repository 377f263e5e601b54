"""Synthetic inputs for tests and bench.py: genomes, assemblies with injected errors,
simulated reads, and a direct writer for Jellyfish "binary/sorted" `.jf` files.

Tooling, not product path: nothing here is called by jasper_b200.jasper or the C-ABI
library.  The `.jf` layout follows SURVEY.md Appendix B (checked against files written by
the reference `jellyfish count`, JF!include/jellyfish/binary_dumper.hpp:36-40,76 and
JF!include/jellyfish/generic_file_header.hpp:88-111); files written here are read back by
the reference's own `QueryMerFile` / `jellyfish dump` in tests/test_oracle.py.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
U64 = np.uint64


# ----------------------------------------------------------------------------- sequences
def codes_to_bytes(codes: np.ndarray) -> bytes:
    return BASES[codes].tobytes()


def bytes_to_codes(seq: bytes) -> np.ndarray:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
        lut[c + 32] = i
    return lut[np.frombuffer(seq, dtype=np.uint8)]


def random_genome(n: int, rng: np.random.Generator, homopolymers: int = 0, hp_len=(4, 12),
                  repeat_copies: int = 0, repeat_len: int = 400) -> np.ndarray:
    """iid uniform ACGT codes; optionally plant homopolymer runs and a repeat family."""
    g = rng.integers(0, 4, n, dtype=np.uint8)
    for _ in range(homopolymers):
        ln = int(rng.integers(hp_len[0], hp_len[1] + 1))
        at = int(rng.integers(0, max(1, n - ln)))
        g[at:min(n, at + ln)] = rng.integers(0, 4)
    if repeat_copies and repeat_len < n:
        unit = rng.integers(0, 4, repeat_len, dtype=np.uint8)
        for _ in range(repeat_copies):
            at = int(rng.integers(0, max(1, n - repeat_len)))
            g[at:at + repeat_len] = unit
    return g


def inject_errors(g: np.ndarray, rate: float, rng: np.random.Generator,
                  mix=(0.5, 0.25, 0.25)) -> np.ndarray:
    """Assembly errors per position: substitution / 1-bp deletion / 1-bp insertion
    (SURVEY.md Appendix I recipe, same draw order)."""
    n = len(g)
    x = rng.random(n)
    s, d, _ = mix
    sub = x < rate * s
    dele = (rate * s <= x) & (x < rate * (s + d))
    ins = (rate * (s + d) <= x) & (x < rate)
    a = g.copy()
    a[sub] = (a[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) % 4
    insb = rng.integers(0, 4, n, dtype=np.uint8)
    rep = np.ones(n, dtype=np.int64)
    rep[dele] = 0
    rep[ins] = 2
    a2 = np.repeat(a, rep)
    a2[(np.cumsum(rep) - 1)[ins]] = insb[ins]
    return a2


def simulate_reads(haps, coverage: float, read_len: int, err: float, rng: np.random.Generator) -> np.ndarray:
    """n x read_len code matrix; reads drawn uniformly from the haplotypes, substitution errors,
    half reverse-complemented."""
    out = []
    for g in haps:
        G = len(g)
        n = int(G * coverage / len(haps)) // read_len
        starts = rng.integers(0, G - read_len, n)
        r = g[starts[:, None] + np.arange(read_len)]
        e = rng.random(r.shape) < err
        r = np.where(e, (r + rng.integers(1, 4, r.shape, dtype=np.uint8)) % 4, r).astype(np.uint8)
        rev = rng.random(n) < 0.5
        r[rev] = (3 - r[rev])[:, ::-1]
        out.append(r)
    return np.concatenate(out, axis=0)


def write_reads_fasta(path: str, reads: np.ndarray) -> None:
    n, L = reads.shape
    buf = np.empty((n, L + 4), dtype=np.uint8)
    buf[:, 0] = ord(">")
    buf[:, 1] = ord("r")
    buf[:, 2] = ord("\n")
    buf[:, 3:3 + L] = BASES[reads]
    buf[:, 3 + L] = ord("\n")
    with open(path, "wb") as f:
        f.write(buf.tobytes())


def write_fasta(path: str, records, width: int = 80) -> None:
    """records: iterable of (name, bytes)."""
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")


# ----------------------------------------------------------------------------- k-mers
def kmer_keys(codes: np.ndarray, k: int):
    """All k-mers of a code array as (lo, hi) uint64 pairs; first base most significant
    (JF!include/jellyfish/mer_dna.hpp:525-542)."""
    n = len(codes) - k + 1
    if n <= 0:
        z = np.zeros(0, dtype=U64)
        return z, z.copy()
    c = codes.astype(U64)
    lo = np.zeros(n, dtype=U64)
    hi = np.zeros(n, dtype=U64)
    for t in range(k):
        sh = 2 * (k - 1 - t)
        if sh < 64:
            lo |= c[t:t + n] << U64(sh)
        else:
            hi |= c[t:t + n] << U64(sh - 64)
    return lo, hi


def revcomp_keys(codes: np.ndarray, k: int):
    n = len(codes) - k + 1
    if n <= 0:
        z = np.zeros(0, dtype=U64)
        return z, z.copy()
    c = (3 - codes).astype(U64)
    lo = np.zeros(n, dtype=U64)
    hi = np.zeros(n, dtype=U64)
    for t in range(k):
        sh = 2 * t
        if sh < 64:
            lo |= c[t:t + n] << U64(sh)
        else:
            hi |= c[t:t + n] << U64(sh - 64)
    return lo, hi


def canonical_keys(codes: np.ndarray, k: int):
    flo, fhi = kmer_keys(codes, k)
    rlo, rhi = revcomp_keys(codes, k)
    take_r = (rhi < fhi) | ((rhi == fhi) & (rlo < flo))
    return np.where(take_r, rlo, flo), np.where(take_r, rhi, fhi)


def unique_keys(lo: np.ndarray, hi: np.ndarray, weights: np.ndarray | None = None):
    """Distinct (lo, hi) pairs and their summed weights."""
    order = np.lexsort((lo, hi))
    lo, hi = lo[order], hi[order]
    w = np.ones(len(lo), dtype=np.int64) if weights is None else weights[order].astype(np.int64)
    if len(lo) == 0:
        return lo, hi, w
    new = np.ones(len(lo), dtype=bool)
    new[1:] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
    idx = np.flatnonzero(new)
    return lo[idx], hi[idx], np.add.reduceat(w, idx)


# ----------------------------------------------------------------------------- .jf writer
def low_block_rank(cols: np.ndarray, r: int) -> int:
    """GF(2) rank of the r x r block over the LOW r key bits (the last r columns)."""
    m = [int(x) & ((1 << r) - 1) for x in cols[len(cols) - r:]]
    rank = 0
    for bit in range(r - 1, -1, -1):
        p = next((i for i in range(rank, len(m)) if (m[i] >> bit) & 1), None)
        if p is None:
            continue
        m[rank], m[p] = m[p], m[rank]
        for i in range(len(m)):
            if i != rank and (m[i] >> bit) & 1:
                m[i] ^= m[rank]
        rank += 1
    return rank


def random_matrix(r: int, c: int, rng: np.random.Generator, jellyfish_like: bool = True) -> np.ndarray:
    """c columns of r random bits each.  `jellyfish count` redraws until the block over the low r key bits is
    invertible (randomize_pseudo_inverse, JF!lib/rectangular_binary_matrix.cc:160-204,240-247), which is what lets the
    compact table layout keep key >> r only; jellyfish_like=False returns the first draw whose low block is SINGULAR
    (any matrix is readable: such a file makes the loader fall back to the wide layout)."""
    while True:
        hi = rng.integers(0, 1 << 32, c, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, c, dtype=np.uint64)
        cols = ((hi << U64(32)) | lo) & U64((1 << r) - 1)
        if r > c or (low_block_rank(cols, r) == r) == jellyfish_like:
            return cols


def matrix_times(columns: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    """pos = XOR_{b set in key} columns[c-1-b]  (times_loop,
    JF!include/jellyfish/rectangular_binary_matrix.hpp:223-261), vectorised with byte LUTs."""
    c = len(columns)
    res = np.zeros(len(lo), dtype=U64)
    nbytes = (c + 7) // 8
    for byte in range(nbytes):
        lut = np.zeros(256, dtype=U64)
        for bit in range(8):
            b = byte * 8 + bit
            if b >= c:
                break
            col = columns[c - 1 - b]
            sel = ((np.arange(256) >> bit) & 1).astype(bool)
            lut[sel] ^= col
        word = lo if byte < 8 else hi
        sh = U64(8 * (byte % 8))
        res ^= lut[((word >> sh) & U64(0xFF)).astype(np.int64)]
    return res


def write_jf(path: str, lo: np.ndarray, hi: np.ndarray, counts: np.ndarray, k: int, r: int | None = None,
             seed: int = 7, counter_len: int = 4, canonical: bool = True, identity: bool = False,
             jellyfish_like: bool = True) -> dict:
    """Write distinct keys + counts as a Jellyfish binary/sorted file.  Returns the header dict."""
    n = len(lo)
    if r is None:
        r = max(4, int(np.ceil(np.log2(max(16, n) / 0.7))))
    r = min(r, 2 * k)
    rng = np.random.default_rng(seed)
    cols = random_matrix(r, 2 * k, rng, jellyfish_like)
    mask = U64((1 << r) - 1)
    pos = (lo & mask) if identity else (matrix_times(cols, lo, hi) & mask)
    order = np.lexsort((lo, hi, pos))
    lo, hi, counts = lo[order], hi[order], np.asarray(counts)[order]
    key_bytes = (2 * k + 7) // 8
    rec = np.zeros((n, key_bytes + counter_len), dtype=np.uint8)
    lo_b = lo.view(np.uint8).reshape(n, 8)
    hi_b = hi.view(np.uint8).reshape(n, 8)
    rec[:, :min(8, key_bytes)] = lo_b[:, :min(8, key_bytes)]
    if key_bytes > 8:
        rec[:, 8:key_bytes] = hi_b[:, :key_bytes - 8]
    cmax = (1 << (8 * counter_len)) - 1
    cnt = np.minimum(np.asarray(counts, dtype=np.uint64), U64(cmax))
    rec[:, key_bytes:] = cnt.view(np.uint8).reshape(n, 8)[:, :counter_len]
    header = {
        "alignment": 8, "canonical": bool(canonical),
        "cmdline": ["synth", "-m", str(k)], "counter_len": counter_len,
        "exe_path": "jasper_b200.synth", "format": "binary/sorted", "hostname": "synthetic",
        "key_len": 2 * k,
        "matrix1": {"c": 2 * k, "columns": [int(x) for x in cols], "identity": bool(identity), "r": r},
        "max_reprobe": 126, "pwd": "/", "reprobes": [1], "size": 1 << r,
        "time": "synthetic", "val_len": 7,
    }
    js = json.dumps(header, separators=(",", ":")).encode()
    pad = (-(9 + len(js))) % 8
    js += b"\0" * pad
    with open(path, "wb") as f:
        f.write(b"%09d" % len(js))
        f.write(js)
        f.write(rec.tobytes())
    return header


def kmer_db_from_reads(reads: np.ndarray, k: int):
    """Exact canonical k-mer counts of a read matrix (what `jellyfish count -C` computes)."""
    n, L = reads.shape
    los, his = [], []
    for t0 in range(0, n, 20000):
        blk = reads[t0:t0 + 20000]
        flat = blk.reshape(-1)
        lo, hi = canonical_keys(flat, k)
        valid = (np.arange(len(flat) - k + 1) % L) <= L - k
        los.append(lo[valid])
        his.append(hi[valid])
    lo = np.concatenate(los)
    hi = np.concatenate(his)
    return unique_keys(lo, hi)


# ----------------------------------------------------------------------------- test cases
@dataclass
class Case:
    """One polishing test case on disk."""
    name: str
    k: int
    passes: int
    threshold: int
    fasta: str
    db: str
    truth: list = field(default_factory=list)


def solid_threshold_from_histogram(rows) -> int | None:
    """Host-side restatement of src/jellyfish.py:1-22 (threshold = int(count/2) at the first
    local minimum of the count histogram; None if that is < 2)."""
    last = -1
    thr = 0
    for cnt, freq in rows:
        if last == -1:
            last = freq
        elif last >= freq:
            last = freq
            thr = int(cnt / 2)
        else:
            return thr if thr >= 2 else None
    return None


def histogram_rows(counts: np.ndarray, high: int = 10000):
    """`jellyfish histo` rows (JF!sub_commands/histo_main.cc:64-84): count, frequency for
    frequency > 0, counts above `high` lumped into high+1."""
    c = np.minimum(np.asarray(counts, dtype=np.int64), high + 1)
    h = np.bincount(c, minlength=high + 2)
    return [(i, int(h[i])) for i in range(1, high + 2) if h[i] > 0]

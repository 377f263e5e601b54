"""Large synthetic workloads for bench.py, generated with torch (on the GPU when there is one):
a random genome, an assembly of it with injected errors, and the Jellyfish `.jf` database a 30x
Illumina read set of that genome would give -- written DIRECTLY in binary/sorted format
(each genome k-mer count ~ Poisson(cov*(L-k+1)/L), plus E singleton error k-mers).

Bench tooling only: PyTorch is used here for data preparation (sorting a few hundred million
keys takes milliseconds on the GPU), never on the product path.  The file layout is the one
synth.write_jf documents and tests/test_oracle.py reads back with the reference's own reader.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import synth

MIN64 = -(1 << 63)


def _keys_from_codes(g: torch.Tensor, k: int):
    """forward (lo, hi) and reverse-complement (lo, hi) keys of every window, as int64 bit patterns."""
    n = g.numel() - k + 1
    flo = torch.zeros(n, dtype=torch.int64, device=g.device)
    fhi = torch.zeros_like(flo)
    rlo = torch.zeros_like(flo)
    rhi = torch.zeros_like(flo)
    c = 3 - g
    for t in range(k):
        sh = 2 * (k - 1 - t)
        piece = g[t:t + n]
        if sh < 64:
            flo |= piece << sh
        else:
            fhi |= piece << (sh - 64)
        sh = 2 * t
        piece = c[t:t + n]
        if sh < 64:
            rlo |= piece << sh
        else:
            rhi |= piece << (sh - 64)
    return flo, fhi, rlo, rhi


def _ult(alo, ahi, blo, bhi):
    """(ahi, alo) < (bhi, blo) as unsigned 128-bit."""
    return (ahi < bhi) | ((ahi == bhi) & ((alo ^ MIN64) < (blo ^ MIN64)))


def _revcomp(lo, hi, k):
    rlo = torch.zeros_like(lo)
    rhi = torch.zeros_like(lo)
    for t in range(k):
        src = 2 * t
        v = ((lo >> src) if src < 64 else (hi >> (src - 64))) & 3
        v = 3 - v
        dst = 2 * (k - 1 - t)
        if dst < 64:
            rlo |= v << dst
        else:
            rhi |= v << (dst - 64)
    return rlo, rhi


def _sort_pairs(lo, hi, *extra):
    """Stable lexicographic sort by unsigned (hi, lo)."""
    o = torch.sort(lo ^ MIN64, stable=True).indices
    lo, hi = lo[o], hi[o]
    extra = [e[o] for e in extra]
    o = torch.sort(hi, stable=True).indices          # hi < 2^32: signed order == unsigned order
    return (lo[o], hi[o]) + tuple(e[o] for e in extra)


def _unique_sum(lo, hi, w):
    lo, hi, w = _sort_pairs(lo, hi, w)
    new = torch.ones(lo.numel(), dtype=torch.bool, device=lo.device)
    new[1:] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
    idx = torch.nonzero(new).squeeze(1)
    csum = torch.cumsum(w, 0)
    ends = torch.cat([idx[1:], torch.tensor([lo.numel()], device=lo.device)]) - 1
    tot = csum[ends]
    tot[1:] -= csum[ends[:-1]]
    return lo[idx], hi[idx], tot


DIRT = {"repeat_copies": 150, "repeat_len": 400, "repeat_divergence": 0.02, "homopolymer_every": 2000, "homopolymer_len": (6, 16),
        "heterozygosity": 0.004, "n_runs": 24, "n_run_len": (20, 400), "lower_stretches": 200, "lower_len": (200, 2000)}


def dirty_genome(G: int, seed: int, dev) -> torch.Tensor:
    """A random genome with what real assemblies hold and iid bases do not: a repeat family (DIRT["repeat_copies"] diverged
    copies of one unit) and homopolymer runs.  int64 codes on `dev`."""
    rng = np.random.default_rng(seed + 77)
    g = rng.integers(0, 4, G, dtype=np.uint8)
    L = DIRT["repeat_len"]
    unit = rng.integers(0, 4, L, dtype=np.uint8)
    for p in rng.integers(0, max(1, G - L), DIRT["repeat_copies"]):
        c = unit.copy()
        m = rng.random(L) < DIRT["repeat_divergence"]
        c[m] = (c[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
        g[p:p + L] = c[:max(0, min(L, G - p))]
    lo, hi = DIRT["homopolymer_len"]
    for p in rng.integers(0, max(1, G - hi), max(1, G // DIRT["homopolymer_every"])):
        g[p:p + int(rng.integers(lo, hi))] = rng.integers(0, 4)
    return torch.from_numpy(g).to(dev).to(torch.int64)


def make_workload(out_dir: str, genome_bp: int, k: int, seed: int, coverage: float = 30.0, read_len: int = 150,
                  err_kmers_per_bp: float = 2.2, device: str | None = None, r_bits: int | None = None, dirty: bool = False) -> dict:
    """Writes out_dir/db.jf and returns {"genome": uint8 codes (cpu numpy), "db": path, "n_records": N, ...}.
    dirty: the genome carries a repeat family and homopolymer runs (dirty_genome) and the reads come from TWO haplotypes that
    differ at DIRT["heterozygosity"] of the sites, half of the coverage each; the assembly is made from the first."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    G = int(genome_bp)
    lam = coverage * (read_len - k + 1) / read_len
    if dirty:
        g = dirty_genome(G, seed, dev)
        h2 = g.clone()
        het = torch.rand(G, device=dev, generator=gen) < DIRT["heterozygosity"]
        h2[het] = (h2[het] + torch.randint(1, 4, (int(het.sum()),), device=dev, generator=gen)) & 3
        flo, fhi, rlo, rhi = _keys_from_codes(g, k)              # (error k-mers below are cut from the first haplotype)
        take = _ult(rlo, rhi, flo, fhi)
        clo = torch.where(take, rlo, flo)
        chi = torch.where(take, rhi, fhi)
        w = torch.poisson(torch.full((clo.numel(),), lam / 2, device=dev), generator=gen).to(torch.int64)
        f2lo, f2hi, r2lo, r2hi = _keys_from_codes(h2, k)
        take = _ult(r2lo, r2hi, f2lo, f2hi)
        clo = torch.cat([clo, torch.where(take, r2lo, f2lo)]); chi = torch.cat([chi, torch.where(take, r2hi, f2hi)])
        w = torch.cat([w, torch.poisson(torch.full((f2lo.numel(),), lam / 2, device=dev), generator=gen).to(torch.int64)])
        del h2, het, f2lo, f2hi, r2lo, r2hi
        keep = w > 0                                             # (a window no read covers is not in the database)
        clo, chi, w = clo[keep], chi[keep], w[keep]
        del keep
    else:
        g = torch.randint(0, 4, (G,), dtype=torch.int64, device=dev, generator=gen)
        flo, fhi, rlo, rhi = _keys_from_codes(g, k)
        take = _ult(rlo, rhi, flo, fhi)
        clo = torch.where(take, rlo, flo)
        chi = torch.where(take, rhi, fhi)
        w = torch.poisson(torch.full((clo.numel(),), lam, device=dev), generator=gen).to(torch.int64).clamp_(min=1)
    # singleton error k-mers: a genome window with one base substituted
    E = int(err_kmers_per_bp * G)
    if E > 0:
        p = torch.randint(0, flo.numel(), (E,), device=dev, generator=gen)
        b = torch.randint(0, k, (E,), device=dev, generator=gen)
        d = torch.randint(1, 4, (E,), device=dev, generator=gen)
        elo, ehi = flo[p], fhi[p]
        sh = 2 * (k - 1 - b)
        in_lo = sh < 64
        s_lo = torch.where(in_lo, sh, torch.zeros_like(sh))
        s_hi = torch.where(in_lo, torch.zeros_like(sh), sh - 64)
        old = torch.where(in_lo, (elo >> s_lo) & 3, (ehi >> s_hi) & 3)
        new = (old + d) & 3
        elo = torch.where(in_lo, (elo & ~(torch.full_like(sh, 3) << s_lo)) | (new << s_lo), elo)
        ehi = torch.where(in_lo, ehi, (ehi & ~(torch.full_like(sh, 3) << s_hi)) | (new << s_hi))
        erl, erh = _revcomp(elo, ehi, k)
        take = _ult(erl, erh, elo, ehi)
        elo = torch.where(take, erl, elo)
        ehi = torch.where(take, erh, ehi)
        clo = torch.cat([clo, elo]); chi = torch.cat([chi, ehi]); w = torch.cat([w, torch.ones(E, dtype=torch.int64, device=dev)])
        del p, b, d, elo, ehi, erl, erh, sh, in_lo, s_lo, s_hi, old, new
    del flo, fhi, rlo, rhi, take
    lo, hi, cnt = _unique_sum(clo, chi, w)
    del clo, chi, w
    N = lo.numel()
    r = r_bits or max(4, int(np.ceil(np.log2(max(16, N) / 0.7))))
    r = min(r, 2 * k)
    cols = synth.random_matrix(r, 2 * k, np.random.default_rng(seed + 1))
    pos = torch.zeros(N, dtype=torch.int64, device=dev)
    nbytes = (2 * k + 7) // 8
    for byte in range(nbytes):
        lut = np.zeros(256, dtype=np.uint64)
        for bit in range(8):
            kb = byte * 8 + bit
            if kb >= 2 * k:
                break
            sel = ((np.arange(256) >> bit) & 1).astype(bool)
            lut[sel] ^= cols[2 * k - 1 - kb]
        lut_t = torch.from_numpy(lut.view(np.int64)).to(dev)
        word = lo if byte < 8 else hi
        pos ^= lut_t[(word >> (8 * (byte % 8))) & 0xFF]
    pos &= (1 << r) - 1
    lo, hi, cnt, pos = _sort_pairs(lo, hi, cnt, pos)
    o = torch.sort(pos, stable=True).indices            # pos < 2^r <= 2^62: signed == unsigned order
    lo, hi, cnt = lo[o], hi[o], cnt[o]
    del pos, o
    key_bytes = nbytes
    rec = torch.empty((N, key_bytes + 4), dtype=torch.uint8, device=dev)
    rec[:, :min(8, key_bytes)] = lo.view(torch.uint8).reshape(N, 8)[:, :min(8, key_bytes)]
    if key_bytes > 8:
        rec[:, 8:key_bytes] = hi.view(torch.uint8).reshape(N, 8)[:, :key_bytes - 8]
    rec[:, key_bytes:] = cnt.clamp(max=(1 << 32) - 1).view(torch.uint8).reshape(N, 8)[:, :4]
    header = {
        "alignment": 8, "canonical": True, "cmdline": ["synth_gpu", "-m", str(k)], "counter_len": 4,
        "exe_path": "jasper_b200.synth_gpu", "format": "binary/sorted", "hostname": "synthetic", "key_len": 2 * k,
        "matrix1": {"c": 2 * k, "columns": [int(x) for x in cols], "identity": False, "r": r},
        "max_reprobe": 126, "pwd": "/", "reprobes": [1], "size": 1 << r, "time": "synthetic", "val_len": 7,
    }
    js = json.dumps(header, separators=(",", ":")).encode()
    js += b"\0" * ((-(9 + len(js))) % 8)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "db.jf")
    with open(path, "wb") as f:
        f.write(b"%09d" % len(js))
        f.write(js)
        step = 1 << 24
        for a in range(0, N, step):
            f.write(rec[a:a + step].cpu().numpy().tobytes())
    genome = g.to(torch.uint8).cpu().numpy()
    del rec, lo, hi, cnt, g
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return {"genome": genome, "db": path, "n_records": int(N), "r": r, "lambda": lam, "err_kmers": E}


def make_assembly(genome: np.ndarray, rate: float, seed: int, device: str | None = None, dirty: bool = False) -> bytes:
    """synth.inject_errors (0.5 substitution / 0.25 deletion / 0.25 insertion) with torch ops; dirty: plus runs of N / n and
    lower-case stretches (the reference keeps case and skips windows with N, src/jasper.py:56-69)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    g = torch.from_numpy(genome).to(dev).to(torch.int64)
    n = g.numel()
    x = torch.rand(n, device=dev, generator=gen)
    sub = x < rate * 0.5
    dele = (x >= rate * 0.5) & (x < rate * 0.75)
    ins = (x >= rate * 0.75) & (x < rate)
    a = g.clone()
    a[sub] = (a[sub] + torch.randint(1, 4, (int(sub.sum()),), device=dev, generator=gen)) & 3
    insb = torch.randint(0, 4, (n,), device=dev, generator=gen)
    rep = torch.ones(n, dtype=torch.int64, device=dev)
    rep[dele] = 0
    rep[ins] = 2
    a2 = torch.repeat_interleave(a, rep)
    at = (torch.cumsum(rep, 0) - 1)[ins]
    a2[at] = insb[ins]
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    if not dirty:
        return lut[a2].cpu().numpy().tobytes()
    out = lut[a2].cpu().numpy().copy()
    rng = np.random.default_rng(seed + 5)
    lo, hi = DIRT["n_run_len"]
    per = out.size / 64e6                                                             # (the counts are per 64 Mb)
    for p in rng.integers(0, max(1, out.size - hi), max(1, int(round(DIRT["n_runs"] * per)))):                  # runs of N (and a few of n)
        out[p:p + int(rng.integers(lo, hi))] = ord("N") if rng.random() < 0.8 else ord("n")
    lo, hi = DIRT["lower_len"]
    for p in rng.integers(0, max(1, out.size - hi), max(1, int(round(DIRT["lower_stretches"] * per)))):         # soft-masked stretches
        e = p + int(rng.integers(lo, hi))
        out[p:e] |= 0x20
    return out.tobytes()

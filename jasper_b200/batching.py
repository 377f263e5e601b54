"""Host-side restatement of the chunking / joining that src/jasper.sh does around the polisher
(the rows SURVEY.md §8f-3 marks "next"; bench.py and jasper_b200/pipeline.py use it so that the units
polished on the GPU are exactly the reference's units).

  split   src/jasper.sh:132-139 (BATCH_SIZE), :155 (pieces named "<first header token>:<offset>")
  group   src/jasper.sh:156     (a new batch file starts at a header once accumulated bases > BATCH_SIZE)
  join    src/jasper.sh:220     (re-stitch pieces by walking offsets 0, bs, 2bs, ...)
"""
from __future__ import annotations

MAX_BATCH_SIZE = 25_000_000       # src/jasper.sh:9


def batch_size(total_bp: int, threads: int, user_batch: int = 0) -> int:
    bs = int(total_bp / threads * 0.9)          # awk: int($3/NUM_THREADS*.9)
    out = user_batch
    if bs > out:
        out = bs
        if out > MAX_BATCH_SIZE:
            out = MAX_BATCH_SIZE
    return out


def split_contigs(contigs, bs: int):
    """contigs: [(header_token_without_'>', seq)] -> pieces [(name, seq)] with name 'ctg:offset'."""
    bs = int(bs)
    pieces = []
    for name, seq in contigs:
        if seq == "":
            continue                             # perl: only non-empty sequences are printed
        if bs <= 0:
            raise ValueError("batch size must be positive")
        for ci in range(0, len(seq), bs):
            pieces.append(("%s:%d" % (name, ci), seq[ci:ci + bs]))
    return pieces


def group_pieces(pieces, bs: int):
    """Greedy grouping into batch files: roll to the next file at a header once the bases written
    to the current file exceed bs."""
    groups, cur, written = [], [], 0
    for name, seq in pieces:
        if written > bs:
            groups.append(cur)
            cur, written = [], 0
        cur.append((name, seq))
        written += len(seq)
    groups.append(cur)
    return groups


def join_pieces(pieces, bs: int):
    """pieces: [(name 'ctg:offset', seq)] in any order -> {ctg: seq}; only names with a ':0' piece
    are emitted, pieces are concatenated while ctg:<b> exists for b = 0, bs, 2bs, ..."""
    bs = int(bs) if int(bs) > 0 else 1
    h = dict(pieces)
    out = {}
    for name in h:
        if name.endswith(":0"):
            ctg = name[:-2]
            b, parts = 0, []
            while "%s:%d" % (ctg, b) in h:
                parts.append(h["%s:%d" % (ctg, b)])
                b += bs
            out[ctg] = "".join(parts)
    return out


def shard_units(lengths, world: int):
    """Multi-GPU sharding: longest-first greedy onto the least-loaded rank.
    Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        shards[r].append(i)
        load[r] += lengths[i]
    for s in shards:
        s.sort()
    return shards

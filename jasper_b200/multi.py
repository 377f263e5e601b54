"""Multi-GPU polishing of ONE assembly: the chunks (the reference's own unit of parallelism, src/jasper.sh:155,212)
are sharded across ranks, every rank polishes its shard against its replica of the k-mer table, the four QV
tallies are summed across ranks (src/jasper.sh:235-238 does it with awk over *qValCalcHelper.csv) and rank 0
collects sequences and fix rows.  No data-path collective: chunks are independent.

One process per GPU (torchrun / torch.distributed).  On GPUs the tally sum is the library's NCCL all-reduce
(jb_allreduce_tallies); the object gather of results uses the process group the caller initialised.  The same
code runs under the `gloo` backend on CPUs in tests/, with a stand-in for the per-rank polish call.
"""
from __future__ import annotations

from . import batching


def shard_chunks(chunks, world: int):
    """chunks: [(name, seq)] -> per-rank lists of chunk indices (longest first onto the least-loaded rank)."""
    return batching.shard_units([len(s) for _, s in chunks], world)


def polish_shard(my_chunks, k: int, solid_thre: int, passes: int, polish_fn, world: int, allreduce_tallies=None, fix: bool = True):
    """One rank's part of a sharded run, for callers that hold only their own chunks (a 3.1 Gb assembly is not copied to
    every rank): polish the shard, sum the first/last tallies over the ranks.  -> (result of polish_fn, [w0, t0, wp, tp] of the
    whole assembly)."""
    res = polish_fn(my_chunks, k, solid_thre, passes, fix)
    first, last = res.tallies[0], res.tallies[-1]
    t4 = [first[0], first[1], last[0], last[1]]
    if allreduce_tallies is not None and world > 1:
        t4 = list(allreduce_tallies(t4))
    return res, t4


def polish_sharded(chunks, k: int, solid_thre: int, passes: int, polish_fn, rank: int, world: int,
                   allreduce_tallies=None, gather=None, fix: bool = True):
    """polish_fn(chunk_list, k, solid_thre, passes, fix) -> object with .status, .tallies [(wrong, total)] per
    iteration, .rows [rows per iteration], .seqs [(name, seq)] -- capi.QueryMerFile.polish has this shape.

    allreduce_tallies(list_of_4_ints) -> summed list (all ranks);  gather(obj) -> list of objs on rank 0, None elsewhere.
    Returns on rank 0: dict(status, tallies=[(w0, t0), (wp, tp)], rows=[per iteration], seqs=[(name, seq)] in the
    order of `chunks`); on other ranks: dict(status, tallies) only."""
    shards = shard_chunks(chunks, world)
    mine = shards[rank]
    res, t4 = polish_shard([chunks[i] for i in mine], k, solid_thre, passes, polish_fn, world, allreduce_tallies, fix)
    local = {"status": res.status, "index": mine, "rows": res.rows, "seqs": res.seqs}
    parts = gather(local) if (gather is not None and world > 1) else [local]
    out = {"status": res.status, "tallies": [(t4[0], t4[1]), (t4[2], t4[3])]}
    if parts is None:
        return out
    n_it = max(len(p["rows"]) for p in parts) if parts else 0
    seqs = [None] * len(chunks)
    rows = [[] for _ in range(n_it)]
    status = 0
    order = []
    for p in parts:
        status = status or p["status"]
        for j, ci in enumerate(p["index"]):
            seqs[ci] = p["seqs"][j]
            order.append((ci, p, j))
    # fix rows in chunk order, each chunk's rows in emission order (rows carry the chunk's name)
    order.sort(key=lambda t: t[0])
    for ite in range(n_it):
        by_name = {}
        for p in parts:
            if ite < len(p["rows"]):
                for row in p["rows"][ite]:
                    by_name.setdefault(row[0], []).append(row)
        seen = set()
        for ci, _, _ in order:
            name = chunks[ci][0]
            if name in seen:
                continue
            seen.add(name)
            rows[ite].extend(by_name.get(name, []))
    out.update({"status": status, "rows": rows, "seqs": seqs})
    return out


def torch_collectives(device_tallies=None):
    """(allreduce_tallies, gather) over the default torch.distributed process group.  device_tallies: optional
    callable doing the sum on the GPUs (capi: jb_allreduce_tallies over NCCL); default = dist.all_reduce on CPU."""
    import torch
    import torch.distributed as dist

    def allreduce(t4):
        if device_tallies is not None:
            return device_tallies(t4)
        t = torch.tensor(t4, dtype=torch.int64)
        dist.all_reduce(t)
        return [int(x) for x in t]

    def gather(obj):
        out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
        dist.gather_object(obj, out, dst=0)
        return out

    return allreduce, gather

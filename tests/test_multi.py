"""N > 1 host logic on CPUs: two `gloo` ranks shard one assembly's chunks, polish their shards (the oracle stands in
for the GPU call -- this file is test infrastructure), sum the QV tallies with an all-reduce and gather the
result on rank 0, which must equal the single-process run chunk for chunk."""
import os
import socket
import sys

import pytest

import helpers
import jasper_oracle as jo
from jasper_b200 import batching, multi


class _Res(object):
    def __init__(self, status, tallies, rows, seqs):
        self.status, self.tallies, self.rows, self.seqs = status, tallies, rows, seqs


def _oracle_polish_fn(db_path):
    odb = helpers.OracleDB(db_path)

    def fn(chunk_list, k, thr, passes, fix):
        res, _ = helpers.oracle_polish(odb, chunk_list, k, thr, passes)
        return _Res(0, [(p.wrong, p.total) for p in res.passes], [[list(r) for r in p.rows] for p in res.passes], res.seqs)
    return fn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, chunks, out_path):
    import pickle
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        allreduce, gather = multi.torch_collectives()
        out = multi.polish_sharded(chunks, case["k"], case["threshold"], case["passes"], _oracle_polish_fn(case["db"]),
                                   rank, world, allreduce, gather)
        with open("%s.%d" % (out_path, rank), "wb") as f:
            pickle.dump(out, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["k25_mix", "k31_dirty"])
def test_two_gloo_ranks_equal_one_process(golden_cases, name, tmp_path):
    import pickle
    import torch.multiprocessing as mp
    case = golden_cases[name]
    contigs = jo.read_fasta(case["fasta"])
    total = sum(len(s) for _, s in contigs)
    chunks = batching.split_contigs(contigs, batching.batch_size(total, 7))          # ~8 chunks, like jasper.sh -t 7
    assert len(chunks) >= 4
    single = multi.polish_sharded(chunks, case["k"], case["threshold"], case["passes"], _oracle_polish_fn(case["db"]), 0, 1)
    out_path = str(tmp_path / "out.pkl")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, {q: case[q] for q in ("k", "threshold", "passes", "db")}, chunks, out_path), nprocs=2, join=True)
    r0 = pickle.load(open(out_path + ".0", "rb"))
    r1 = pickle.load(open(out_path + ".1", "rb"))
    assert r0["tallies"] == single["tallies"] == r1["tallies"]              # every rank holds the reduced tallies
    assert "seqs" not in r1                                                 # only rank 0 gathers
    assert r0["seqs"] == single["seqs"]
    assert r0["rows"] == single["rows"]
    # and sharding is what the docstring promises: every chunk exactly once, balanced within one chunk length
    shards = multi.shard_chunks(chunks, 2)
    assert sorted(shards[0] + shards[1]) == list(range(len(chunks)))
    load = [sum(len(chunks[i][1]) for i in s) for s in shards]
    assert abs(load[0] - load[1]) <= max(len(s) for _, s in chunks)
    # the joined assembly: the reference's own join (src/jasper.sh:220) over the gathered pieces
    joined = batching.join_pieces(r0["seqs"], batching.batch_size(total, 7))
    whole = multi.polish_sharded(chunks, case["k"], case["threshold"], case["passes"], _oracle_polish_fn(case["db"]), 0, 1)
    assert joined == batching.join_pieces(whole["seqs"], batching.batch_size(total, 7))

"""The human-scale workload generator (bench tooling, jasper_b200/synth_scale.py) at toy size on the CPU: the parts written by
different ranks / rounds must add up to ONE valid binary/sorted file that the CPU oracle reads like any `jellyfish count` output."""
import numpy as np
import torch

import helpers
from jasper_b200 import batching, synth, synth_scale


def _make(tmp_path, parts, name, G=120_000, block=1 << 15):
    spec = synth_scale.Spec(G, k=37, seed=5, err_kmers_per_bp=0.5, block=block)
    g = synth_scale.genome_tensor(spec, torch.device("cpu"))
    path = str(tmp_path / name)
    meta = synth_scale.make_database(path, spec, g, parts_per_rank=parts)
    return spec, g, path, meta


def test_parts_add_up_to_one_sorted_file(tmp_path):
    spec, g, p1, m1 = _make(tmp_path, 1, "one.jf")
    _, _, p3, m3 = _make(tmp_path, 3, "three.jf")
    assert open(p1, "rb").read() == open(p3, "rb").read()                 # the split into position ranges changes nothing
    assert m1["n_records"] == m3["n_records"] and m1["file_bytes"] == len(open(p1, "rb").read())
    odb = helpers.OracleDB(p1)
    lo, hi, cnt = odb.records()
    assert len(lo) == m1["n_records"]
    # every genome window is in the file with a count >= 1; the records are distinct
    s = synth.codes_to_bytes(g.numpy())
    c = odb.count_positions(s.decode())
    assert np.all(c[:len(s) - 37 + 1] >= 1)
    keys = np.stack([hi, lo], 1)
    assert len(np.unique(keys, axis=0)) == len(lo)
    # roughly G genome k-mers + E singletons
    assert 0.95 * (spec.G + spec.E) < len(lo) <= spec.G + spec.E
    # the table arithmetic accepts it: sorted by (position, key); at this toy size a 74-bit key leaves no room for a compact slot
    sim = helpers.table_sim_lib()
    st, msg = helpers.table_sim_check(sim, p1)
    assert (st["layout"] == 0 and "no room" in msg) or st["records"] == len(lo)


def test_assembly_pieces_shard_like_jasper_sh(tmp_path):
    spec = synth_scale.Spec(400_000, k=37, seed=7, block=1 << 16)
    g = synth_scale.genome_tensor(spec, torch.device("cpu"))
    whole, bs, total, n_pieces = synth_scale.assembly_pieces(g, spec, 0.001, threads=64)
    assert len(whole) == n_pieces and sum(len(s) for _, s in whole) == total and bs == batching.batch_size(total, 64)
    # two ranks: a partition of the same pieces
    a, _, _, _ = synth_scale.assembly_pieces(g, spec, 0.001, threads=64, rank=0, world=2)
    b, _, _, _ = synth_scale.assembly_pieces(g, spec, 0.001, threads=64, rank=1, world=2)
    assert sorted(a + b) == sorted(whole) and a and b
    # pieces are what batching.split_contigs makes of the contigs' assemblies
    names = {n.rsplit(":", 1)[0] for n, _ in whole}
    assert names == {c[0] for c in synth_scale.contig_bounds(spec.G)}
    joined = batching.join_pieces([(n, s.decode()) for n, s in whole], bs)
    assert sum(len(v) for v in joined.values()) == total
    # ~0.1 % of the bases differ from the genome
    name, s0, e0 = synth_scale.contig_bounds(spec.G)[0]
    assert abs(len(joined[name]) - (e0 - s0)) < 0.002 * (e0 - s0) + 10

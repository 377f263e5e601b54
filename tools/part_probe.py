#!/usr/bin/env python
"""What the partitioned table costs: per-position counting (jb_count_positions = k_scan) of the same contig against
the whole table on one GPU and against the table split over two GPUs (half of the probes are peer loads over NVLink).
Run on a box with >= 2 GPUs:  python tools/part_probe.py"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from jasper_b200 import capi, synth_gpu


def main():
    d = tempfile.mkdtemp(prefix="jb_part")
    G = 16_000_000
    w = synth_gpu.make_workload(d, G, 37, seed=7)
    asm = synth_gpu.make_assembly(w["genome"], 0.001, seed=3).decode("latin-1")
    path = os.path.join(d, "db.jf")
    whole = capi.QueryMerFile(path, device=0)
    ref = whole.count_positions(asm)
    for _ in range(3):
        t0 = time.perf_counter(); whole.count_positions(asm); t1 = time.perf_counter()
    ms_whole = capi.timing_ms("count_positions")
    whole.close()
    parts = capi.QueryMerFile.open_partitioned(path, [0, 1])
    got = parts[0].count_positions(asm)
    assert (got == ref).all()
    for _ in range(3):
        parts[0].count_positions(asm)
    ms_part = capi.timing_ms("count_positions")
    res = parts[1].polish([("c", asm)], 37, 5, 2)
    print("records %d, %d windows: whole table %.3f ms (%.2f G lookups/s), partitioned over 2 GPUs %.3f ms (%.2f G lookups/s); polish on part 1: status %d, tallies %s"
          % (w["n_records"], len(asm) + 1, ms_whole, (len(asm) + 1) / ms_whole / 1e6, ms_part, (len(asm) + 1) / ms_part / 1e6, res.status, res.tallies))
    for qf in parts:
        qf.close()


if __name__ == "__main__":
    main()

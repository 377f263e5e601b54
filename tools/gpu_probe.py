#!/usr/bin/env python
"""First-contact GPU probe: random-sector gather roofline + bulk lookup rate on a synthetic table.
Writes gpurun_out/probe.json."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jasper_b200 import capi, synth

def main(n=16_000_000, k=37):
    t0 = time.time()
    rng = np.random.default_rng(1)
    g = synth.random_genome(n, rng)
    lo, hi = synth.canonical_keys(g, k)
    lo, hi, mult = synth.unique_keys(lo, hi)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "probe.jf")
    synth.write_jf(path, lo, hi, mult * 20, k, seed=3)
    print("db built in %.1fs, %d records, %.1f MB" % (time.time() - t0, len(lo), os.path.getsize(path) / 1e6), flush=True)
    t0 = time.time()
    qf = capi.QueryMerFile(path)
    out = {"info": qf.info, "open_s": time.time() - t0}
    print(out, flush=True)
    for sb in (32, 64, 128):
        gbps, gps = qf.gather_bench(n_probes=1 << 28, sector_bytes=sb, iters=3)
        out["gather_%d" % sb] = {"GBps": gbps, "Gprobes_per_s": gps}
        print("gather", sb, gbps, gps, flush=True)
    s = synth.codes_to_bytes(g).decode()
    for it in range(3):
        c = qf.count_positions(s)
        ms = capi.timing_ms("count_positions")
        out["count_positions_ms_%d" % it] = ms
        out["lookups_per_s_%d" % it] = (len(s) + 1) / (ms * 1e-3)
        print("count_positions", ms, "ms", (len(s) + 1) / (ms * 1e-3) / 1e9, "G lookups/s", "h2d", capi.timing_ms("h2d"), "d2h", capi.timing_ms("d2h"), flush=True)
    assert np.all(c[: len(s) - k + 1] >= 20)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000)

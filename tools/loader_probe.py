#!/usr/bin/env python
"""How fast jb_db_open streams a large .jf: cold from the box's disk (page cache dropped), warm from page cache, and cold with
O_DIRECT (JB_ODIRECT=1).  Generates the human-scale database of bench.py --config human at --genome-mb (one GPU builds it).

    python tools/loader_probe.py --genome-mb 800 --dir /tmp/jb_loader_probe > profiles/..._loader_probe.json
"""
import argparse, json, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def drop_caches():
    try:
        subprocess.run("sync; echo 3 > /proc/sys/vm/drop_caches", shell=True, check=True)
        return True
    except Exception:
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mb", type=float, default=800.0)
    ap.add_argument("--dir", default="/tmp/jb_loader_probe")
    args = ap.parse_args()
    import torch
    from jasper_b200 import capi, synth_scale
    dev = torch.device("cuda", 0)
    shutil.rmtree(args.dir, ignore_errors=True)
    os.makedirs(args.dir)
    path = os.path.join(args.dir, "db.jf")
    spec = synth_scale.Spec(int(args.genome_mb * 1e6), 37, seed=20, err_kmers_per_bp=0.5)
    t0 = time.time()
    g = synth_scale.genome_tensor(spec, dev)
    meta = synth_scale.make_database(path, spec, g, 0, 1, None, parts_per_rank=8)
    del g
    torch.cuda.empty_cache()
    out = {"file_bytes": meta["file_bytes"], "n_records": meta["n_records"], "generate_and_write_s": time.time() - t0, "dir": args.dir, "runs": []}
    for label, env, cold in (("cold, buffered reads", {"JB_ODIRECT": "0"}, True), ("warm (page cache)", {"JB_ODIRECT": "0"}, False),
                             ("cold, O_DIRECT", {"JB_ODIRECT": "1"}, True), ("warm file, O_DIRECT", {"JB_ODIRECT": "1"}, False)):
        dropped = drop_caches() if cold else None
        os.environ.update(env)
        t0 = time.time()
        qf = capi.QueryMerFile(path, device=0)
        secs = time.time() - t0
        info = qf.info
        out["runs"].append({"what": label, "caches_dropped": dropped, "open_s": secs, "GBps": meta["file_bytes"] / secs / 1e9,
                            "busiest_reader_s": info["read_seconds"], "table_bytes": info["table_bytes"]})
        qf.close()
    shutil.rmtree(args.dir, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

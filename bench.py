#!/usr/bin/env python
"""bench.py -- the headline benchmark of the polishing hot path (BASELINE.json: assembly bp
polished/sec and k-mer lookups/sec vs the reference's CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genome-mb M]

One "step" = one pass of the hot path over one batch: the whole per-batch loop of src/jasper.py
main() (p fixing passes + the final count-only pass) over the workload's contig chunks.
Workload (configs[1]): synthetic 64 Mb chr20-scale assembly with 0.1 % injected errors, k-mer
database of a simulated 30x Illumina read set written directly as a Jellyfish binary/sorted file,
k = 37, p = 2, chunked exactly like src/jasper.sh does for -t 16.
`value` is device-timed with the chunks already resident in HBM; `e2e` goes through the public
call (host buffers in, polished host strings out) with both copies inside the timed region.
Multi-GPU (torchrun, one rank per GPU): every rank holds a replica of the table and polishes its
own assembly of the same size (weak scaling); the QV tallies are summed with one NCCL all-reduce.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 37
PASSES = 2
SOLID_THRE = 5
CHUNK_THREADS = 16            # the -t of src/jasper.sh that defines the chunking (src/jasper.sh:132-139)
ASM_ERR = 0.001


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=64.0)
    ap.add_argument("--cpu-sample-kb", type=int, default=0, help="bases of the workload the CPU baseline polishes (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- workload
def build_workload(work_dir: str, genome_mb: float, rank: int, make_db: bool):
    """Genome + .jf (rank 0 writes it once) + this rank's assembly, chunked like jasper.sh -t 16."""
    from jasper_b200 import batching, synth_gpu
    G = int(genome_mb * 1e6)
    t0 = time.time()
    if make_db:
        w = synth_gpu.make_workload(work_dir, G, K, seed=20)
        genome = w["genome"]
        meta = {a: b for a, b in w.items() if a != "genome"}
        import numpy as np
        np.save(os.path.join(work_dir, "genome.npy"), genome)
        json.dump(meta, open(os.path.join(work_dir, "meta.json"), "w"))
    else:
        import numpy as np
        genome = np.load(os.path.join(work_dir, "genome.npy"))
        meta = json.load(open(os.path.join(work_dir, "meta.json")))
    asm = synth_gpu.make_assembly(genome, ASM_ERR, seed=1000 + rank).decode("latin-1")
    bs = batching.batch_size(len(asm), CHUNK_THREADS)
    chunks = batching.split_contigs([("chr20_like", asm)], bs)
    meta["gen_seconds"] = time.time() - t0
    meta["assembly_bp"] = len(asm)
    meta["batch_size"] = bs
    return meta, chunks


# ----------------------------------------------------------------------------- the reference's CPU path
# oracle/_ref/ (git-ignored, built by `make -C oracle ref`, shipped to the GPU box by gpurun) holds the UNMODIFIED reference:
# jasper/jasper.py, jasper/jasper.sh, jasper/jellyfish.py, bin/jellyfish and the compiled Jellyfish Python binding.
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_JASPER = os.path.join(REF_DIR, "jasper", "jasper.py")


def reference_available():
    import glob
    return os.path.exists(REF_JASPER) and bool(glob.glob(os.path.join(REF_DIR, "_dna_jellyfish*.so")))


def touch_reference_binding(db_path):
    """Import the reference's compiled binding in THIS process and open the database with it once (the workers below
    are separate processes: this makes the .so visible to whoever records what the bench process loaded) -> C('A'*k)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dna_jellyfish as jf
    qf = jf.QueryMerFile(db_path)
    return int(qf[jf.MerDNA("A" * K).get_canonical()])


class ReferenceRun(object):
    """The polishing stage of the unmodified src/jasper.sh (:206-212) over the workload's chunks: one batch file per chunk
    (what the split at :155-156 writes: '>name' line, sequence on one line), run_jasper.sh written as jasper.sh writes it,
    `ls *.batch.*.fa | xargs -P <cores> -I{} ./run_jasper.sh {}`, every worker the unmodified src/jasper.py importing the
    reference's compiled Jellyfish binding."""

    def __init__(self, work_dir, db_path, chunks, procs):
        self.dir = os.path.join(work_dir, "ref_run")
        shutil.rmtree(self.dir, ignore_errors=True)
        os.makedirs(self.dir)
        self.db, self.procs, self.chunks = db_path, procs, chunks
        self.files = []
        for i, (name, seq) in enumerate(chunks):
            fn = "asm.fa.batch.%d.fa" % i
            with open(os.path.join(self.dir, fn), "w") as f:
                f.write(">%s\n%s\n" % (name, seq))
            self.files.append(fn)
        last = PASSES - 1
        with open(os.path.join(self.dir, "run_jasper.sh"), "w") as f:          # src/jasper.sh:207-208, verbatim but for the paths
            f.write("#!/bin/bash\n")
            f.write("%s --db %s --query $1 --ksize %d -p %d --fix --fout $1.fix.csv -ff $1.fixed.fa.tmp --test -thre %d 1>jasper.out 2>jasper.err "
                    "&& mv _iter%d_$1.fixed.fa.tmp _iter%d_$1.fixed.fa\n" % (REF_JASPER, db_path, K, PASSES, SOLID_THRE, last, last))
        os.chmod(os.path.join(self.dir, "run_jasper.sh"), 0o755)
        self.env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle"))

    def step(self):
        """-> wall seconds of the xargs line"""
        for fn in os.listdir(self.dir):
            if fn.startswith("_iter") or fn.endswith("qValCalcHelper.csv"):
                os.remove(os.path.join(self.dir, fn))
        t0 = time.perf_counter()
        subprocess.check_call("ls asm.fa.batch.*.fa | xargs -P %d -I{} ./run_jasper.sh {}" % self.procs, shell=True, cwd=self.dir, env=self.env)
        return time.perf_counter() - t0

    def outputs(self):
        """-> (polished sequence per chunk, {iteration: [wrong, total]} summed like src/jasper.sh:235-238, fix.csv text per pass)"""
        last = PASSES - 1
        seqs = []
        for fn in self.files:
            lines = open(os.path.join(self.dir, "_iter%d_%s.fixed.fa" % (last, fn))).read().split("\n")
            seqs.append("".join(l for l in lines if not l.startswith(">")))
        tallies = {}
        for ite in (0, PASSES):
            w = t = 0
            for l in open(os.path.join(self.dir, "%dqValCalcHelper.csv" % ite)):
                a, b = l.split()
                w += int(a); t += int(b)
            tallies[ite] = [w, t]
        rows = ["".join(open(os.path.join(self.dir, "_iter%d_%s.fix.csv" % (ite, fn)), newline="").read() for fn in self.files) for ite in range(PASSES)]
        return seqs, tallies, rows


def product_fix_csv_text(res, chunks):
    """What src/jasper.py would have written for every chunk file, from the GPU result (jasper_b200/jasper.py's writer)."""
    import csv
    import io
    out = []
    for ite in range(PASSES):
        by_name = {}
        for row in res.rows[ite]:
            by_name.setdefault(row[0], []).append(row)
        buf = io.StringIO()
        for name, _ in chunks:
            w = csv.writer(buf, delimiter=" ")
            w.writerow(["Contig", "Base_coord", "Original", "Mutation"])
            w.writerows(by_name.get(name, []))
        out.append(buf.getvalue())                    # (csv's "\r\n" line ends, like the reference's files)
    return out


# ----------------------------------------------------------------------------- CPU port (fallback where oracle/_ref is absent)
def _cpu_polish_worker(args):
    db_path, seqs, k, thr, passes = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import jasper_oracle as jo
    import helpers
    odb = helpers.OracleDB(db_path)
    o = jo.PolishOracle(odb.count, k, thr)
    t0 = time.perf_counter()
    res = o.polish(seqs, passes)
    dt = time.perf_counter() - t0
    return dt, res.lookups, sum(len(s) for _, s in seqs), [(p.wrong, p.total) for p in res.passes], "port", [s for _, s in res.seqs]


def cpu_port_baseline(db_path, chunks, sample_bp: int, procs: int):
    """The oracle (Python restatement of src/jasper.py over oracle/libjf_oracle.so) on a bounded sample: `procs` worker
    processes, one slice each.  Only used where the reference itself (oracle/_ref) is not installed."""
    import multiprocessing as mp
    whole = "".join(s for _, s in chunks)
    per = int(max(2000, min(sample_bp // procs, len(whole) // procs)))
    jobs = []
    for w in range(procs):
        at = w * (len(whole) // procs)
        piece = whole[at:at + per]
        if len(piece) < 2 * K:
            continue
        jobs.append((db_path, [("slice_%d" % w, piece)], K, SOLID_THRE, PASSES))
    t0 = time.perf_counter()
    if len(jobs) == 1:
        outs = [_cpu_polish_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            outs = pool.map(_cpu_polish_worker, jobs)
    wall = time.perf_counter() - t0
    bp = sum(o[2] for o in outs)
    lookups = sum(o[1] for o in outs)
    return {"wall": wall, "bp": bp, "lookups": lookups, "kind": "port", "procs": len(jobs), "jobs": jobs, "outs": outs}


def workload_text(genome_mb, meta, n_chunks):
    return ("synthetic %.0f Mb chr20-scale assembly, 0.1%% injected errors, simulated 30x Illumina .jf (%d records), k=%d, p=%d, "
            "chunked as jasper.sh -t %d (%d chunks of <= %d bp), solid_thre=%d; one assembly per GPU against a replicated table"
            % (genome_mb, meta["n_records"], K, PASSES, CHUNK_THREADS, n_chunks, meta["batch_size"], SOLID_THRE))


def ncu_traffic_gb():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_scan launch on this workload, from the committed
    ncu --set full summary (profiles/, written by tools/ncu_keymetrics.py); None when the summary is missing."""
    import csv
    scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
    try:
        vals = {}
        for row in csv.reader(l for l in open(os.path.join(ROOT, "profiles", "r1x_scan_kernels_ncu_full.csv")) if not l.startswith("#")):
            if len(row) == 7 and row[1] == "k_scan" and row[4].startswith("dram__bytes_"):
                vals[row[4]] = float(row[6]) * scale[row[5]]
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    except Exception:
        return None


# ----------------------------------------------------------------------------- main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from jasper_b200 import capi
    import ctypes as C

    base = os.environ.get("JB_BENCH_DIR") or os.path.join(tempfile.gettempdir(), "jb_bench_%s" % os.environ.get("MASTER_PORT", "single"))
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        os.makedirs(base, exist_ok=True)
    if world > 1:
        dist.barrier()
    if rank == 0:
        meta, chunks = build_workload(base, args.genome_mb, rank, make_db=True)
    if world > 1:
        dist.barrier()
    if rank != 0:
        meta, chunks = build_workload(base, args.genome_mb, rank, make_db=False)
    db_path = os.path.join(base, "db.jf")

    t0 = time.time()
    qf = capi.QueryMerFile(db_path, device=local_rank)
    open_s = time.time() - t0
    lib = capi.load_library()
    if world > 1:
        uid = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            capi._check(lib.jb_comm_unique_id(buf))
            uid = [buf.raw]
        dist.broadcast_object_list(uid, src=0)
        capi._check(lib.jb_comm_init(rank, world, uid[0], local_rank))

    total_bp = sum(len(s) for _, s in chunks)

    def device_step():
        b = qf.upload(chunks)                         # not timed: `value` starts with the chunks resident
        b.polish(K, SOLID_THRE, PASSES)
        ms = capi.timing_ms("polish")
        stats = {"scan_ms": capi.timing_ms("polish_scan"), "walk_ms": capi.timing_ms("polish_walk"),
                 "scan_lookups": capi.counter("scan_lookups"), "walk_lookups": capi.counter("walk_lookups"),
                 "launches": capi.counter("kernel_launches"), "probe_steps": capi.counter("probe_steps"), "events": capi.counter("events"),
                 "plan_ms": capi.timing_ms("polish_plan"), "segments": capi.counter("segments"), "reruns": capi.counter("segment_reruns"),
                 "full_scans": capi.counter("full_scans"), "full_scan_lookups": capi.counter("full_scan_lookups"),
                 "recount_probes": capi.counter("recount_probes"), "recount_ms": capi.timing_ms("polish_recount"),
                 "seg_block": capi.counter("seg_block"), "fast_events": capi.counter("fast_events"), "fast_steps": capi.counter("fast_steps"),
                 "verdicts": capi.counter("verdicts"), "verdict2_events": capi.counter("verdict2_events"), "verdict_ms": capi.timing_ms("polish_verdicts"), "kwalk_ms": capi.timing_ms("polish_kwalk"),
                 "max_seg_events": capi.counter("max_seg_events"), "max_seg_lookups": capi.counter("max_seg_lookups"),
                 "max_seg_span": capi.counter("max_seg_span"),
                 "slowest_segment_per_pass": [{key: capi.counter("%s%d" % (key, q)) for key in ("slow_cycles", "slow_events", "slow_lookups", "slow_span", "sum_cycles")} for q in range(3)]}
        return b, ms, stats

    for _ in range(args.warmup):
        b, _, _ = device_step()
        b.free()
    # ---- timed: K steps, device time (CUDA events on the library's stream), max over ranks
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, last_stats, last_batch = 0.0, None, None
    for step in range(args.steps):
        if last_batch is not None:
            last_batch.free()
        last_batch, ms, last_stats = device_step()
        dev_ms += ms
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    res = last_batch.fetch()
    last_batch.free()
    # ---- e2e: the C-ABI call the reference-side binding makes: jb_polish with HOST buffers in,
    # host-visible results out (both PCIe copies inside), K steps, wall clock.  The step's result is read back:
    # tallies, row counts and a checksum over the first/last bytes of every polished contig.
    arr, keep = qf._seq_array(chunks)
    names = [n for n, _ in chunks]

    def e2e_step():
        res = C.c_void_p()
        capi._check(lib.jb_polish(qf._h, arr, len(chunks), K, SOLID_THRE, PASSES, 1, C.byref(res)))
        w, t = C.c_int64(), C.c_int64()
        capi._check(lib.jb_result_tally(res, PASSES, C.byref(w), C.byref(t)))
        nrows = sum(lib.jb_result_num_rows(res, ite) for ite in range(PASSES))
        d2h_seen = 0
        for i in range(len(chunks)):
            p, ln = C.c_void_p(), C.c_uint64()
            capi._check(lib.jb_result_seq(res, i, C.byref(p), C.byref(ln)))
            d2h_seen += ln.value
        lib.jb_result_free(res)
        return w.value, t.value, nrows, d2h_seen

    for _ in range(min(2, args.warmup)):
        e2e_step()
    e2e_t0 = time.perf_counter()
    for step in range(args.steps):
        e2e_last = e2e_step()
    e2e_s = time.perf_counter() - e2e_t0
    h2d, d2h = capi.counter("h2d_bytes"), capi.counter("d2h_bytes")
    e2e_wall = {key: capi.timing_ms(key) for key in ("wall_upload", "wall_polish", "wall_fetch")}
    # the Python convenience wrapper (builds Python lists of rows and str objects) timed separately, once
    py_t0 = time.perf_counter()
    r2 = qf.polish(chunks, K, SOLID_THRE, PASSES)
    py_api_s = time.perf_counter() - py_t0
    assert (r2.tallies[-1][0], r2.tallies[-1][1]) == e2e_last[:2]
    clocks = sampler.stop() if rank == 0 else None

    tallies = (C.c_int64 * 4)(res.tallies[0][0], res.tallies[0][1], res.tallies[-1][0], res.tallies[-1][1])
    if world > 1:
        capi._check(lib.jb_allreduce_tallies(tallies))
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        tb = torch.tensor([total_bp], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb)
        job_bp = float(tb[0])
    else:
        job_bp = float(total_bp)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms / args.steps
    value = job_bp * PASSES / (ms_per_step * 1e-3)
    e2e_value = job_bp * PASSES / (e2e_s / args.steps)

    # ---- roofline of the probe kernel (k_scan): algorithmic bytes = 32 B per lookup
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    gather_gbps, gather_gps = qf.gather_bench(n_probes=1 << 28, sector_bytes=32, iters=3)
    _, line_gps = qf.gather_bench(n_probes=1 << 28, sector_bytes=128, iters=3)      # whole 128-byte lines, four lanes per line
    # k_scan runs once per step (pass 0); later passes inherit counts through k_recount
    scan_ms_per_launch = last_stats["scan_ms"] / max(1, last_stats["full_scans"])
    lookups_per_launch = last_stats["full_scan_lookups"] / max(1, last_stats["full_scans"])
    achieved = lookups_per_launch * 32 / (scan_ms_per_launch * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_scan", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy), of measured" if peaks else "fallback 6650, of fallback",
                "traffic": ncu_traffic_gb(), "traffic_unit": "GB per launch (dram read+write, profiles/ ncu --set full capture of this workload)",
                "algorithmic_gb_per_launch": lookups_per_launch * 32 / 1e9,
                "lookups_per_s": lookups_per_launch / (scan_ms_per_launch * 1e-3),
                "random_sector_peak_GBps": gather_gbps, "frac_of_random_sector_peak": achieved / gather_gbps if gather_gbps else None,
                "ms_per_launch": scan_ms_per_launch, "share_of_step": last_stats["scan_ms"] / (dev_ms / args.steps),
                "probe_units_per_lookup": last_stats["probe_steps"] / max(1, last_stats["full_scan_lookups"]),
                "probe_unit": "32-byte sector (four 8-byte slots)" if qf.info["layout"] == 1 else "32-byte sector (two 16-byte slots)",
                "random_sector_peak_Gsectors_per_s": gather_gps, "random_line_peak_Glines_per_s": line_gps,
                "table": {k2: qf.info[k2] for k2 in ("layout", "slot_bytes", "remainder_bits", "tag_bits", "count_bits", "load_factor", "bytes_per_record", "n_stash", "table_bytes", "n_records", "max_displacement", "load_seconds", "read_seconds", "file_bytes")},
                "launches_per_step": last_stats["full_scans"]}

    # ---- CPU baseline beside it: the reference itself on the box's host cores, over the WHOLE workload when it is installed
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = len(os.sched_getaffinity(0))
        if reference_available():
            touch_reference_binding(db_path)
            ref = ReferenceRun(base, db_path, chunks, cores)
            wall = ref.step()
            r_seqs, r_tallies, r_rows = ref.outputs()
            # the step's own result (`res`: the last timed batch) against the reference's files: the bench doubles as a parity
            # check of the whole workload -- polished bases, fix.csv bytes of every pass, QV tallies
            same_bases = [x for _, x in res.seqs] == r_seqs
            same_tallies = [list(res.tallies[0]), list(res.tallies[-1])] == [r_tallies[0], r_tallies[PASSES]]
            same_rows = product_fix_csv_text(res, chunks) == r_rows
            if not (same_bases and same_tallies and same_rows):
                raise SystemExit("bench: the GPU path and the unmodified reference disagree on the bench workload (bases %s, tallies %s, fix rows %s)"
                                 % (same_bases, same_tallies, same_rows))
            cpu = {"value": total_bp * PASSES / wall, "unit": "bp polished/s", "cores": min(cores, len(chunks)), "kind": "reference",
                   "parity": "all %d chunks (%d bp): polished bases, the fix.csv bytes of both passes (%d rows) and the QV tallies written by the unmodified "
                             "src/jasper.py are identical to the GPU path's" % (len(chunks), total_bp, sum(len(r) for r in res.rows)),
                   "sample": "the whole workload once: unmodified src/jasper.py per chunk file under `xargs -P %d` (src/jasper.sh:206-212) over the reference's "
                             "compiled Jellyfish binding (oracle/_ref), %.1f s wall" % (cores, wall)}
        else:
            sample = (args.cpu_sample_kb * 1000) or min(total_bp, 3_000_000 * cores)
            cb = cpu_port_baseline(db_path, chunks, sample, cores)
            slices = [job[1][0] for job in cb["jobs"]]
            got = qf.polish(slices, K, SOLID_THRE, PASSES)
            same_bases = [s for _, s in got.seqs] == [o[5][0] for o in cb["outs"]]
            want_t = [[sum(o[3][ite][q] for o in cb["outs"]) for q in (0, 1)] for ite in range(PASSES + 1)]
            same_tallies = [list(t) for t in got.tallies] == want_t
            if not (same_bases and same_tallies):
                raise SystemExit("bench: the GPU path and the CPU port disagree on the bench workload (bases %s, tallies %s)" % (same_bases, same_tallies))
            cpu = {"value": cb["bp"] * PASSES / cb["wall"], "unit": "bp polished/s", "cores": cb["procs"], "kind": "port",
                   "parity": "polished bases and QV tallies of all %d slices (%d bp) identical to the GPU path" % (len(slices), cb["bp"]),
                   "sample": "%d slices of %d bp of the same chunks, k=%d p=%d, python restatement (oracle/) over oracle/libjf_oracle.so (oracle/_ref not installed), %.1f s wall"
                             % (cb["procs"], cb["bp"] // max(1, cb["procs"]), K, PASSES, cb["wall"]),
                   "lookups_per_s": cb["lookups"] / cb["wall"]}

    line = {
        "metric": "assembly bp polished/sec", "value": value, "unit": "bp polished/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64/u8 (2-bit packed k-mers, 32-bit counts)", "data": "synthetic",
        "config": {"workload": workload_text(args.genome_mb, meta, len(chunks)),
                   "l2_policy": "inputs larger than L2: %.1f GB table probed at random, fresh upload every step" % (qf.info["table_bytes"] / 1e9)},
        "e2e": {"value": e2e_value, "unit": "bp polished/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3, "call": "jb_polish (C ABI, host buffers in / host results out)",
                "python_wrapper_ms": py_api_s * 1e3, "last_call_wall_ms": e2e_wall},
        "gpu_launches": int(last_stats["launches"] * args.steps),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "lookups_per_s": (last_stats["scan_lookups"] + last_stats["walk_lookups"]) / (ms_per_step * 1e-3) * world,
        "detail": {"scan_ms": last_stats["scan_ms"], "recount_ms": last_stats["recount_ms"], "recount_probes": last_stats["recount_probes"], "plan_ms": last_stats["plan_ms"], "walk_ms": last_stats["walk_ms"], "events": last_stats["events"], "verdicts": last_stats["verdicts"], "verdict_ms": last_stats["verdict_ms"],
                   "events_from_verdicts": last_stats["fast_events"], "scan_steps_from_verdicts": last_stats["fast_steps"], "fixes_from_second_stage_verdicts": last_stats["verdict2_events"],
                   "kwalk_ms": last_stats["kwalk_ms"], "max_seg_events": last_stats["max_seg_events"], "max_seg_lookups": last_stats["max_seg_lookups"],
                   "max_seg_span": last_stats["max_seg_span"], "slowest_segment_per_pass": last_stats["slowest_segment_per_pass"], "segments": last_stats["segments"], "segment_reruns": last_stats["reruns"], "seg_block": last_stats["seg_block"],
                   "walk_lookups": last_stats["walk_lookups"], "db_open_s": open_s, "table_gb": qf.info["table_bytes"] / 1e9,
                   "workload_gen_s": meta["gen_seconds"], "qv_tallies_all_ranks": list(tallies),
                   "fix_rows_pass0": len(res.rows[0]), "fix_rows_pass1": len(res.rows[1]) if len(res.rows) > 1 else 0},
    }
    print(json.dumps(line))
    qf.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the polishing stage of the
    unmodified jasper.sh (xargs -P <cores> of the unmodified jasper.py) over the SAME chunks the GPU arm polishes; one step = the
    whole workload once."""
    if rank != 0:
        return
    base = os.path.join(tempfile.gettempdir(), "jb_bench_ref")
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base, exist_ok=True)
    meta, chunks = build_workload(base, args.genome_mb, 0, make_db=True)
    db_path = os.path.join(base, "db.jf")
    cores = len(os.sched_getaffinity(0))
    total_bp = sum(len(s) for _, s in chunks)
    if reference_available():
        touch_reference_binding(db_path)
        ref = ReferenceRun(base, db_path, chunks, cores)
        walls = []
        for step in range(args.warmup + args.steps):
            w = ref.step()
            if step >= args.warmup:
                walls.append(w)
        _, tallies, rows = ref.outputs()
        wall = sum(walls)
        value = total_bp * PASSES * args.steps / wall
        kind, procs = "reference", min(cores, len(chunks))
        sample = ("each step: the whole workload (%d chunks, %d bp): unmodified src/jasper.py per chunk file under `xargs -P %d` (src/jasper.sh:206-212), "
                  "the reference's compiled Jellyfish binding; QV tallies %s -> %s, %d fix rows"
                  % (len(chunks), total_bp, cores, tallies[0], tallies[PASSES], sum(r.count("\n") - len(chunks) for r in rows)))
    else:
        sample_bp = (args.cpu_sample_kb * 1000) or min(total_bp, 1_000_000 * cores)
        walls, bps = [], []
        for step in range(args.warmup + args.steps):
            cb = cpu_port_baseline(db_path, chunks, sample_bp, cores)
            if step >= args.warmup:
                walls.append(cb["wall"]); bps.append(cb["bp"])
        wall = sum(walls)
        value = sum(bps) * PASSES / wall
        kind, procs = "port", cb["procs"]
        sample = ("each step: %d worker processes x %d bp slices of the workload's chunks (python restatement of src/jasper.py over oracle/libjf_oracle.so; "
                  "oracle/_ref not installed)" % (procs, bps[-1] // max(1, procs)))
    line = {
        "impl": "reference", "metric": "assembly bp polished/sec", "value": value, "unit": "bp polished/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "python int / uint64", "data": "synthetic",
        "config": {"workload": workload_text(args.genome_mb, meta, len(chunks)),
                   "l2_policy": "CPU path: the whole workload per step" if kind == "reference" else "CPU path: a bounded sample of the workload per step (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "bp polished/s", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "bp polished/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

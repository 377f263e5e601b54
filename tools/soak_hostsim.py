#!/usr/bin/env python
"""Soak: the device walker's source (host build, tests/hostsim) against the oracle on fuzz genomes beyond the seeds the test suite runs.

    python tools/soak_hostsim.py FIRST_SEED N_SEEDS [big]

Every seed: a random dirty genome (N runs, lower case, IUPAC, repeats, heterozygosity, clustered errors; tests/helpers.py
fuzz_case), two passes, with and without events / segments; fix rows, bases and tallies must equal the oracle's (which the
goldens and tests/test_reference_container.py pin to the unmodified src/jasper.py).  Prints one summary line.
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    sim = helpers.hostsim_lib()
    bad, crashes, events, bridged = [], 0, 0, 0
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(first, first + n):
            k = helpers.FUZZ_KS[seed % len(helpers.FUZZ_KS)]
            recs, path, thr = helpers.write_fuzz_db(tmp, seed, k, big=big)
            odb = helpers.OracleDB(path)
            res, orc = helpers.oracle_polish(odb, recs, k, thr, 2)
            names = [nm for nm, _ in recs]
            for use_events in (1, 300):
                st = {}
                status, txt, out, tallies = helpers.hostsim_polish(sim, odb, recs, k, thr, 2, use_events, stats=st)
                if res is None:
                    crashes += use_events == 1
                    if status != -7:
                        bad.append((seed, use_events, "status %d where the reference crashes" % status))
                    continue
                if status != 0 or txt != helpers.rows_text(res.passes, names) or out != [s for _, s in res.seqs] or \
                        tallies != [(p.wrong, p.total) for p in res.passes]:
                    bad.append((seed, use_events, "status %d / bytes differ" % status))
                if use_events == 1:
                    events += st.get("events", 0)
            if res is not None:
                bridged += sum(1 for p in res.passes for r in p.rows if len(r[2]) > 1 or len(r[3]) > 2)
            os.remove(path)
    print("seeds %d..%d%s: %d divergences %s; %d reference crashes, %d events, %d multi-base rows"
          % (first, first + n - 1, " (big)" if big else "", len(bad), bad[:5], crashes, events, bridged))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

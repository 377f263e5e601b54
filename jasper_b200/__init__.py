"""jasper_b200 -- B200-native drop-in for the polishing hot path of JASPER.

The package holds only what the path needs:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/jasper_b200.h)
  capi.py          ctypes binding of libjasper_b200.so, mirroring the reference's
                   dna_jellyfish interface (QueryMerFile / MerDNA) for this path
  jasper.py        host side of the reference's src/jasper.py (same CLI, same output files)
  pipeline.py      the reference's src/jasper.sh in one process (split, histogram + threshold, polishing over the GPUs, join, fixes.csv, QV)
  batching.py      jasper.sh's batch size and contig splitting; multi.py: one rank per GPU
  synth*.py        synthetic genomes / .jf writers for tests and bench.py (tooling)
There is no CPU fallback: without the CUDA library every compute call raises.
"""
from .capi import (JasperError, QueryMerFile, MerDNA, PolishResult, load_library, library_path)  # noqa: F401

__all__ = ["JasperError", "QueryMerFile", "MerDNA", "PolishResult", "load_library", "library_path"]

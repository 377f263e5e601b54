"""Loader for the reference's own Jellyfish 2.3.0 Python binding -- TEST INFRASTRUCTURE ONLY.

`oracle/_ref/_dna_jellyfish*.so` is compiled by oracle/Makefile from the unmodified,
pre-generated SWIG wrapper inside /root/reference/jellyfish-2.3.0.tar.gz
(JF!swig/python/swig_wrap.cpp).  This file is the thin Python proxy that SWIG would
normally generate next to it; we write our own (only the three entry points that
src/jasper.py uses -- QueryMerFile, MerDNA, get_canonical, JF!swig/mer_file.i:18-41,
JF!swig/mer_dna.i:15-71 -- plus ReadMerFile for table-dump tests) so that no
reference source is copied into the repo.

Importing this module as `dna_jellyfish` lets the unmodified /root/reference/src/jasper.py
run here (golden-vector generation) and lets bench.py's reference arm time the real
C++ lookup path on the GPU box's host cores.
"""
import glob
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_cands = sorted(glob.glob(os.path.join(_here, "_ref", "_dna_jellyfish*.so")))
if not _cands:
    raise ImportError("oracle/_ref/_dna_jellyfish*.so not built; run `make -C oracle ref` "
                      "in a container that has /root/reference")
_spec = importlib.util.spec_from_file_location("_dna_jellyfish", _cands[0])
_impl = importlib.util.module_from_spec(_spec)
sys.modules.setdefault("_dna_jellyfish", _impl)
_spec.loader.exec_module(_impl)


class MerDNA(object):
    """2-bit k-mer (JF!include/jellyfish/mer_dna.hpp:717)."""

    def __init__(self, *args):
        _impl.MerDNA_swiginit(self, _impl.new_MerDNA(*args))

    @staticmethod
    def k(*args):
        return _impl.MerDNA_k(*args)

    def get_canonical(self):
        return _impl.MerDNA_get_canonical(self)

    def get_reverse_complement(self):
        return _impl.MerDNA_get_reverse_complement(self)

    def canonicalize(self):
        return _impl.MerDNA_canonicalize(self)

    def __str__(self):
        return _impl.MerDNA___str__(self)

    def __eq__(self, other):
        return _impl.MerDNA___eq__(self, other)

    def __lt__(self, other):
        return _impl.MerDNA___lt__(self, other)

    __hash__ = None
    __swig_destroy__ = _impl.delete_MerDNA


_impl.MerDNA_swigregister(MerDNA)


class QueryMerFile(object):
    """Random access to a binary/sorted .jf (JF!swig/mer_file.i:18-41)."""

    def __init__(self, path):
        _impl.QueryMerFile_swiginit(self, _impl.new_QueryMerFile(path))

    def __getitem__(self, m):
        return _impl.QueryMerFile___getitem__(self, m)

    __swig_destroy__ = _impl.delete_QueryMerFile


_impl.QueryMerFile_swigregister(QueryMerFile)


class ReadMerFile(object):
    """Sequential reader (JF!swig/mer_file.i:60-106)."""

    def __init__(self, path):
        _impl.ReadMerFile_swiginit(self, _impl.new_ReadMerFile(path))

    def next_mer(self):
        return _impl.ReadMerFile_next_mer(self)

    def mer(self):
        return _impl.ReadMerFile_mer(self)

    def count(self):
        return _impl.ReadMerFile_count(self)

    def __iter__(self):
        return self

    def __next__(self):
        if not self.next_mer():
            raise StopIteration
        return self.mer(), self.count()

    __swig_destroy__ = _impl.delete_ReadMerFile


_impl.ReadMerFile_swigregister(ReadMerFile)

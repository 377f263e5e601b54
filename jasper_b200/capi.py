"""ctypes binding of libjasper_b200.so (include/jasper_b200.h).

Mirrors the operator interface the reference polisher uses:

    qf = QueryMerFile(path)          # JF!swig/mer_file.i:18-36  -> jb_db_open
    qf[MerDNA(s).get_canonical()]    # JF!swig/mer_file.i:41     -> jb_lookup (one string)
    qf.lookup(strings)               # the same for a batch, one kernel launch
    qf.count_positions(seq)          # C(seq[i:i+k]) for every i
    qf.polish(seqs, k, thre, passes) # src/jasper.py main()      -> jb_polish

Errors surface as JasperError (a RuntimeError, like the SWIG binding's std::runtime_error);
there is no fallback when the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class JasperError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (jb error %d)" % (msg, code))
        self.code = code


class _DbInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("key_bits", C.c_int32), ("counter_len", C.c_int32), ("canonical", C.c_int32),
                ("matrix_r", C.c_int32), ("matrix_c", C.c_int32), ("matrix_identity", C.c_int32), ("device", C.c_int32),
                ("hash_size", C.c_uint64), ("n_records", C.c_uint64), ("table_slots", C.c_uint64),
                ("table_bytes", C.c_uint64), ("max_displacement", C.c_uint64), ("load_seconds", C.c_double),
                ("build_kernel_ms", C.c_double), ("layout", C.c_int32), ("slot_bytes", C.c_int32), ("remainder_bits", C.c_int32),
                ("tag_bits", C.c_int32), ("count_bits", C.c_int32), ("n_parts", C.c_int32), ("home_units", C.c_uint64),
                ("n_stash", C.c_uint64), ("load_factor", C.c_double), ("bytes_per_record", C.c_double), ("read_seconds", C.c_double),
                ("file_bytes", C.c_uint64), ("index_bytes", C.c_uint64), ("index_buckets", C.c_uint64), ("index_mmer", C.c_int32),
                ("pad_", C.c_int32), ("index_seconds", C.c_double)]


LAYOUT_WIDE, LAYOUT_COMPACT = 0, 1


class _Seq(C.Structure):
    _fields_ = [("name", C.c_char_p), ("seq", C.c_void_p), ("len", C.c_uint64)]


class _FixRow(C.Structure):
    _fields_ = [("seq_index", C.c_uint32), ("coord", C.c_int64), ("mutation", C.c_char * 64), ("original", C.c_char * 64)]


def _fixrow_dtype():
    import numpy as np
    return np.dtype({"names": ["seq_index", "coord", "mutation", "original"], "formats": ["<u4", "<i8", "S64", "S64"],
                     "offsets": [_FixRow.seq_index.offset, _FixRow.coord.offset, _FixRow.mutation.offset, _FixRow.original.offset],
                     "itemsize": C.sizeof(_FixRow)})


try:
    _FIXROW_DTYPE = _fixrow_dtype()
except Exception:                                  # numpy is optional: the slow path needs nothing
    _FIXROW_DTYPE = None


# every symbol include/jasper_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "jb_init", "jb_device_count", "jb_last_error", "jb_version",
    "jb_db_open", "jb_db_open_partitioned", "jb_db_open_part", "jb_db_part_handle", "jb_db_attach_parts", "jb_db_info", "jb_db_replicate", "jb_db_close", "jb_db_dump", "jb_db_histogram", "jb_solid_threshold",
    "jb_lookup", "jb_lookup_flat", "jb_count_positions",
    "jb_polish", "jb_polish_capacity", "jb_host_alloc", "jb_host_free", "jb_batch_upload", "jb_batch_polish", "jb_batch_fetch", "jb_batch_free",
    "jb_result_status", "jb_result_iterations", "jb_result_tally", "jb_result_num_rows", "jb_result_rows",
    "jb_result_num_seqs", "jb_result_seq", "jb_result_free",
    "jb_last_timing_ms", "jb_last_counter",
    "jb_comm_unique_id", "jb_comm_init", "jb_allreduce_tallies", "jb_comm_destroy",
    "jb_gather_bench",
]


def library_path() -> str:
    return os.environ.get("JASPER_B200_LIB", os.path.join(_HERE, "libjasper_b200.so"))


def load_library():
    """Load libjasper_b200.so (built in-tree by __graft_entry__.build() / jasper_b200/build.py)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise JasperError(-4, "libjasper_b200.so is not built (%s): run `python -m jasper_b200.build`; "
                              "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, u64p, u32p, i64p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
    sig = {
        "jb_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
        "jb_device_count": (C.c_int, []),
        "jb_last_error": (C.c_char_p, []),
        "jb_version": (C.c_char_p, []),
        "jb_db_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "jb_db_open_partitioned": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]),
        "jb_db_open_part": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "jb_db_part_handle": (C.c_int, [vp, C.c_char_p]),
        "jb_db_attach_parts": (C.c_int, [vp, C.c_char_p]),
        "jb_db_info": (C.c_int, [vp, C.POINTER(_DbInfo)]),
        "jb_db_replicate": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]),
        "jb_db_close": (None, [vp]),
        "jb_db_dump": (C.c_int, [vp, u64p, u64p, u32p, C.c_uint64, u64p]),
        "jb_db_histogram": (C.c_int, [vp, C.c_uint64, u64p]),
        "jb_solid_threshold": (C.c_int, [u64p, u64p, C.c_size_t]),
        "jb_lookup": (C.c_int, [vp, C.POINTER(C.c_char_p), C.c_size_t, u32p]),
        "jb_lookup_flat": (C.c_int, [vp, C.c_char_p, u64p, C.c_size_t, u32p]),
        "jb_count_positions": (C.c_int, [vp, C.c_char_p, C.c_size_t, u32p]),
        "jb_polish": (C.c_int, [vp, C.POINTER(_Seq), C.c_size_t, C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]),
        "jb_polish_capacity": (C.c_int, [vp, u64p]),
        "jb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
        "jb_host_free": (None, [vp]),
        "jb_batch_upload": (C.c_int, [vp, C.POINTER(_Seq), C.c_size_t, C.POINTER(vp)]),
        "jb_batch_polish": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_int, C.c_int]),
        "jb_batch_fetch": (C.c_int, [vp, C.POINTER(vp)]),
        "jb_batch_free": (None, [vp]),
        "jb_result_status": (C.c_int, [vp]),
        "jb_result_iterations": (C.c_int, [vp]),
        "jb_result_tally": (C.c_int, [vp, C.c_int, i64p, i64p]),
        "jb_result_num_rows": (C.c_int64, [vp, C.c_int]),
        "jb_result_rows": (C.c_int, [vp, C.c_int, C.POINTER(_FixRow), C.c_size_t]),
        "jb_result_num_seqs": (C.c_int64, [vp]),
        "jb_result_seq": (C.c_int, [vp, C.c_size_t, C.POINTER(C.c_void_p), u64p]),
        "jb_result_free": (None, [vp]),
        "jb_last_timing_ms": (C.c_double, [C.c_char_p]),
        "jb_last_counter": (C.c_int64, [C.c_char_p]),
        "jb_comm_unique_id": (C.c_int, [C.c_char_p]),
        "jb_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int]),
        "jb_allreduce_tallies": (C.c_int, [i64p]),
        "jb_comm_destroy": (None, []),
        "jb_gather_bench": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _check(rc: int):
    if rc < 0:
        raise JasperError(rc, load_library().jb_last_error().decode("utf-8", "replace"))
    return rc


def timing_ms(what: str) -> float:
    return load_library().jb_last_timing_ms(what.encode())


def counter(what: str) -> int:
    return load_library().jb_last_counter(what.encode())


class MerDNA(object):
    """Stand-in for jf.MerDNA: it only carries the string; packing, reverse complement and the
    canonical choice happen on the device inside the lookup (JF!swig/mer_dna.i:15,71)."""
    __slots__ = ("s",)

    def __init__(self, s: str = ""):
        self.s = s

    def get_canonical(self):
        return self

    def __str__(self):
        return self.s


@dataclass
class PolishResult:
    status: int
    tallies: list                      # [(wrong, total)] per iteration 0..p
    rows: list                         # per iteration: [[name, coord, mutation, original], ...]
    seqs: list                         # [(name, str)]; raw mode: [(name, memoryview over the library's pinned result buffer)]
    timings: dict = field(default_factory=dict)
    counters: dict = field(default_factory=dict)
    _owner: object = None              # raw mode: keeps the jb_result (and with it the memoryviews) alive


class _ResultOwner(object):
    """Frees a jb_result when the last PolishResult / memoryview user lets go of it."""

    def __init__(self, lib, res):
        self.lib, self.res = lib, res

    def __del__(self):
        try:
            if self.res:
                self.lib.jb_result_free(self.res)
                self.res = None
        except Exception:
            pass


def rows_as_lists(a, names):
    """A raw row array (numpy, _FixRow layout) -> [[name, coord, field3, field4]] as src/jasper.py appends them."""
    idx, coord = a["seq_index"].tolist(), a["coord"].tolist()
    mut = [m.decode("latin-1") for m in a["mutation"].tolist()]
    org = [o.decode("latin-1") for o in a["original"].tolist()]
    return [[names[i], c, m, o] for i, c, m, o in zip(idx, coord, mut, org)]


def rows_csv_bytes(a, names) -> bytes:
    """The body of src/jasper.py's fix.csv (:114-119: csv.writer, delimiter ' ', "\\r\\n" line ends) for a raw row array.  The four
    fields never hold a blank, a quote or a line break (names are first header tokens), so plain joining equals csv's quoting."""
    if len(a) == 0:
        return b""
    idx, coord = a["seq_index"].tolist(), a["coord"].tolist()
    mut, org = a["mutation"].tolist(), a["original"].tolist()
    nb = [n.encode("utf-8") for n in names]
    if any(b'"' in n for n in nb):               # csv would quote such a name: let the csv module do it
        import csv
        import io
        buf = io.StringIO()
        csv.writer(buf, delimiter=" ").writerows(rows_as_lists(a, names))
        return buf.getvalue().encode("utf-8")
    return b"".join(b"%s %d %s %s\r\n" % (nb[i], c, m, o) for i, c, m, o in zip(idx, coord, mut, org))


class PinnedBuffer(object):
    """Page-locked host memory from the library (jb_host_alloc).  Contigs placed in it (`place`) travel to the GPU without a
    staging copy; `view` is a writable memoryview over the whole buffer."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        p = C.c_void_p()
        _check(self._lib.jb_host_alloc(nbytes, C.byref(p)))
        self._p = p
        self.nbytes = nbytes
        self.view = memoryview((C.c_char * max(1, nbytes)).from_address(p.value)).cast("B")[:nbytes]

    def place(self, seqs):
        """Copy [(name, str / bytes-like)] into the buffer back to back -> [(name, memoryview slice)] for polish() / upload()."""
        out, at = [], 0
        for name, s in seqs:
            b = s.encode("latin-1") if isinstance(s, str) else s
            n = len(b)
            if at + n > self.nbytes:
                raise ValueError("PinnedBuffer of %d bytes is too small" % self.nbytes)
            self.view[at:at + n] = b
            out.append((name, self.view[at:at + n]))
            at += n
        return out

    def close(self):
        if getattr(self, "_p", None) is not None and self._p:
            # (views handed out by place() point into this memory: the caller must be done with them)
            self._lib.jb_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QueryMerFile(object):
    """A Jellyfish binary/sorted database resident in the HBM of one GPU."""

    def __init__(self, path: str, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        _check(self._lib.jb_init(1, (C.c_int * 1)(device)))
        _check(self._lib.jb_db_open(os.fsencode(path), device, C.byref(self._h)))
        info = _DbInfo()
        _check(self._lib.jb_db_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in _DbInfo._fields_}
        self.k = info.k
        self.device = device

    @classmethod
    def open_partitioned(cls, path: str, devices):
        """ONE database whose slots are spread over `devices` (the fallback for a table beyond one GPU's HBM, SURVEY.md
        section 8e): returns one QueryMerFile per device; each owns its part and probes the others by peer loads."""
        lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        _check(lib.jb_init(len(devices), devs))
        hs = (C.c_void_p * len(devices))()
        _check(lib.jb_db_open_partitioned(os.fsencode(path), len(devices), devs, hs))
        return [cls._wrap(lib, h, d) for d, h in zip(devices, hs)]

    PART_HANDLE_BYTES = 128

    @classmethod
    def open_part(cls, path: str, device: int, rank: int, world: int):
        """One process per GPU: build and keep part `rank` of the table on `device`.  Exchange part_handle() between the
        ranks (any host-side all-gather), then attach_parts(handles in rank order) before the first lookup."""
        lib = load_library()
        _check(lib.jb_init(1, (C.c_int * 1)(device)))
        h = C.c_void_p()
        _check(lib.jb_db_open_part(os.fsencode(path), device, rank, world, C.byref(h)))
        return cls._wrap(lib, h.value, device)

    def part_handle(self) -> bytes:
        buf = C.create_string_buffer(self.PART_HANDLE_BYTES)
        _check(self._lib.jb_db_part_handle(self._h, buf))
        return buf.raw

    def attach_parts(self, handles) -> None:
        blob = b"".join(handles)
        _check(self._lib.jb_db_attach_parts(self._h, blob))

    @classmethod
    def _wrap(cls, lib, handle, device):
        qf = cls.__new__(cls)
        qf._lib = lib
        qf._h = C.c_void_p(handle)
        info = _DbInfo()
        _check(lib.jb_db_info(qf._h, C.byref(info)))
        qf.info = {f: getattr(info, f) for f, _ in _DbInfo._fields_}
        qf.k = info.k
        qf.device = device
        return qf

    def replicate(self, devices):
        """Copies of this (whole) table on other GPUs of the box: one disk read, device-to-device copies over NVLink."""
        n = len(devices)
        if n == 0:
            return []
        devs = (C.c_int * n)(*devices)
        hs = (C.c_void_p * n)()
        _check(self._lib.jb_db_replicate(self._h, n, devs, hs))
        return [self._wrap(self._lib, h, d) for d, h in zip(devices, hs)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.jb_db_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- qf[mer] ------------------------------------------------------------------------------
    def __getitem__(self, mer) -> int:
        return self.lookup([str(mer)])[0]

    def lookup(self, strings) -> list:
        n = len(strings)
        enc = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
        k = self.k
        parts, offs, pos = [], [0], 0
        for b in enc:
            z = b.find(b"\0")
            b = b[:k] if z < 0 else b[:min(k, z)]          # MerDNA(const char*) stops at NUL / k bytes
            parts.append(b)
            pos += len(b)
            offs.append(pos)
        flat = b"".join(parts)
        out = (C.c_uint32 * max(1, n))()
        _check(self._lib.jb_lookup_flat(self._h, flat, (C.c_uint64 * (n + 1))(*offs), n, out))
        return list(out[:n])

    def count_positions(self, seq):
        import numpy as np
        b = seq.encode("latin-1") if isinstance(seq, str) else bytes(seq)
        out = np.empty(len(b) + 1, dtype=np.uint32)
        _check(self._lib.jb_count_positions(self._h, b, len(b), out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def dump(self):
        import numpy as np
        n = self.info["n_records"]
        lo = np.empty(max(1, n), dtype=np.uint64)
        hi = np.empty(max(1, n), dtype=np.uint64)
        cnt = np.empty(max(1, n), dtype=np.uint32)
        got = C.c_uint64()
        p64, p32 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        _check(self._lib.jb_db_dump(self._h, lo.ctypes.data_as(p64), hi.ctypes.data_as(p64), cnt.ctypes.data_as(p32), n, C.byref(got)))
        return lo[:n], hi[:n], cnt[:n]

    def histogram(self, high: int = 10000):
        """Rows (count, frequency) as `jellyfish histo` prints them."""
        hist = (C.c_uint64 * (high + 2))()
        _check(self._lib.jb_db_histogram(self._h, high, hist))
        return [(c, hist[c]) for c in range(1, high + 2) if hist[c] > 0]

    def gather_bench(self, n_probes: int = 1 << 28, sector_bytes: int = 32, iters: int = 3):
        gbps, gps = C.c_double(), C.c_double()
        _check(self._lib.jb_gather_bench(self._h, n_probes, sector_bytes, iters, C.byref(gbps), C.byref(gps)))
        return gbps.value, gps.value

    # -- polishing ----------------------------------------------------------------------------
    def polish_capacity(self) -> int:
        """Bases one polish() call can take on this device right now (jb_polish_capacity)."""
        n = C.c_uint64()
        _check(self._lib.jb_polish_capacity(self._h, C.byref(n)))
        return n.value

    def _seq_array(self, seqs):
        n = len(seqs)
        arr = (_Seq * max(1, n))()
        keep = []
        for i, (name, s) in enumerate(seqs):
            nb = name.encode("utf-8")
            arr[i].name = nb
            if isinstance(s, memoryview):                    # e.g. a slice of a PinnedBuffer: passed by address, no copy
                s = s.cast("B") if s.format != "B" else s
                view = (C.c_char * len(s)).from_buffer(s) if len(s) else None
                keep.append((s, view, nb))
                arr[i].seq = C.addressof(view) if view is not None else None
                arr[i].len = len(s)
                continue
            b = s.encode("latin-1") if isinstance(s, str) else bytes(s)
            keep.append((b, nb))
            arr[i].seq = C.cast(C.c_char_p(b), C.c_void_p)
            arr[i].len = len(b)
        return arr, keep

    def _collect(self, res, names, raw: bool = False) -> PolishResult:
        """raw=True keeps the result in bulk form: sequences as memoryviews of the library's pinned result buffer (valid while the
        PolishResult lives), the rows of a pass as one numpy structured array (fields seq_index, coord, mutation, original) -- no
        per-row / per-base Python objects and no copy of the bases.  raw results own `res`; otherwise the caller frees it."""
        lib = self._lib
        n_it = lib.jb_result_iterations(res)
        tallies, rows = [], []
        for ite in range(n_it):
            w, t = C.c_int64(), C.c_int64()
            _check(lib.jb_result_tally(res, ite, C.byref(w), C.byref(t)))
            tallies.append((w.value, t.value))
            nr = lib.jb_result_num_rows(res, ite)
            if raw and _FIXROW_DTYPE is not None:
                import numpy as np
                a = np.empty(nr, dtype=_FIXROW_DTYPE)
                _check(lib.jb_result_rows(res, ite, a.ctypes.data_as(C.POINTER(_FixRow)), nr))
                rows.append(a)
                continue
            buf = (_FixRow * max(1, nr))()
            _check(lib.jb_result_rows(res, ite, buf, nr))
            if nr > 256 and _FIXROW_DTYPE is not None:                           # bulk conversion through numpy: per-field ctypes access costs ~2 us each
                import numpy as np
                rows.append(rows_as_lists(np.frombuffer(buf, dtype=_FIXROW_DTYPE, count=nr), names))
            else:
                rows.append([[names[buf[i].seq_index], buf[i].coord, buf[i].mutation.decode("latin-1"), buf[i].original.decode("latin-1")]
                             for i in range(nr)])
        seqs = []
        for i in range(lib.jb_result_num_seqs(res)):
            p, ln = C.c_void_p(), C.c_uint64()
            _check(lib.jb_result_seq(res, i, C.byref(p), C.byref(ln)))
            if raw:                                          # no copy: a view of the pinned buffer the device wrote into
                seqs.append((names[i], memoryview((C.c_char * ln.value).from_address(p.value)).cast("B") if ln.value else memoryview(b"")))
            else:
                seqs.append((names[i], C.string_at(p, ln.value).decode("latin-1") if ln.value else ""))
        out = PolishResult(status=lib.jb_result_status(res), tallies=tallies, rows=rows, seqs=seqs)
        if raw:
            out._owner = _ResultOwner(lib, res)
        for key in ("polish", "polish_scan", "polish_plan", "polish_walk", "polish_kwalk", "polish_recount", "polish_verdicts", "h2d", "d2h"):
            out.timings[key] = timing_ms(key)
        for key in ("scan_lookups", "walk_lookups", "kernel_launches", "events", "probe_steps", "h2d_bytes", "d2h_bytes",
                    "segments", "segment_reruns", "seg_block", "full_scans", "full_scan_lookups", "recount_probes", "max_seg_events", "max_seg_lookups", "max_seg_span",
                    "verdicts", "fast_events", "fast_steps", "verdict2_events", "speculative_passes"):
            out.counters[key] = counter(key)
        return out

    def polish(self, seqs, k: int, solid_thre: int, passes: int, fix: bool = True, raw: bool = False) -> PolishResult:
        """seqs: [(name, str or bytes)].  One call = the whole per-batch loop of src/jasper.py main().  raw: see _collect."""
        import time
        out = None
        t0 = time.perf_counter()
        arr, keep = self._seq_array(seqs)
        res = C.c_void_p()
        t1 = time.perf_counter()
        _check(self._lib.jb_polish(self._h, arr, len(seqs), k, solid_thre, passes, 1 if fix else 0, C.byref(res)))
        t2 = time.perf_counter()
        try:
            out = self._collect(res, [n for n, _ in seqs], raw=raw)
            self.last_wall_ms = {"marshal": (t1 - t0) * 1e3, "jb_polish": (t2 - t1) * 1e3, "collect": (time.perf_counter() - t2) * 1e3,
                                 "jb_upload": timing_ms("wall_upload"), "jb_device": timing_ms("wall_polish"), "jb_fetch": timing_ms("wall_fetch")}
            return out
        finally:
            if not (raw and out is not None):
                self._lib.jb_result_free(res)

    # split form: keep the batch resident and time the device part alone
    def upload(self, seqs):
        arr, keep = self._seq_array(seqs)
        b = C.c_void_p()
        _check(self._lib.jb_batch_upload(self._h, arr, len(seqs), C.byref(b)))
        return _Batch(self, b, [n for n, _ in seqs])


class _Batch(object):
    def __init__(self, qf: QueryMerFile, handle, names):
        self.qf, self.h, self.names = qf, handle, names

    def polish(self, k, solid_thre, passes, fix=True):
        _check(self.qf._lib.jb_batch_polish(self.h, k, solid_thre, passes, 1 if fix else 0))

    def fetch(self, raw: bool = False) -> PolishResult:
        res = C.c_void_p()
        _check(self.qf._lib.jb_batch_fetch(self.h, C.byref(res)))
        out = None
        try:
            out = self.qf._collect(res, self.names, raw=raw)
            return out
        finally:
            if not (raw and out is not None):
                self.qf._lib.jb_result_free(res)

    def free(self):
        if self.h:
            self.qf._lib.jb_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def solid_threshold(rows) -> int:
    """src/jellyfish.py on histogram rows [(count, freq)]; raises JasperError where the reference exits 1."""
    n = len(rows)
    cs = (C.c_uint64 * max(1, n))(*[r[0] for r in rows])
    fs = (C.c_uint64 * max(1, n))(*[r[1] for r in rows])
    return _check(load_library().jb_solid_threshold(cs, fs, n))

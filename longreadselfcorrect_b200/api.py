"""ctypes binding of libpbsc.so (include/pbsc.h) — the host-side mirror of the reference interface.

The names follow the reference: ``Params`` is the option block of ``stride pbcorrect``
(StriDe/PacBioSelfCorrection.cpp:71-101), ``Index`` stands for ``BWTIndexSet`` (pBWT + pRBWT),
``find_interval`` for ``BWTAlgorithms::findInterval``, ``search_seeds`` for
``LongReadProbe::searchSeedsWithHybridKmers``, ``extend_overlap`` for
``LongReadSelfCorrectByOverlap::extendOverlap`` and ``correct_reads`` for
``PacBioSelfCorrectionProcess::process`` over a batch.  There is no CPU fallback: importing this
module fails loudly when the CUDA library has not been built, and every compute call fails when no
GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBSC_LIB") or os.path.join(_PKG, "libpbsc.so")   # PBSC_LIB: an alternative build of the same library (experiments)

PBSC_BWT, PBSC_RBWT = 0, 1


class PbscError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pbsc error {code}: {msg}")
        self.code = code


class CParams(C.Structure):
    _fields_ = [
        ("pb_coverage", C.c_int32), ("error_rate", C.c_double), ("start_kmer", C.c_int32), ("next_target", C.c_int32),
        ("max_leaves", C.c_int32), ("idmer_len", C.c_int32), ("min_kmer", C.c_int32), ("genome", C.c_int32),
        ("mode", C.c_int32), ("manual", C.c_int32), ("adjust", C.c_int32), ("split", C.c_int32), ("no_dp", C.c_int32),
        ("offset", C.c_int32 * 3), ("pool", C.c_int32 * 8), ("n_pool", C.c_int32), ("scan_kmer", C.c_int32),
        ("kmer_up_bound", C.c_int32), ("radius", C.c_int32), ("hh_ratio", C.c_float),
        ("threshold", (C.c_float * 52) * 3), ("freqs_of_kmer", C.c_double * 101),
        ("debug_seed", C.c_int32), ("reserved", C.c_int32),
    ]


class CSeed(C.Structure):
    _fields_ = [("start", C.c_int32), ("len", C.c_int32), ("max_fixed_freq", C.c_int32), ("is_repeat", C.c_int32),
                ("start_best_k", C.c_int32), ("end_best_k", C.c_int32), ("hitchhiked", C.c_int32), ("static_k", C.c_int32)]


class CReadStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("total_reads_len", "corrected_len", "total_seed_num", "total_walk_num",
                                          "high_error_num", "exceed_depth_num", "exceed_leave_num", "fm_num", "dp_num",
                                          "seed_dis")] + [("merge", C.c_int32), ("n_pieces", C.c_int32)]


class CTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("seed_ms", C.c_float), ("extend_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("kernel_launches", C.c_uint64), ("seed_pairs", C.c_uint64),
                ("rank_queries", C.c_uint64), ("dp_ms", C.c_float), ("dp_jobs", C.c_uint64), ("dp_rows", C.c_uint64),
                ("walk_ms", C.c_float), ("walk_launches", C.c_uint64), ("dp_thread_rows", C.c_uint64)]


SEED_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4"), ("max_fixed_freq", "<i4"), ("is_repeat", "<i4"),
                       ("start_best_k", "<i4"), ("end_best_k", "<i4"), ("hitchhiked", "<i4"), ("static_k", "<i4")])
STATS_DTYPE = np.dtype([(n, "<i8") for n in ("total_reads_len", "corrected_len", "total_seed_num", "total_walk_num",
                                             "high_error_num", "exceed_depth_num", "exceed_leave_num", "fm_num",
                                             "dp_num", "seed_dis")] + [("merge", "<i4"), ("n_pieces", "<i4")])

WALK_LOG_DTYPE = np.dtype([("src_start", "<i4"), ("trg_start", "<i4"), ("code", "<i4"), ("dp_failed", "<i4")])

# every symbol include/pbsc.h declares
EXPORTED = ["pbsc_last_error", "pbsc_device_count", "pbsc_params_default", "pbsc_params_derive",
            "pbsc_threshold_table_text", "pbsc_index_create", "pbsc_index_load", "pbsc_index_create_synthetic",
            "pbsc_index_build_prefix_table", "pbsc_index_destroy", "pbsc_index_num_symbols", "pbsc_index_num_strings",
            "pbsc_index_device_bytes", "pbsc_index_get_symbols", "pbsc_findinterval_batch", "pbsc_findinterval_device",
            "pbsc_seed_batch", "pbsc_extend_batch", "pbsc_correct_batch", "pbsc_last_timing",
            "pbsc_batch_upload", "pbsc_batch_run", "pbsc_batch_result_size", "pbsc_batch_fetch", "pbsc_batch_destroy",
            "pbsc_host_alloc", "pbsc_host_free", "pbsc_trim", "pbsc_random_sector_bench",
            "pbsc_index_save", "pbsc_index_load_fmg", "pbsc_index_open", "pbsc_index_clone", "pbsc_index_blob_size",
            "pbsc_index_export_blob", "pbsc_index_import_blob", "pbsc_index_set_lanes", "pbsc_index_lanes",
            "pbsc_host_register", "pbsc_host_unregister", "pbsc_occ_counts",
            "pbsc_batch_debug_size", "pbsc_batch_fetch_debug",
            "pbsc_build_index_files", "pbsc_build_bwt", "pbsc_free"]

_lib = None


def lib() -> C.CDLL:
    """Load libpbsc.so; raise if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m longreadselfcorrect_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.pbsc_last_error.restype = C.c_char_p
    for name in ("pbsc_index_num_symbols", "pbsc_index_num_strings", "pbsc_index_device_bytes"):
        getattr(L, name).restype = C.c_uint64
    L.pbsc_index_destroy.restype = None
    L.pbsc_batch_destroy.restype = None
    L.pbsc_params_default.restype = None
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc < 0:
        raise PbscError(rc, lib().pbsc_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(lib().pbsc_device_count())


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class _Pinned:
    """Owner of one page-locked host block (pbsc_host_alloc); numpy views keep it alive through .base."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        L = lib()
        L.pbsc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.pbsc_host_free.argtypes = [C.c_void_p]
        L.pbsc_host_free.restype = None
        _check(L.pbsc_host_alloc(C.byref(self.ptr), C.c_size_t(max(int(nbytes), 1))))
        self.nbytes = max(int(nbytes), 1)
        self.buf = (C.c_uint8 * self.nbytes).from_address(self.ptr.value)

    def __del__(self):
        try:
            if self.ptr:
                lib().pbsc_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(count: int, dtype=np.uint8) -> np.ndarray:
    """numpy array over page-locked host memory (not initialised)."""
    dt = np.dtype(dtype)
    own = _Pinned(int(count) * dt.itemsize)
    a = np.frombuffer(own.buf, dtype=dt, count=int(count))
    _PINNED_OWNERS[a.ctypes.data] = own
    return a


def pinned_copy(a: np.ndarray) -> np.ndarray:
    b = pinned_empty(a.size, a.dtype)
    b[...] = a.reshape(-1)
    return b


_PINNED_OWNERS: dict = {}


def pinned_free(a: np.ndarray) -> None:
    """Release a pinned_empty() block now (otherwise it lives until the process ends)."""
    _PINNED_OWNERS.pop(a.ctypes.data, None)


def host_register(a: np.ndarray) -> None:
    """Page-lock an existing numpy buffer (e.g. a np.memmap over /dev/shm)."""
    L = lib()
    L.pbsc_host_register.argtypes = [C.c_void_p, C.c_size_t]
    _check(L.pbsc_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))


def host_unregister(a: np.ndarray) -> None:
    L = lib()
    L.pbsc_host_unregister.argtypes = [C.c_void_p]
    L.pbsc_host_unregister.restype = None
    L.pbsc_host_unregister(C.c_void_p(a.ctypes.data))


def build_index_files(packed, prefix: str, device: int = 0, forward: bool = True, reverse: bool = True) -> None:
    """`stride index` on the GPU (pbsc_build.cu): PREFIX.bwt/.sai and PREFIX.rbwt/.rsai, byte-identical to the reference's files.
    packed = (ASCII bases of all reads as one uint8 array, uint64 offsets[n + 1])."""
    bases, offsets = packed
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    flags = (0 if forward else 1) | (0 if reverse else 2)
    L = lib()
    L.pbsc_build_index_files.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_int, C.c_int]
    _check(L.pbsc_build_index_files(C.c_void_p(bases.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_uint64(offsets.size - 1),
                                    prefix.encode(), C.c_int(device), C.c_int(flags)))


def build_bwt(packed, reverse: bool = False, device: int = 0):
    """One strand in memory: (RLUnit bytes, number of symbols, lexicographic read order); Index.from_runs takes the bytes."""
    bases, offsets = packed
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    L = lib()
    L.pbsc_build_bwt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    L.pbsc_free.argtypes = [C.c_void_p]
    L.pbsc_free.restype = None
    runs, lex = C.c_void_p(), C.c_void_p()
    n_runs, n_sym = C.c_uint64(), C.c_uint64()
    _check(L.pbsc_build_bwt(C.c_void_p(bases.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_uint64(n), C.c_int(1 if reverse else 0),
                            C.c_int(device), C.byref(runs), C.byref(n_runs), C.byref(n_sym), C.byref(lex)))
    try:
        r = np.ctypeslib.as_array(C.cast(runs, C.POINTER(C.c_uint8)), shape=(n_runs.value,)).copy() if n_runs.value else np.zeros(0, np.uint8)
        lx = np.ctypeslib.as_array(C.cast(lex, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
    finally:
        L.pbsc_free(runs)
        L.pbsc_free(lex)
    return r, int(n_sym.value), lx


def _concat(strings):
    """list[str] -> (bytes array, uint64 offsets[n+1])"""
    enc = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    buf = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
    return buf, off


@dataclass
class Params:
    """Options of `stride pbcorrect`; `derive()` fills what PacBioSelfCorrectionMain computes (:195-231)."""
    c: CParams

    @staticmethod
    def make(coverage: int = 90, error_rate: float = 0.15, genome: int = 10, kmer: int | None = None,
             unique_offset: int | None = None, repeat_offset: int | None = None, next_target: int = 1,
             max_leaves: int = 32, idmer_len: int = 9, min_kmer: int = 13, mode: int | None = None,
             split: bool = False, no_dp: bool = False, debug_seed: bool = False) -> "Params":
        p = CParams()
        lib().pbsc_params_default(C.byref(p))
        p.pb_coverage, p.error_rate, p.genome = coverage, error_rate, genome
        p.next_target, p.max_leaves, p.idmer_len, p.min_kmer = next_target, max_leaves, idmer_len, min_kmer
        p.split, p.no_dp = int(split), int(no_dp)
        p.debug_seed = int(debug_seed)
        if kmer is not None:
            p.start_kmer, p.adjust = kmer, 1
        if unique_offset is not None:
            p.offset[1], p.adjust = unique_offset, 1
        if repeat_offset is not None:
            p.offset[2], p.adjust = repeat_offset, 1
        if mode is not None:
            p.mode, p.manual = mode, 1
        _check(lib().pbsc_params_derive(C.byref(p)))
        return Params(p)

    def threshold_table_text(self) -> str:
        buf = C.create_string_buffer(8192)
        n = lib().pbsc_threshold_table_text(C.byref(self.c), buf, 8192)
        _check(n)
        return buf.value.decode()

    @property
    def pool(self):
        return [self.c.pool[i] for i in range(self.c.n_pool)]


class Index:
    """Both FM-index strands resident on one GPU (BWTIndexSet::pBWT / pRBWT of the reference)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def load(prefix: str, device: int = 0, require_sai: bool = True) -> "Index":
        h = C.c_void_p()
        _check(lib().pbsc_index_load(prefix.encode(), C.c_int(device), C.c_int(int(require_sai)), C.byref(h)))
        return Index(h)

    @staticmethod
    def from_runs(bwt_runs: np.ndarray, bwt_nsym: int, bwt_nstr: int, rbwt_runs: np.ndarray, rbwt_nsym: int,
                  rbwt_nstr: int, device: int = 0) -> "Index":
        h = C.c_void_p()
        a = np.ascontiguousarray(bwt_runs, dtype=np.uint8)
        b = np.ascontiguousarray(rbwt_runs, dtype=np.uint8)
        _check(lib().pbsc_index_create(_ptr(a, C.c_uint8), C.c_uint64(a.size), C.c_uint64(bwt_nsym), C.c_uint64(bwt_nstr),
                                       _ptr(b, C.c_uint8), C.c_uint64(b.size), C.c_uint64(rbwt_nsym), C.c_uint64(rbwt_nstr),
                                       C.c_int(device), C.byref(h)))
        return Index(h)

    @staticmethod
    def synthetic(n_symbols: int, n_strings: int, seed: int, device: int = 0) -> "Index":
        h = C.c_void_p()
        _check(lib().pbsc_index_create_synthetic(C.c_uint64(n_symbols), C.c_uint64(n_strings), C.c_uint64(seed),
                                                 C.c_int(device), C.byref(h)))
        return Index(h)

    @staticmethod
    def load_fmg(path: str, device: int = 0) -> "Index":
        """PREFIX.fmg: the flat tables as persisted by save()."""
        h = C.c_void_p()
        _check(lib().pbsc_index_load_fmg(path.encode(), C.c_int(device), C.byref(h)))
        return Index(h)

    @staticmethod
    def open(prefix: str, device: int = 0, require_sai: bool = True, k0: int = 13, write_fmg: bool = False):
        """What `pbcorrect -p PREFIX` does: PREFIX.fmg when it matches PREFIX.bwt/.rbwt, else the run-length files.
        Returns (Index, from_fmg)."""
        h = C.c_void_p()
        used = C.c_int(0)
        _check(lib().pbsc_index_open(prefix.encode(), C.c_int(device), C.c_int(int(require_sai)), C.c_int(k0), C.c_int(int(write_fmg)),
                                     C.byref(used), C.byref(h)))
        return Index(h), bool(used.value)

    @staticmethod
    def import_blob(ptr: int, nbytes: int, src_device: int, device: int) -> "Index":
        """New index on `device` from a blob at address `ptr` on GPU src_device (host memory when src_device < 0)."""
        h = C.c_void_p()
        _check(lib().pbsc_index_import_blob(C.c_void_p(ptr), C.c_uint64(nbytes), C.c_int(src_device), C.c_int(device), C.byref(h)))
        return Index(h)

    def save(self, path: str) -> None:
        _check(lib().pbsc_index_save(self._h, path.encode()))

    def clone(self, device: int) -> "Index":
        """A copy on another GPU of the box (peer copies over NVLink)."""
        h = C.c_void_p()
        _check(lib().pbsc_index_clone(self._h, C.c_int(device), C.byref(h)))
        return Index(h)

    def blob_size(self) -> int:
        n = C.c_uint64(0)
        _check(lib().pbsc_index_blob_size(self._h, C.byref(n)))
        return int(n.value)

    def export_blob(self, ptr: int, cap: int) -> None:
        """Write the blob to device memory at `ptr` (on this index's GPU)."""
        _check(lib().pbsc_index_export_blob(self._h, C.c_void_p(ptr), C.c_uint64(cap)))

    def set_lanes(self, lanes: int) -> None:
        """How many batches of this index may run at once (each on its own stream and scratch arena)."""
        _check(lib().pbsc_index_set_lanes(self._h, C.c_int(lanes)))

    def lanes(self) -> int:
        return int(lib().pbsc_index_lanes(self._h))

    def build_prefix_table(self, k0: int) -> None:
        _check(lib().pbsc_index_build_prefix_table(self._h, C.c_int(k0)))

    def close(self) -> None:
        if self._h:
            lib().pbsc_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_symbols(self, which: int) -> int:
        return int(lib().pbsc_index_num_symbols(self._h, C.c_int(which)))

    def num_strings(self, which: int) -> int:
        return int(lib().pbsc_index_num_strings(self._h, C.c_int(which)))

    def device_bytes(self) -> int:
        return int(lib().pbsc_index_device_bytes(self._h))

    def symbols(self, which: int, first: int, count: int) -> bytes:
        buf = np.zeros(max(count, 1), dtype=np.uint8)
        _check(lib().pbsc_index_get_symbols(self._h, C.c_int(which), C.c_uint64(first), C.c_uint64(count), _ptr(buf, C.c_char)))
        return buf[:count].tobytes()

    # BWTAlgorithms::findInterval(pBWT, w) for a list of strings; returns (lower, upper, steps)
    def find_interval(self, which: int, kmers):
        buf, off = _concat(kmers)
        n = len(kmers)
        lo = np.zeros(max(n, 1), dtype=np.int64)
        hi = np.zeros(max(n, 1), dtype=np.int64)
        st = np.zeros(max(n, 1), dtype=np.uint8)
        _check(lib().pbsc_findinterval_batch(self._h, C.c_int(which), _ptr(buf, C.c_char), _ptr(off, C.c_uint64), C.c_uint64(n),
                                             _ptr(lo, C.c_int64), _ptr(hi, C.c_int64), _ptr(st, C.c_uint8)))
        return lo[:n], hi[:n], st[:n]

    # LongReadProbe::searchSeedsWithHybridKmers over a batch; returns (seeds structured array, offsets[n+1])
    def search_seeds(self, params: Params, reads):
        buf, off = _concat(reads)
        n = len(reads)
        cap = int(off[-1] // 10 + 16 * n + 16)
        seeds = np.zeros(cap, dtype=SEED_DTYPE)
        soff = np.zeros(n + 1, dtype=np.uint64)
        need = C.c_uint64(0)
        _check(lib().pbsc_seed_batch(self._h, C.byref(params.c), _ptr(buf, C.c_char), _ptr(off, C.c_uint64), C.c_uint64(n),
                                     seeds.ctypes.data_as(C.POINTER(CSeed)), C.c_uint64(cap), _ptr(soff, C.c_uint64),
                                     C.byref(need), C.c_int(0)))
        return seeds[: int(soff[-1])], soff

    # LongReadSelfCorrectByOverlap(...).extendOverlap for explicit pairs; returns (status[n], merged strings)
    def extend_overlap(self, params: Params, src, path, trg, dis, k, min_sa):
        n = len(src)
        sb, so = _concat(src)
        pb, po = _concat(path)
        tb, to = _concat(trg)
        d = np.asarray(dis, dtype=np.int32)
        kk = np.asarray(k, dtype=np.int32)
        sa = np.asarray(min_sa, dtype=np.int32)
        status = np.zeros(max(n, 1), dtype=np.int32)
        cap = int(sum(int(1.2 * (len(p) + 10)) + 2 * int(x) + len(t) + 32 for p, x, t in zip(path, k, trg))) + 64
        out = np.zeros(cap, dtype=np.uint8)
        ooff = np.zeros(n + 1, dtype=np.uint64)
        _check(lib().pbsc_extend_batch(self._h, C.byref(params.c), C.c_uint64(n), _ptr(sb, C.c_char), _ptr(so, C.c_uint64),
                                       _ptr(pb, C.c_char), _ptr(po, C.c_uint64), _ptr(tb, C.c_char), _ptr(to, C.c_uint64),
                                       _ptr(d, C.c_int32), _ptr(kk, C.c_int32), _ptr(sa, C.c_int32), _ptr(status, C.c_int32),
                                       _ptr(out, C.c_char), C.c_uint64(cap), _ptr(ooff, C.c_uint64)))
        raw = out.tobytes()
        merged = [raw[int(ooff[i]):int(ooff[i + 1])].decode() for i in range(n)]
        return status[:n], merged

    # PacBioSelfCorrectionProcess::process over a batch; returns (pieces per read, stats structured array)
    def correct_reads(self, params: Params, reads=None, packed=None, pinned_out: bool = False, out_bufs=None):
        """PacBioSelfCorrectionProcess::process over a batch on host buffers.  With pinned_out the result buffers are
        page-locked blocks owned by this Index and reused by the next call (their contents are overwritten then).
        out_bufs = (pieces uint8[cap], piece_offsets uint64[n+2], first_piece uint64[n+1], stats STATS_DTYPE[n]): caller-owned
        result buffers (any host memory); raises PbscError(-5) when they are too small."""
        if packed is not None:
            buf, off = packed
            n = off.size - 1
        else:
            buf, off = _concat(reads)
            n = len(reads)
        total = int(off[-1])
        if out_bufs is not None:
            out, poff, first, stats = out_bufs
            need = C.c_uint64(0)
            _check(lib().pbsc_correct_batch(self._h, C.byref(params.c), _ptr(buf, C.c_char), _ptr(off, C.c_uint64), C.c_uint64(n),
                                            _ptr(out, C.c_char), C.c_uint64(out.size), _ptr(poff, C.c_uint64), C.c_uint64(poff.size),
                                            _ptr(first, C.c_uint64), stats.ctypes.data_as(C.POINTER(CReadStats)), C.byref(need)))
            return out, poff, first[: n + 1], stats[:n]
        cap = int(total * 1.3) + 4096 * 4
        while True:
            poff_cap = (total // 10 + 4 * n + 16) if params.c.split else (n + 2)
            if pinned_out:
                have = getattr(self, "_pin_caps", None)
                if have is None or have[0] < cap or have[1] < poff_cap or have[2] < n:
                    for old in getattr(self, "_pin", ()):
                        pinned_free(old)
                    caps = (max(cap, have[0] if have else 0), max(poff_cap, have[1] if have else 0), max(n, have[2] if have else 0))
                    self._pin = (pinned_empty(caps[0], np.uint8), pinned_empty(caps[1], np.uint64), pinned_empty(caps[2] + 1, np.uint64),
                                 pinned_empty(max(caps[2], 1) * STATS_DTYPE.itemsize, np.uint8))
                    self._pin_caps = caps
                out, poff, first = self._pin[0], self._pin[1], self._pin[2][: n + 1]
                stats = np.frombuffer(self._pin[3], dtype=STATS_DTYPE)[: max(n, 1)]
                cap, poff_cap = out.size, poff.size
            else:
                out = np.zeros(cap, dtype=np.uint8)
                poff = np.zeros(poff_cap, dtype=np.uint64)
                first = np.zeros(n + 1, dtype=np.uint64)
                stats = np.zeros(max(n, 1), dtype=STATS_DTYPE)
            need = C.c_uint64(0)
            rc = lib().pbsc_correct_batch(self._h, C.byref(params.c), _ptr(buf, C.c_char), _ptr(off, C.c_uint64), C.c_uint64(n),
                                          _ptr(out, C.c_char), C.c_uint64(cap), _ptr(poff, C.c_uint64), C.c_uint64(poff_cap),
                                          _ptr(first, C.c_uint64), stats.ctypes.data_as(C.POINTER(CReadStats)), C.byref(need))
            if rc == -5 and need.value > cap:
                cap = int(need.value) + 64
                continue
            _check(rc)
            break
        return out, poff, first, stats[:n]

    @staticmethod
    def pieces_as_strings(out, poff, first):
        raw = out.tobytes()
        res = []
        for r in range(first.size - 1):
            res.append([raw[int(poff[j]):int(poff[j + 1])].decode() for j in range(int(first[r]), int(first[r + 1]))])
        return res


class Batch:
    """A batch of reads resident on the device: upload -> run (kernels only) -> fetch (pbsc_batch_* in pbsc.h)."""

    def __init__(self, index: Index, params: Params, reads=None, packed=None):
        if packed is not None:
            self.buf, self.off = packed
        else:
            self.buf, self.off = _concat(reads)
        self.n = self.off.size - 1
        self.params = params
        self.index = index
        self._h = C.c_void_p()
        _check(lib().pbsc_batch_upload(index._h, C.byref(params.c), _ptr(self.buf, C.c_char), _ptr(self.off, C.c_uint64),
                                       C.c_uint64(self.n), C.byref(self._h)))

    def run(self) -> float:
        ms = C.c_float(0)
        _check(lib().pbsc_batch_run(self._h, C.byref(ms)))
        return float(ms.value)

    def fetch(self):
        nb, npieces = C.c_uint64(0), C.c_uint64(0)
        _check(lib().pbsc_batch_result_size(self._h, C.byref(nb), C.byref(npieces)))
        out = np.zeros(max(int(nb.value), 1), dtype=np.uint8)
        poff = np.zeros(int(npieces.value) + 1, dtype=np.uint64)
        first = np.zeros(self.n + 1, dtype=np.uint64)
        stats = np.zeros(max(self.n, 1), dtype=STATS_DTYPE)
        _check(lib().pbsc_batch_fetch(self._h, _ptr(out, C.c_char), C.c_uint64(out.size), _ptr(poff, C.c_uint64), C.c_uint64(poff.size),
                                      _ptr(first, C.c_uint64), stats.ctypes.data_as(C.POINTER(CReadStats))))
        return out, poff, first, stats[: self.n]

    def fetch_debug(self):
        """--debugseed data of a batch that ran with Params.make(debug_seed=True): (seeds, seed_offsets, n_surviving, ratio, log,
        log_offsets); see pbsc_batch_fetch_debug in pbsc.h."""
        ns, nl = C.c_uint64(0), C.c_uint64(0)
        _check(lib().pbsc_batch_debug_size(self._h, C.byref(ns), C.byref(nl)))
        seeds = np.zeros(max(int(ns.value), 1), dtype=SEED_DTYPE)
        soff = np.zeros(self.n + 1, dtype=np.uint64)
        nsurv = np.zeros(max(self.n, 1), dtype=np.uint32)
        nb = int(self.off[-1])
        ratio = np.zeros(max(nb, 1), dtype=np.float32)
        log = np.zeros(max(int(nl.value), 1), dtype=WALK_LOG_DTYPE)
        loff = np.zeros(self.n + 1, dtype=np.uint64)
        _check(lib().pbsc_batch_fetch_debug(self._h, seeds.ctypes.data_as(C.POINTER(CSeed)), C.c_uint64(seeds.size), _ptr(soff, C.c_uint64),
                                            _ptr(nsurv, C.c_uint32), _ptr(ratio, C.c_float), C.c_uint64(ratio.size), C.c_void_p(log.ctypes.data),
                                            C.c_uint64(log.size), _ptr(loff, C.c_uint64)))
        return seeds[: int(ns.value)], soff, nsurv[: self.n], ratio[:nb], log[: int(nl.value)], loff

    def close(self):
        if self._h:
            lib().pbsc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_sector_peak(nbytes: int, device: int = 0) -> float:
    """Measured GB/s of independent random 32-byte sector reads over nbytes of HBM (roofline denominator)."""
    g = C.c_float(0)
    _check(lib().pbsc_random_sector_bench(C.c_int(device), C.c_uint64(int(nbytes)), C.byref(g)))
    return float(g.value)


def occ_counts(reset: bool = True):
    """(is_measurement_build, {family: distinct 32-byte index sectors asked for}) -- zeros unless PBSC_LIB points at the
    -DPBSC_COUNT_OCC build (libpbsc_count.so)."""
    out = (C.c_uint64 * 5)()
    rc = lib().pbsc_occ_counts(out, C.c_int(int(reset)))
    _check(rc)
    return bool(rc), dict(zip(("seed", "setup", "walk", "dp", "other"), (int(x) for x in out)))


def last_timing() -> dict:
    t = CTiming()
    _check(lib().pbsc_last_timing(C.byref(t)))
    return {k: getattr(t, k) for k, _ in CTiming._fields_}

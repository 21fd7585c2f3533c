"""ctypes binding of libraer_b200.so (include/raer_b200.h).

This is the stand-in for R's ``.Call`` into raer.so: it marshals the positional arguments built by
``glue.py`` into ``raer_pileup_args`` / ``raer_scpileup_args`` and unpacks the column vectors. There
is no fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libraer_b200.so")
_LIB = None

COUNT_COLS = ("nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX")
MAX_FILES = 32


class RaerError(RuntimeError):
    pass


class PileupArgs(C.Structure):
    _fields_ = [
        ("nbam", C.c_int32),
        ("bampaths", C.POINTER(C.c_char_p)),
        ("indexes", C.POINTER(C.c_char_p)),
        ("fapath", C.c_char_p),
        ("region", C.c_char_p),
        ("n_regions", C.c_int32),
        ("reg_seqnames", C.POINTER(C.c_char_p)),
        ("reg_start", C.POINTER(C.c_int32)),
        ("reg_end", C.POINTER(C.c_int32)),
        ("int_args", C.c_int32 * 14),
        ("dbl_args", C.c_double * 5),
        ("lgl_args", C.c_int32 * 2),
        ("libtype", C.POINTER(C.c_int32)),
        ("only_keep_variants", C.POINTER(C.c_int32)),
        ("min_mapq", C.POINTER(C.c_int32)),
        ("umi_tag", C.c_char_p),
        ("overlap_mode", C.c_int32),
        ("device", C.c_int32),
        ("host_threads", C.c_int32),
        ("shard_index", C.c_int32),
        ("shard_count", C.c_int32),
        ("aei_mode", C.c_int32),
        ("pad1", C.c_int32),
        ("poll", C.c_void_p),
        ("poll_ctx", C.c_void_p),
    ]


POLL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)  # raer_poll_fn: nonzero return stops the call (RAER_ERR_INTERRUPT)


class PileupResult(C.Structure):
    _fields_ = [
        ("nrows", C.c_int64),
        ("nbam", C.c_int32),
        ("n_targets", C.c_int32),
        ("target_names", C.POINTER(C.c_char_p)),
        ("tid", C.POINTER(C.c_int32)),
        ("pos", C.POINTER(C.c_int32)),
        ("strand", C.POINTER(C.c_char)),
        ("ref", C.POINTER(C.c_char)),
        ("rpbz", C.POINTER(C.c_double)),
        ("vdb", C.POINTER(C.c_double)),
        ("sor", C.POINTER(C.c_double)),
        ("alt", C.POINTER(C.c_char)),
        ("counts", C.POINTER(C.c_int32)),
        ("t_decode", C.c_double),
        ("t_pack", C.c_double),
        ("t_h2d", C.c_double),
        ("t_kernel", C.c_double),
        ("t_d2h", C.c_double),
        ("t_total", C.c_double),
        ("n_records", C.c_int64),
        ("n_aligned_bases", C.c_int64),
        ("gpu_launches", C.c_int64),
        ("aei", C.POINTER(C.c_int64)),
        ("n_inflated_bytes", C.c_int64),
        ("n_seeks", C.c_int64),
        ("n_device_contigs", C.c_int64),
        ("n_host_contigs", C.c_int64),
    ]


class ScArgs(C.Structure):
    _fields_ = [
        ("nbam", C.c_int32),
        ("bampaths", C.POINTER(C.c_char_p)),
        ("indexes", C.POINTER(C.c_char_p)),
        ("region", C.c_char_p),
        ("n_sites", C.c_int32),
        ("site_seqnames", C.POINTER(C.c_char_p)),
        ("site_pos", C.POINTER(C.c_int32)),
        ("site_strand", C.POINTER(C.c_int32)),
        ("site_ref", C.POINTER(C.c_char_p)),
        ("site_alt", C.POINTER(C.c_char_p)),
        ("site_idx", C.POINTER(C.c_int32)),
        ("n_barcodes", C.c_int32),
        ("barcodes", C.POINTER(C.c_char_p)),
        ("cb_tag", C.c_char_p),
        ("int_args", C.c_int32 * 14),
        ("dbl_args", C.c_double * 5),
        ("libtype", C.c_int32),
        ("outfns", C.c_char_p * 3),
        ("umi_tag", C.c_char_p),
        ("remove_overlaps", C.c_int32),
        ("min_mapq", C.c_int32),
        ("min_counts", C.c_int32),
        ("overlap_mode", C.c_int32),
        ("device", C.c_int32),
        ("host_threads", C.c_int32),
        ("pad1", C.c_int32),
        ("poll", C.c_void_p),
        ("poll_ctx", C.c_void_p),
    ]


class ReadHdr(C.Structure):
    _fields_ = [
        ("pos", C.c_int32), ("end", C.c_int32), ("flag", C.c_uint16), ("mapq", C.c_uint8), ("bits", C.c_uint8),
        ("l_qseq", C.c_uint16), ("n_cigar", C.c_uint16), ("cigar_off", C.c_uint32), ("sq_off", C.c_uint32),
        ("mate", C.c_int32), ("aux", C.c_uint32),
    ]


READ_HDR_DTYPE = np.dtype([
    ("pos", "<i4"), ("end", "<i4"), ("flag", "<u2"), ("mapq", "u1"), ("bits", "u1"), ("l_qseq", "<u2"),
    ("n_cigar", "<u2"), ("cigar_off", "<u4"), ("sq_off", "<u4"), ("mate", "<i4"), ("aux", "<u4")])
assert READ_HDR_DTYPE.itemsize == 32 and C.sizeof(ReadHdr) == 32


class ReadBatch(C.Structure):
    _fields_ = [
        ("n_files", C.c_int32), ("tid", C.c_int32), ("beg", C.c_int32), ("end", C.c_int32),
        ("n_reads", C.c_int64),
        ("file_off", C.c_void_p), ("hdr", C.c_void_p), ("maxend", C.c_void_p), ("cigar", C.c_void_p),
        ("seq", C.c_void_p), ("qual", C.c_void_p),
        ("n_cigar_words", C.c_int64), ("n_sq_bytes", C.c_int64),
        ("pair_b", C.c_void_p), ("n_pairs", C.c_int64),
        ("cb", C.c_void_p), ("umi", C.c_void_p),
        ("n_aligned_bases", C.c_int64),
        ("umi_holes", C.c_void_p), ("n_umi_holes", C.c_int64),
    ]


class BulkConf(C.Structure):
    _fields_ = [
        ("n_files", C.c_int32), ("min_depth", C.c_int32), ("min_bq", C.c_int32), ("trim_5p", C.c_int32),
        ("trim_3p", C.c_int32), ("indel_dist", C.c_int32), ("splice_dist", C.c_int32), ("min_overhang", C.c_int32),
        ("nmer", C.c_int32), ("min_var_reads", C.c_int32), ("n_mm_type", C.c_int32), ("n_mm", C.c_int32),
        ("report_multiallelic", C.c_int32), ("remove_overlaps", C.c_int32), ("want_stats", C.c_int32),
        ("umi_mode", C.c_int32), ("aei_mode", C.c_int32), ("pad0", C.c_int32),
        ("ftrim_5p", C.c_double), ("ftrim_3p", C.c_double), ("min_af", C.c_double),
        ("min_mapq", C.c_int32 * MAX_FILES), ("only_keep_variants", C.c_int32 * MAX_FILES),
        ("reg_beg", C.c_int32), ("reg_end", C.c_int32),
        ("site_mask", C.c_void_p), ("mask_beg", C.c_int32), ("mask_bits", C.c_int32),
    ]


class ScConf(C.Structure):
    _fields_ = [
        ("min_bq", C.c_int32), ("min_mapq", C.c_int32), ("trim_5p", C.c_int32), ("trim_3p", C.c_int32),
        ("indel_dist", C.c_int32), ("splice_dist", C.c_int32), ("min_overhang", C.c_int32),
        ("remove_overlaps", C.c_int32), ("has_umi", C.c_int32), ("n_barcodes", C.c_int32), ("n_umi", C.c_uint32),
        ("n_sites", C.c_int32),
        ("ftrim_5p", C.c_double), ("ftrim_3p", C.c_double),
        ("site_pos", C.c_void_p), ("site_strand", C.c_void_p), ("site_ref", C.c_void_p), ("site_alt", C.c_void_p),
    ]


class ScCells(C.Structure):
    _fields_ = [("n", C.c_int64), ("cells", C.POINTER(C.c_int32)), ("n_sites", C.c_int64),
                ("site_totals", C.POINTER(C.c_int32))]


class Rows(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("n_files", C.c_int32),
        ("pos", C.POINTER(C.c_int32)), ("meta", C.POINTER(C.c_uint32)), ("stats", C.POINTER(C.c_double)),
        ("counts", C.POINTER(C.c_int32)), ("altcode", C.POINTER(C.c_uint32)),
        ("grow_pos", C.c_int32 * 64),
    ]


def lib() -> C.CDLL:
    """Load libraer_b200.so; raise loudly if it was not built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise RaerError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(make -C raer_b200/csrc). raer_b200 has no CPU fallback.")
        L = C.CDLL(_SO)
        L.raer_last_error.restype = C.c_char_p
        L.raer_version.restype = C.c_char_p
        L.raer_device_count.restype = C.c_int
        L.raer_pileup.argtypes = [C.POINTER(PileupArgs), C.POINTER(PileupResult)]
        L.raer_pileup.restype = C.c_int
        L.raer_pileup_result_free.argtypes = [C.POINTER(PileupResult)]
        L.raer_scpileup.argtypes = [C.POINTER(ScArgs)]
        L.raer_scpileup.restype = C.c_int
        L.raer_get_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.raer_get_region.restype = C.c_int
        L.raer_fisher_exact.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double)]
        L.raer_fisher_exact.restype = C.c_int
        L.raer_ctx_create.argtypes = [C.c_int32]
        L.raer_ctx_create.restype = C.c_void_p
        L.raer_ctx_destroy.argtypes = [C.c_void_p]
        L.raer_ctx_stream.argtypes = [C.c_void_p]
        L.raer_ctx_stream.restype = C.c_void_p
        L.raer_ctx_set_reference.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int64]
        L.raer_ctx_set_reference.restype = C.c_int
        L.raer_ctx_set_reference_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.raer_ctx_set_reference_device.restype = C.c_int
        L.raer_batch_pileup_host.argtypes = [C.c_void_p, C.POINTER(BulkConf), C.POINTER(ReadBatch), C.POINTER(Rows)]
        L.raer_batch_pileup_host.restype = C.c_int
        L.raer_batch_pileup_device.argtypes = [C.c_void_p, C.POINTER(BulkConf), C.POINTER(ReadBatch),
                                               C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.raer_batch_pileup_device.restype = C.c_int
        L.raer_pipe_create.argtypes = [C.c_int32]
        L.raer_pipe_create.restype = C.c_void_p
        L.raer_pipe_destroy.argtypes = [C.c_void_p]
        L.raer_pipe_destroy.restype = None
        L.raer_pipe_submit.argtypes = [C.c_void_p, C.POINTER(BulkConf), C.POINTER(ReadBatch), C.c_int32, C.c_void_p, C.c_int64]
        L.raer_pipe_submit.restype = C.c_int
        L.raer_pipe_collect.argtypes = [C.c_void_p, C.POINTER(Rows)]
        L.raer_pipe_collect.restype = C.c_int
        L.raer_rows_free.argtypes = [C.POINTER(Rows)]
        L.raer_batch_algorithmic_bytes.argtypes = [C.POINTER(ReadBatch), C.c_int64, C.c_int64]
        L.raer_batch_algorithmic_bytes.restype = C.c_int64
        L.raer_ctx_profile.argtypes = [C.c_void_p, C.c_int]
        L.raer_ctx_profile.restype = None
        L.raer_ctx_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.raer_ctx_profile_read.restype = C.c_int
        L.raer_altcode_to_string.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.c_char_p]
        L.raer_sc_ctx_create.argtypes = [C.c_int32]
        L.raer_sc_ctx_create.restype = C.c_void_p
        L.raer_sc_ctx_destroy.argtypes = [C.c_void_p]
        L.raer_sc_ctx_destroy.restype = None
        L.raer_sc_ctx_stream.argtypes = [C.c_void_p]
        L.raer_sc_ctx_stream.restype = C.c_void_p
        L.raer_sc_batch_device.argtypes = [C.c_void_p, C.POINTER(ScConf), C.POINTER(ReadBatch), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.raer_sc_batch_device.restype = C.c_int
        L.raer_sc_batch_host.argtypes = [C.c_void_p, C.POINTER(ScConf), C.POINTER(ReadBatch), C.POINTER(ScCells)]
        L.raer_sc_batch_host.restype = C.c_int
        L.raer_sc_cells_free.argtypes = [C.POINTER(ScCells)]
        L.raer_sc_cells_free.restype = None
        L.raer_sc_ctx_profile.argtypes = [C.c_void_p, C.c_int]
        L.raer_sc_ctx_profile.restype = None
        L.raer_sc_ctx_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.raer_sc_ctx_profile_read.restype = C.c_int
        L.raer_sc_algorithmic_bytes.argtypes = [C.POINTER(ReadBatch), C.c_int64]
        L.raer_sc_algorithmic_bytes.restype = C.c_int64
        L.raer_ctx_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def last_error() -> str:
    return lib().raer_last_error().decode(errors="replace")


def _strs(xs):
    """char** of the strings (None -> NULL). Site lists repeat a few contig names millions of times: every distinct string
    is encoded once and the pointer array is gathered with numpy (a ctypes array built element by element costs about a
    microsecond per entry)."""
    table, keepalive, addr = {}, [], []
    codes = np.empty(len(xs), dtype=np.int64)
    get = table.get
    for i, x in enumerate(xs):
        k = get(x)
        if k is None:
            b = x if (x is None or isinstance(x, bytes)) else x.encode()
            k = table[x] = len(addr)
            keepalive.append(b)
            addr.append(0 if b is None else C.cast(C.c_char_p(b), C.c_void_p).value)
        codes[i] = k
    arr = np.asarray(addr, dtype=np.uint64)[codes] if len(xs) else np.zeros(1, dtype=np.uint64)
    p = C.cast(arr.ctypes.data, C.POINTER(C.c_char_p))
    p._keep = (arr, keepalive)  # the pointer array and the encoded strings live as long as the pointer object
    return p


def _ints(xs):
    a = np.ascontiguousarray(xs, dtype=np.int32) if len(xs) else np.zeros(1, dtype=np.int32)
    return (C.c_int32 * max(len(xs), 1)).from_buffer_copy(a.tobytes()) if len(xs) else (C.c_int32 * 1)()


LAST_TIMING = {}


def pileup(pc, device: int = 0, host_threads: int = 0, overlap_mode: int = 1, shard=(0, 1), aei: bool = False) -> dict:
    """.Call(".pileup", ...) -> column dict (same layout as the reference's list-of-lists)."""
    L = lib()
    n = pc.n
    a = PileupArgs()
    keep = [_strs(pc.bampaths), _strs(pc.indexes)]
    a.nbam = n
    a.bampaths, a.indexes = keep[0], keep[1]
    a.fapath = pc.fapath.encode()
    a.region = pc.region.encode() if pc.region else None
    if pc.lst is not None and len(pc.lst[0]) > 0:
        a.n_regions = len(pc.lst[0])
        keep += [_strs(list(pc.lst[0])), _ints(pc.lst[1]), _ints(pc.lst[2])]
        a.reg_seqnames, a.reg_start, a.reg_end = keep[-3], keep[-2], keep[-1]
    a.int_args = (C.c_int32 * 14)(*[int(x) for x in pc.int_args])
    a.dbl_args = (C.c_double * 5)(*[float(x) for x in pc.dbl_args])
    a.lgl_args = (C.c_int32 * 2)(*[int(bool(x)) for x in pc.lgl_args])
    keep += [_ints(pc.libtype), _ints([int(bool(x)) for x in pc.only_keep_variants]), _ints(pc.min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3], keep[-2], keep[-1]
    a.umi_tag = pc.umi.encode() if pc.umi else None
    a.overlap_mode = overlap_mode
    a.device = device
    a.host_threads = host_threads
    a.shard_index, a.shard_count = shard
    a.aei_mode = 1 if aei else 0
    r = PileupResult()
    rc = L.raer_pileup(C.byref(a), C.byref(r))
    if rc != 0:
        raise RaerError(f"[raer internal] error detected during pileup (code {rc}): {last_error()}")
    try:
        nr = int(r.nrows)
        names = [r.target_names[i].decode() for i in range(r.n_targets)]

        def arr(ptr, dt, cnt):
            if cnt == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ptr, shape=(cnt,)).astype(dt, copy=True)

        def chars(ptr, width, cnt):
            """fixed-width ASCII column -> numpy unicode array (bytes widened to UCS4 in one vectorised pass)"""
            if cnt == 0:
                return np.zeros(0, dtype=f"<U{width}")
            raw = np.frombuffer(C.string_at(ptr, cnt * width), dtype=np.uint8)
            return raw.astype("<u4").view(f"<U{width}")

        tid = arr(r.tid, np.int32, nr)
        out = {
            "nrows": nr, "contigs": names, "tid": tid,
            "seqname": np.asarray(names, dtype=str)[tid] if nr else np.zeros(0, dtype=str),
            "pos": arr(r.pos, np.int32, nr),
            "strand": chars(r.strand, 1, nr),
            "REF": chars(r.ref, 1, nr),
            "rpbz": arr(r.rpbz, np.float64, nr), "vdb": arr(r.vdb, np.float64, nr), "sor": arr(r.sor, np.float64, nr),
            "files": [],
        }
        if nr:
            alt = chars(r.alt, 12, nr * n)
            cnt = np.ctypeslib.as_array(r.counts, shape=(n, nr, 8)).copy()
        for f in range(n):
            d = {}
            if nr:
                d["ALT"] = alt[f * nr:(f + 1) * nr]
                for k, nm in enumerate(COUNT_COLS):
                    d[nm] = cnt[f, :, k]
            else:
                d["ALT"] = np.zeros(0, dtype=str)
                for nm in COUNT_COLS:
                    d[nm] = np.zeros(0, dtype=np.int32)
            out["files"].append(d)
        LAST_TIMING.clear()
        LAST_TIMING.update(dict(t_decode=r.t_decode, t_pack=r.t_pack, t_h2d=r.t_h2d, t_kernel=r.t_kernel,
                                t_d2h=r.t_d2h, t_total=r.t_total, n_records=int(r.n_records),
                                n_aligned_bases=int(r.n_aligned_bases), gpu_launches=int(r.gpu_launches),
                                n_inflated_bytes=int(r.n_inflated_bytes), n_seeks=int(r.n_seeks),
                                n_device_contigs=int(r.n_device_contigs), n_host_contigs=int(r.n_host_contigs)))
        out["timing"] = dict(LAST_TIMING)
        if aei and r.aei:
            out["aei"] = np.ctypeslib.as_array(r.aei, shape=(n, 4, 4)).astype(np.int64, copy=True)
        return out
    finally:
        L.raer_pileup_result_free(C.byref(r))


class ScTriplets(C.Structure):
    _fields_ = [("n", C.c_int64), ("site_idx", C.POINTER(C.c_int32)), ("cb_idx", C.POINTER(C.c_int32)),
                ("n_ref", C.POINTER(C.c_int32)), ("n_alt", C.POINTER(C.c_int32))]


def scpileup_triplets(sc, device: int = 0, host_threads: int = 0, overlap_mode: int = 1, write_files: bool = False):
    """raer_scpileup_triplets(): the nonzero cells as (site_idx, cb_idx, nRef, nAlt) arrays, 1-based indices like
    the lines of the counts file; the three text files are only written when write_files is set."""
    L = lib()
    L.raer_scpileup_triplets.argtypes = [C.POINTER(ScArgs), C.POINTER(ScTriplets)]
    L.raer_sc_triplets_free.argtypes = [C.POINTER(ScTriplets)]
    a, keep = _sc_args(sc, device, host_threads, overlap_mode)
    if not write_files:
        a.outfns = (C.c_char_p * 3)(None, None, None)
    t = ScTriplets()
    rc = L.raer_scpileup_triplets(C.byref(a), C.byref(t))
    if rc < 0:
        raise RaerError(f"[raer internal] error detected during pileup (code {rc}): {last_error()}")
    try:
        n = int(t.n)
        get = lambda ptr: np.ctypeslib.as_array(ptr, shape=(n,)).astype(np.int64, copy=True) if n else np.zeros(0, dtype=np.int64)  # noqa: E731
        return get(t.site_idx), get(t.cb_idx), get(t.n_ref), get(t.n_alt)
    finally:
        L.raer_sc_triplets_free(C.byref(t))


def scpileup(sc, device: int = 0, host_threads: int = 0, overlap_mode: int = 1) -> int:
    """.Call(".scpileup", ...): writes the three output files, returns the status integer."""
    L = lib()
    a, keep = _sc_args(sc, device, host_threads, overlap_mode)
    rc = L.raer_scpileup(C.byref(a))
    if rc < 0:
        raise RaerError(f"[raer internal] error detected during pileup (code {rc}): {last_error()}")
    return rc


def _sc_args(sc, device, host_threads, overlap_mode):
    a = ScArgs()
    keep = [_strs(sc.bampaths), _strs(sc.indexes)]
    a.nbam = len(sc.bampaths)
    a.bampaths, a.indexes = keep[0], keep[1]
    a.region = sc.query_region.encode() if sc.query_region else None
    lst = sc.lst
    a.n_sites = len(lst[0])
    keep += [_strs(list(lst[0])), _ints(lst[1]), _ints(lst[2]), _strs(list(lst[3])), _strs(list(lst[4])), _ints(lst[5])]
    a.site_seqnames, a.site_pos, a.site_strand, a.site_ref, a.site_alt, a.site_idx = keep[-6:]
    a.n_barcodes = len(sc.barcodes)
    keep.append(_strs(list(sc.barcodes)))
    a.barcodes = keep[-1]
    a.cb_tag = sc.cbtag.encode() if sc.cbtag else None
    a.int_args = (C.c_int32 * 14)(*[int(x) for x in sc.int_args])
    a.dbl_args = (C.c_double * 5)(*[float(x) for x in sc.dbl_args])
    a.libtype = int(sc.libtype)
    a.outfns = (C.c_char_p * 3)(*[o.encode() for o in sc.outfns])
    a.umi_tag = sc.umi.encode() if sc.umi else None
    a.remove_overlaps = int(bool(sc.remove_overlaps))
    a.min_mapq = int(sc.min_mapq)
    a.min_counts = int(sc.min_counts)
    a.overlap_mode = overlap_mode
    a.device = device
    a.host_threads = host_threads
    return a, keep


def get_region(region: str):
    """.Call(".get_region", region) -> dict(chrom, start, end)."""
    buf = C.create_string_buffer(256)
    b, e = C.c_int32(), C.c_int32()
    if lib().raer_get_region(region.encode(), buf, 256, C.byref(b), C.byref(e)) != 0:
        raise RaerError(f"could not parse region:{region}")
    return {"chrom": buf.value.decode(), "start": int(b.value), "end": int(e.value)}


def fisher_exact(mat: np.ndarray) -> np.ndarray:
    """.Call(".fisher_exact", mat): integer matrix with 4 rows (a_ref, a_alt, b_ref, b_alt) per column."""
    mat = np.asarray(mat)
    if mat.ndim != 2 or mat.shape[0] != 4:
        raise ValueError("'mat' must be matrix with 4 rows")
    if not np.issubdtype(mat.dtype, np.integer):
        raise ValueError("'mat' must be an integer matrix")
    flat = np.ascontiguousarray(mat.T.astype(np.int32)).ravel()
    out = np.zeros(mat.shape[1], dtype=np.float64)
    lib().raer_fisher_exact(flat.ctypes.data_as(C.POINTER(C.c_int32)), mat.shape[1],
                            out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def gpu_inflate(payloads, sizes, crcs=None):
    """Test / bench entry of the device DEFLATE decoder (raer_ginflate.cuh): raw DEFLATE `payloads` (bytes objects) of
    known inflated `sizes` -> (list of inflated bytes, status array, kernel milliseconds)."""
    L = lib()
    n = len(payloads)
    c_len = np.array([len(p) for p in payloads], dtype=np.int32)
    c_off = np.zeros(n, dtype=np.int64)
    c_off[1:] = np.cumsum(c_len[:-1], dtype=np.int64)
    u_len = np.asarray(sizes, dtype=np.int32)
    u_off = np.zeros(n, dtype=np.int64)
    u_off[1:] = np.cumsum(u_len[:-1], dtype=np.int64)
    comp = np.frombuffer(b"".join(payloads) + b"\0" * 8, dtype=np.uint8)
    out_len = int(u_len.sum())
    out = np.zeros(out_len + 8, dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    ms = C.c_double(0)
    L.raer_gpu_inflate.restype = C.c_int
    L.raer_gpu_inflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double)]
    crc = None if crcs is None else np.asarray(crcs, dtype=np.uint32)
    rc = L.raer_gpu_inflate(comp.ctypes.data, len(comp) - 8, c_off.ctypes.data, c_len.ctypes.data, u_off.ctypes.data,
                            u_len.ctypes.data, None if crc is None else crc.ctypes.data, n, out.ctypes.data, out_len,
                            status.ctypes.data, C.byref(ms))
    if rc != 0:
        raise RaerError(last_error())
    res = [out[int(u_off[i]):int(u_off[i]) + int(u_len[i])].tobytes() for i in range(n)]
    return res, status, ms.value


def sc_last_device_contigs() -> int:
    """contigs the last scpileup call took through the device ingest"""
    L = lib()
    L.raer_sc_last_device_contigs.restype = C.c_int64
    return int(L.raer_sc_last_device_contigs())

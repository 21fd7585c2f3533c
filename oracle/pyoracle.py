"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py. The product package (raer_b200/) never imports this module.

The call signatures take the same positional vectors the reference passes through
.Call(".pileup") (src/plp.c:1143-1145) and .Call(".scpileup") (src/sc-plp.c:945-948).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COUNT_COLS = ("nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX")


class _PlpArgs(C.Structure):
    _fields_ = [
        ("nbam", C.c_int),
        ("bampaths", C.POINTER(C.c_char_p)),
        ("indexes", C.POINTER(C.c_char_p)),
        ("fapath", C.c_char_p),
        ("region", C.c_char_p),
        ("n_regions", C.c_int),
        ("reg_seqnames", C.POINTER(C.c_char_p)),
        ("reg_start", C.POINTER(C.c_int)),
        ("reg_end", C.POINTER(C.c_int)),
        ("int_args", C.c_int * 14),
        ("dbl_args", C.c_double * 5),
        ("lgl_args", C.c_int * 2),
        ("libtype", C.POINTER(C.c_int)),
        ("only_keep_variants", C.POINTER(C.c_int)),
        ("min_mapq", C.POINTER(C.c_int)),
        ("umi_tag", C.c_char_p),
        ("overlap_mode", C.c_int),
    ]


class _PlpResult(C.Structure):
    _fields_ = [
        ("nrows", C.c_int64),
        ("nbam", C.c_int),
        ("n_targets", C.c_int),
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
    ]


class _ScArgs(C.Structure):
    _fields_ = [
        ("nbam", C.c_int),
        ("bampaths", C.POINTER(C.c_char_p)),
        ("indexes", C.POINTER(C.c_char_p)),
        ("region", C.c_char_p),
        ("n_sites", C.c_int),
        ("site_seqnames", C.POINTER(C.c_char_p)),
        ("site_pos", C.POINTER(C.c_int)),
        ("site_strand", C.POINTER(C.c_int)),
        ("site_ref", C.POINTER(C.c_char_p)),
        ("site_alt", C.POINTER(C.c_char_p)),
        ("site_idx", C.POINTER(C.c_int)),
        ("n_barcodes", C.c_int),
        ("barcodes", C.POINTER(C.c_char_p)),
        ("cb_tag", C.c_char_p),
        ("int_args", C.c_int * 14),
        ("dbl_args", C.c_double * 5),
        ("libtype", C.c_int),
        ("outfns", C.c_char_p * 3),
        ("umi_tag", C.c_char_p),
        ("remove_overlaps", C.c_int),
        ("min_mapq", C.c_int),
        ("min_counts", C.c_int),
        ("overlap_mode", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_pileup.argtypes = [C.POINTER(_PlpArgs), C.POINTER(_PlpResult)]
        L.oracle_pileup.restype = C.c_int
        L.oracle_plp_result_free.argtypes = [C.POINTER(_PlpResult)]
        L.oracle_scpileup.argtypes = [C.POINTER(_ScArgs)]
        L.oracle_scpileup.restype = C.c_int
        L.oracle_calc_vdb.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.oracle_calc_vdb.restype = C.c_double
        L.oracle_calc_mwu_biasZ.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        L.oracle_calc_mwu_biasZ.restype = C.c_double
        L.oracle_calc_sor.argtypes = [C.c_int] * 4
        L.oracle_calc_sor.restype = C.c_double
        L.oracle_kf_erfc.argtypes = [C.c_double]
        L.oracle_kf_erfc.restype = C.c_double
        L.oracle_count_aligned_bases.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]
        L.oracle_count_aligned_bases.restype = C.c_int64
        _LIB = L
    return _LIB


def _strs(xs):
    arr = (C.c_char_p * max(len(xs), 1))()
    for i, x in enumerate(xs):
        arr[i] = x.encode() if isinstance(x, str) else x
    return arr


def _ints(xs):
    return (C.c_int * max(len(xs), 1))(*[int(x) for x in xs])


def pileup(bampaths, fapath, *, region=None, sites=None, int_args, dbl_args, lgl_args, libtype,
           only_keep_variants, min_mapq, umi_tag=None, indexes=None, overlap_mode=1):
    """Oracle of .Call(".pileup"). ``sites`` is (seqnames, start, end) with 1-based inclusive ends.

    Returns a dict: seqname/pos/strand/REF/rpbz/vdb/sor per row, and per-file lists ALT, nRef, ...
    """
    n = len(bampaths)
    a = _PlpArgs()
    a.nbam = n
    keep = [_strs(bampaths)]
    a.bampaths = keep[0]
    if indexes is not None:
        keep.append(_strs(indexes))
        a.indexes = keep[-1]
    a.fapath = fapath.encode() if fapath else None
    a.region = region.encode() if region else None
    if sites is not None and len(sites[0]) > 0:
        a.n_regions = len(sites[0])
        keep += [_strs(list(sites[0])), _ints(sites[1]), _ints(sites[2])]
        a.reg_seqnames, a.reg_start, a.reg_end = keep[-3], keep[-2], keep[-1]
    a.int_args = (C.c_int * 14)(*[int(x) for x in int_args])
    a.dbl_args = (C.c_double * 5)(*[float(x) for x in dbl_args])
    a.lgl_args = (C.c_int * 2)(*[int(bool(x)) for x in lgl_args])
    keep += [_ints(libtype), _ints([int(bool(x)) for x in only_keep_variants]), _ints(min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3], keep[-2], keep[-1]
    a.umi_tag = umi_tag.encode() if umi_tag else None
    a.overlap_mode = overlap_mode
    r = _PlpResult()
    rc = lib().oracle_pileup(C.byref(a), C.byref(r))
    try:
        if rc < 0:
            raise RuntimeError("[oracle] error detected during pileup")
        nr = int(r.nrows)
        names = [r.target_names[i].decode() for i in range(r.n_targets)]

        def arr(ptr, dt, cnt):
            if cnt == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ptr, shape=(cnt,)).astype(dt, copy=True)

        tid = arr(r.tid, np.int32, nr)
        out = {
            "nrows": nr,
            "contigs": names,
            "tid": tid,
            "seqname": [names[t] for t in tid],
            "pos": arr(r.pos, np.int32, nr),
            "strand": np.frombuffer(C.string_at(r.strand, nr), dtype="S1").astype(str) if nr else np.zeros(0, dtype=str),
            "REF": np.frombuffer(C.string_at(r.ref, nr), dtype="S1").astype(str) if nr else np.zeros(0, dtype=str),
            "rpbz": arr(r.rpbz, np.float64, nr),
            "vdb": arr(r.vdb, np.float64, nr),
            "sor": arr(r.sor, np.float64, nr),
            "files": [],
        }
        if nr:
            alt_raw = C.string_at(r.alt, nr * n * 12)
            cnt = np.ctypeslib.as_array(r.counts, shape=(n, nr, 8)).copy()
        for f in range(n):
            d = {}
            if nr:
                blk = alt_raw[f * nr * 12:(f + 1) * nr * 12]
                d["ALT"] = np.array([blk[i * 12:(i + 1) * 12].split(b"\0", 1)[0].decode() for i in range(nr)])
                for k, nm in enumerate(COUNT_COLS):
                    d[nm] = cnt[f, :, k].astype(np.int32)
            else:
                d["ALT"] = np.zeros(0, dtype=str)
                for nm in COUNT_COLS:
                    d[nm] = np.zeros(0, dtype=np.int32)
            out["files"].append(d)
        return out
    finally:
        lib().oracle_plp_result_free(C.byref(r))


def scpileup(bampaths, *, region=None, sites, barcodes, cb_tag, int_args, dbl_args, libtype, outfns,
             umi_tag, remove_overlaps, min_mapq, min_counts, indexes=None, overlap_mode=1):
    """Oracle of .Call(".scpileup"). ``sites`` = (seqnames, pos, strand(1/2), ref, alt, idx). Writes outfns."""
    a = _ScArgs()
    a.nbam = len(bampaths)
    keep = [_strs(bampaths)]
    a.bampaths = keep[0]
    if indexes is not None:
        keep.append(_strs(indexes))
        a.indexes = keep[-1]
    a.region = region.encode() if region else None
    a.n_sites = len(sites[0])
    keep += [_strs(list(sites[0])), _ints(sites[1]), _ints(sites[2]), _strs(list(sites[3])), _strs(list(sites[4])),
             _ints(sites[5])]
    a.site_seqnames, a.site_pos, a.site_strand, a.site_ref, a.site_alt, a.site_idx = keep[-6:]
    a.n_barcodes = len(barcodes)
    keep.append(_strs(list(barcodes)))
    a.barcodes = keep[-1]
    a.cb_tag = cb_tag.encode() if cb_tag else None
    a.int_args = (C.c_int * 14)(*[int(x) for x in int_args])
    a.dbl_args = (C.c_double * 5)(*[float(x) for x in dbl_args])
    a.libtype = int(libtype)
    a.outfns = (C.c_char_p * 3)(*[o.encode() for o in outfns])
    a.umi_tag = umi_tag.encode() if umi_tag else None
    a.remove_overlaps = int(bool(remove_overlaps))
    a.min_mapq = int(min_mapq)
    a.min_counts = int(min_counts)
    a.overlap_mode = overlap_mode
    rc = lib().oracle_scpileup(C.byref(a))
    return rc


def count_aligned_bases(bampath):
    nrec = C.c_int64(0)
    n = lib().oracle_count_aligned_bases(bampath.encode(), C.byref(nrec))
    return int(n), int(nrec.value)

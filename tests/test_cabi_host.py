"""CPU-side checks of the C-ABI library (no GPU): it loads, exports every symbol include/raer_b200.h
declares, refuses to compute without a device (no CPU fallback), and its host logic (region parsing,
Fisher test, BAM ingest / prefilter / mate pairing) agrees with independent computations.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bamtools
from raer_b200 import FilterParam, _cabi, scan_bam_flag
from raer_b200.glue import prepare_pileup_call

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "raer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(raer_[a-z0-9_]+)\s*\(", hdr))
    assert {"raer_pileup", "raer_scpileup", "raer_get_region", "raer_fisher_exact", "raer_batch_pileup_host",
            "raer_batch_pileup_device"} <= names
    L = _cabi.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/raer_b200.h but not exported: {missing}"


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.ReadHdr) == 32
    # spot-check ABI sizes against the C compiler's view
    import subprocess
    import tempfile

    src = r'''
    #include <stdio.h>
    #include "raer_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(raer_pileup_args), sizeof(raer_pileup_result),
        sizeof(raer_scpileup_args), sizeof(raer_read_batch), sizeof(raer_bulk_conf), sizeof(raer_rows),
        sizeof(raer_sc_conf), sizeof(raer_sc_cells)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")],
                       check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    want = [C.sizeof(x) for x in (_cabi.PileupArgs, _cabi.PileupResult, _cabi.ScArgs, _cabi.ReadBatch, _cabi.BulkConf,
                                  _cabi.Rows, _cabi.ScConf, _cabi.ScCells)]
    assert [int(x) for x in out] == want


def test_no_cpu_fallback(ext):
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    pc = prepare_pileup_call(ext.bam1, ext.fasta)
    with pytest.raises(_cabi.RaerError, match="no usable CUDA device"):
        _cabi.pileup(pc)
    assert _cabi.lib().raer_ctx_create(0) is None


def test_get_region():  # tests/testthat/test_utils.R (.get_region)
    assert _cabi.get_region("chr1:1000-2000") == {"chrom": "chr1", "start": 999, "end": 2000}
    assert _cabi.get_region("chr1") == {"chrom": "chr1", "start": 0, "end": 2**31 - 1}
    assert _cabi.get_region("chr1:1,000-2,000") == {"chrom": "chr1", "start": 999, "end": 2000}
    assert _cabi.get_region("HLA-A*01:01:2-3")["chrom"] == "HLA-A*01:01"


def test_fisher_exact_matches_scipy():  # tests/testthat/test_utils.R compares with fisher.test
    from scipy.stats import fisher_exact

    rng = np.random.default_rng(1)
    mat = rng.integers(0, 60, size=(4, 200))
    got = _cabi.fisher_exact(mat)
    want = np.array([fisher_exact([[a, b], [c, d]])[1] for a, b, c, d in mat.T])
    assert np.allclose(got, want, rtol=1e-6, atol=1e-12)
    with pytest.raises(ValueError):
        _cabi.fisher_exact(np.zeros((3, 2), dtype=int))
    with pytest.raises(ValueError):
        _cabi.fisher_exact(np.zeros((4, 2), dtype=float))


def _ingest(pc, target=0):
    L = _cabi.lib()
    L.raer_ingest_summary.argtypes = [C.POINTER(_cabi.PileupArgs), C.c_int64, C.POINTER(C.c_int64)]
    a = _cabi.PileupArgs()
    keep = [_cabi._strs(pc.bampaths), _cabi._strs(pc.indexes)]
    a.nbam = pc.n
    a.bampaths, a.indexes = keep
    a.fapath = pc.fapath.encode()
    a.region = pc.region.encode() if pc.region else None
    if pc.lst:
        a.n_regions = len(pc.lst[0])
        keep += [_cabi._strs(pc.lst[0]), _cabi._ints(pc.lst[1]), _cabi._ints(pc.lst[2])]
        a.reg_seqnames, a.reg_start, a.reg_end = keep[-3:]
    a.int_args = (C.c_int32 * 14)(*pc.int_args)
    a.dbl_args = (C.c_double * 5)(*pc.dbl_args)
    a.lgl_args = (C.c_int32 * 2)(*[int(x) for x in pc.lgl_args])
    keep += [_cabi._ints(pc.libtype), _cabi._ints([int(x) for x in pc.only_keep_variants]), _cabi._ints(pc.min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3:]
    out = (C.c_int64 * 16)()
    rc = L.raer_ingest_summary(C.byref(a), target, out)
    assert rc == 0, _cabi.last_error()
    return dict(zip(("reads", "pairs", "cigar", "sq", "batches", "aligned_bases", "records", "distinct"), out))


def _expected_kept(bam, keep0, keep1, min_mapq=0):
    """readaln (src/plp.c:32-78) restated over the records of a BAM."""
    _, _, recs = bamtools.read_bam(bam)
    kept, bases = [], 0
    for r in recs:
        if r.tid < 0 or r.flag & 4:
            continue
        bases += sum(l for op, l in r.cigar if op in (0, 7, 8))
        t = (keep0 & ~r.flag) | (keep1 & r.flag)
        if (~t) & 2047:
            continue
        if r.mapq < min_mapq:
            continue
        if (r.flag & 1) and not (r.flag & 2):
            continue
        kept.append(r)
    return kept, bases, len(recs)


@pytest.mark.parametrize("target", [0, 50, 7])
def test_ingest_prefilter_matches_readaln(ext, target, monkeypatch):
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta)
    k0, k1 = pc.int_args[12], pc.int_args[13]
    kept1, bases1, n1 = _expected_kept(ext.bam1, k0, k1)
    kept2, bases2, n2 = _expected_kept(ext.bam2, k0, k1)
    # the call carries a region list (one interval per contig, R/pileup.R:251-306): the ingest goes contig by contig
    # through the index and never reads the unmapped tail of the files
    s = _ingest(pc, target)
    assert s["distinct"] == len(kept1) + len(kept2)
    assert s["records"] <= n1 + n2
    assert s["reads"] >= s["distinct"]  # halo reads are re-sent with later batches
    # the single stream reads every record
    monkeypatch.setenv("RAER_SERIAL_INGEST", "1")
    s1 = _ingest(pc, target)
    assert s1["distinct"] == s["distinct"] and s1["reads"] == s["reads"] and s1["pairs"] == s["pairs"]
    assert s1["records"] == n1 + n2
    assert s1["aligned_bases"] == bases1 + bases2


def test_ingest_flags_mapq_and_region(ext):
    fp = FilterParam(min_mapq=255, bam_flags=scan_bam_flag(isMinusStrand=False))
    pc = prepare_pileup_call(ext.bam1, ext.fasta, param=fp)
    s = _ingest(pc)
    kept, _, _ = _expected_kept(ext.bam1, pc.int_args[12], pc.int_args[13], 255)
    assert s["distinct"] == len(kept)
    pc = prepare_pileup_call(ext.bam1, ext.fasta, region="SSR3:203-205")
    s = _ingest(pc)
    kept, _, _ = _expected_kept(ext.bam1, pc.int_args[12], pc.int_args[13])
    n = sum(1 for r in kept if r.tid == 0 and r.pos < 205 and r.pos + (r.ref_len() or 1) > 202)
    assert s["distinct"] == n


def test_ingest_mate_pairs(ext):
    """overlap_push pairing: proper pairs on one contig whose later mate arrives
    while the earlier one is still in the pileup buffer."""
    pc = prepare_pileup_call(ext.bam1, ext.fasta)
    s = _ingest(pc)
    kept, _, _ = _expected_kept(ext.bam1, pc.int_args[12], pc.int_args[13])
    first, pairs, prev_pos, prev_tid = {}, 0, 0, 0
    for r in kept:
        end = r.pos + r.ref_len()
        ok = not (r.flag & 8) and (r.flag & 2)
        if (r.mtid >= 0 and r.tid != r.mtid) or (abs(r.isize) >= 2 * r.l_seq and r.mpos >= end):
            ok = False
        if ok:
            e = first.get(r.qname)
            if e is not None and (e[0] != r.tid or e[1] < prev_pos):
                del first[r.qname]
                e = None
            if e is None:
                if r.mpos >= r.pos:
                    first[r.qname] = (r.tid, end)
            else:
                pairs += 1
                del first[r.qname]
        prev_pos, prev_tid = r.pos, r.tid
    assert s["pairs"] == pairs
    assert pairs > 40


def test_documented_limit_is_refused_before_any_work(ext):
    """umi_tag / the fused AEI reduction with min_base_quality > 128: RAER_ERR_LIMIT up front, with or
    without a device, instead of failing after the qualities were rewritten"""
    pc = prepare_pileup_call(ext.bam1, ext.fasta, param=FilterParam(min_base_quality=200), umi_tag="UB")
    with pytest.raises(_cabi.RaerError, match="needs min_base_quality <= 128"):
        _cabi.pileup(pc)
    pc = prepare_pileup_call(ext.bam1, ext.fasta, param=FilterParam(min_base_quality=200))
    with pytest.raises(_cabi.RaerError, match="needs min_base_quality <= 128"):
        _cabi.pileup(pc, aei=True)


def test_max_depth_replay_matches_the_literal_rule():
    """raer_replay_max_depth (what the device ingest runs when the depth can reach max_depth) against bam_plp_push's rule
    written out with a plain list: a read that starts where the previous pushed read did is dropped while the buffer
    (pushed reads whose end has not fallen behind the iterator) holds max_depth reads"""
    L = _cabi.lib()
    L.raer_replay_max_depth.restype = C.c_int64
    L.raer_replay_max_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    rng = np.random.default_rng(9)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        pos = np.sort(rng.integers(0, max(2, n // int(rng.integers(1, 8))), n)).astype(np.int32)
        end = (pos + rng.integers(0, 12, n)).astype(np.int32)  # zero-length reads included
        md = int(rng.integers(1, 12))
        want = np.zeros(n, dtype=np.uint8)
        live, prev = [], None
        for i in range(n):
            if prev is not None and pos[i] == prev and len(live) + 1 > md:
                live = [e for e in live if e >= prev]
                if len(live) + 1 > md:
                    want[i] = 1
                    continue
            kept = prev is None or end[i] > prev
            prev = pos[i]
            if kept:
                live.append(end[i])
        got = np.ones(n, dtype=np.uint8)
        nd = L.raer_replay_max_depth(pos.ctypes.data, end.ctypes.data, n, md, got.ctypes.data)
        assert np.array_equal(got, want) and nd == int(want.sum())


def test_ingest_max_depth_drop(ext, monkeypatch):
    """the host ingest's max_depth rule (min-heap of live ends) counted against the replay entry and bam_plp_push's
    keep rule, per contig"""
    L = _cabi.lib()
    L.raer_replay_max_depth.restype = C.c_int64
    L.raer_replay_max_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    monkeypatch.setenv("RAER_SERIAL_INGEST", "1")
    for md in (3, 5, 12):
        pc = prepare_pileup_call(ext.bam1, ext.fasta, param=FilterParam(max_depth=md))
        s = _ingest(pc)
        passing, _, _ = _expected_kept(ext.bam1, pc.int_args[12], pc.int_args[13])
        expected = 0
        for tid in sorted({r.tid for r in passing}):
            rs = [r for r in passing if r.tid == tid]
            pos = np.array([r.pos for r in rs], dtype=np.int32)
            end = np.array([r.pos + r.ref_len() for r in rs], dtype=np.int32)
            drop = np.zeros(len(rs), dtype=np.uint8)
            L.raer_replay_max_depth(pos.ctypes.data, end.ctypes.data, len(rs), md, drop.ctypes.data)
            prev = None
            for i in range(len(rs)):
                if drop[i]:
                    continue
                expected += prev is None or end[i] > prev
                prev = pos[i]
        assert s["distinct"] == expected
        if md == 3:
            assert expected < len(passing) - 20  # the rule bites at this depth

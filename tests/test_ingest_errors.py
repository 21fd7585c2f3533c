"""A BAM that cannot be read to its end must fail the call, not shorten the result (the reference raises an error,
src/plp.c:1195; htslib reports truncated files, CRC mismatches and unsorted input). Host-only: raer_ingest_summary
streams the files exactly like raer_pileup() / raer_scpileup() do. Covers: a file cut in the middle of a block, cut at a
block boundary inside a record, a flipped payload byte (CRC32 / inflate), corruption beyond the first inflate group
(the background prefetch task reports it), records out of coordinate order, mutated headers, record fields and BAI
bytes (no crash, no out-of-bounds read: run under the normal allocator, every call must return)."""
import ctypes as C
import os
import random
import shutil

import numpy as np
import pytest

import bamtools
from raer_b200 import _cabi
from raer_b200.glue import prepare_pileup_call


def _summary(bam, fasta, region=None, index=None):
    L = _cabi.lib()
    L.raer_ingest_summary.argtypes = [C.POINTER(_cabi.PileupArgs), C.c_int64, C.POINTER(C.c_int64)]
    pc = prepare_pileup_call(bam, fasta, region=region) if region else prepare_pileup_call(bam, fasta)
    a = _cabi.PileupArgs()
    keep = [_cabi._strs(pc.bampaths), _cabi._strs([index] if index else pc.indexes)]
    a.nbam = pc.n
    a.bampaths, a.indexes = keep
    a.fapath = pc.fapath.encode()
    a.region = pc.region.encode() if pc.region else None
    a.int_args = (C.c_int32 * 14)(*pc.int_args)
    a.dbl_args = (C.c_double * 5)(*pc.dbl_args)
    a.lgl_args = (C.c_int32 * 2)(*[int(x) for x in pc.lgl_args])
    keep += [_cabi._ints(pc.libtype), _cabi._ints([int(x) for x in pc.only_keep_variants]), _cabi._ints(pc.min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3:]
    out = (C.c_int64 * 16)()
    rc = L.raer_ingest_summary(C.byref(a), 0, out)
    return rc, _cabi.last_error(), int(out[6])


def _copy_with_index(src, dst):
    shutil.copy(src, dst)
    shutil.copy(src + ".bai", dst + ".bai")


@pytest.fixture(scope="module")
def big_bam(tmp_path_factory):
    """about 30 MB of inflated records: several 12 MB inflate groups, so the later ones come from the prefetch task"""
    from raer_b200 import synth

    d = str(tmp_path_factory.mktemp("big"))
    return synth.make_bulk_dataset(d, n_bams=1, n_contigs=2, contig_len=400_000, pairs_per_contig=45_000, seed=31)


def test_intact_file_reads_to_the_end(ext):
    rc, msg, n = _summary(ext.bam1, ext.fasta)
    assert rc == 0, msg
    assert n == len(bamtools.read_bam(ext.bam1)[2])


@pytest.mark.parametrize("frac", [0.5, 0.9])
def test_truncated_file_fails(ext, tmp_path, frac):
    p = str(tmp_path / "cut.bam")
    _copy_with_index(ext.bam1, p)
    data = open(ext.bam1, "rb").read()
    open(p, "wb").write(data[: int(len(data) * frac)])
    rc, msg, n = _summary(p, ext.fasta)
    assert rc == -3, (rc, msg, n)  # RAER_ERR_IO
    assert "truncated" in msg or "corrupt" in msg, msg


def test_file_cut_at_a_block_boundary_inside_a_record_fails(ext, tmp_path):
    # whole BGZF blocks only: the framing is intact, the last record is not (blocks of 5000 inflated bytes cut records)
    recs = bamtools.read_bam(ext.bam1)[2]
    raw = bamtools._bgzf_blocks(ext.bam1)
    src = str(tmp_path / "small_blocks.bam")
    with open(src, "wb") as f:
        for k in range(0, len(raw), 5000):
            f.write(bamtools._bgzf_block(raw[k:k + 5000]))
        f.write(bamtools._bgzf_block(b""))
    shutil.copy(ext.bam1 + ".bai", src + ".bai")  # not consulted: the call has no region
    rc, msg, n = _summary(src, ext.fasta)
    assert rc == 0 and n == len(recs), (rc, msg, n)
    data = open(src, "rb").read()
    offs = [0]
    o = 0
    while o + 18 <= len(data):
        o += int.from_bytes(data[o + 16:o + 18], "little") + 1
        offs.append(o)
    assert len(offs) > 20
    failed = 0
    for cut in offs[5:15]:
        p = str(tmp_path / "cutb.bam")
        open(p, "wb").write(data[:cut])
        shutil.copy(src + ".bai", p + ".bai")
        rc, msg, n = _summary(p, ext.fasta)
        if rc != 0:  # a cut that happens to fall between two records reads as a short but well-formed file
            assert rc == -3 and "truncated" in msg, (rc, msg)
            failed += 1
    assert failed >= 8, failed


def test_flipped_payload_byte_fails(ext, tmp_path):
    data = bytearray(open(ext.bam1, "rb").read())
    rng = random.Random(7)
    bad = 0
    for _ in range(12):
        d = bytearray(data)
        i = rng.randrange(200, len(d) - 100)
        d[i] ^= 0x5A
        p = str(tmp_path / "flip.bam")
        open(p, "wb").write(d)
        shutil.copy(ext.bam1 + ".bai", p + ".bai")
        rc, msg, n = _summary(p, ext.fasta)
        bad += rc != 0
        if rc != 0:
            assert rc == -3, (rc, msg)
    # a flip inside a header field that is not checked (MTIME, XFL, OS) is harmless; everything else must be noticed
    assert bad >= 10, bad


def test_corruption_beyond_the_first_inflate_group_fails(big_bam, tmp_path):
    src = big_bam["bams"][0]
    rc, msg, n_ok = _summary(src, big_bam["fasta"])
    assert rc == 0 and n_ok == 180_000, (rc, msg, n_ok)
    data = bytearray(open(src, "rb").read())
    i = int(len(data) * 0.8)
    data[i] ^= 0xFF
    p = str(tmp_path / "late.bam")
    open(p, "wb").write(data)
    shutil.copy(src + ".bai", p + ".bai")
    rc, msg, n = _summary(p, big_bam["fasta"])
    assert rc == -3, (rc, msg, n)
    assert msg, "the message of the background inflate task must reach the caller"
    # the contig-parallel path of raer_pileup() streams through the index: same answer region by region
    rc2, msg2, _ = _summary(p, big_bam["fasta"], region="chr2")
    assert rc2 == -3, (rc2, msg2)


def test_unsorted_input_fails(ext, tmp_path):
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    # two records the prefilter keeps (src/plp.c:32-78): only reads that reach the pileup are order-checked, like htslib
    mapped = [i for i, r in enumerate(recs) if r.tid >= 0 and not (r.flag & (4 | 256 | 512 | 1024 | 2048))
              and (not (r.flag & 1) or (r.flag & 2))]
    i, j = mapped[10], mapped[len(mapped) // 2]
    assert (recs[i].tid, recs[i].pos) != (recs[j].tid, recs[j].pos)
    recs[i], recs[j] = recs[j], recs[i]
    p = str(tmp_path / "unsorted.bam")
    bamtools.write_bam(p, hdr, refs, recs)
    rc, msg, _ = _summary(p, ext.fasta)
    assert rc == -3 and "not sorted" in msg, (rc, msg)
    # chromosomes out of order
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    tids = sorted({r.tid for r in recs if r.tid >= 0})
    assert len(tids) >= 2
    first = [r for r in recs if r.tid == tids[0]]
    second = [r for r in recs if r.tid == tids[1]]
    rest = [r for r in recs if r.tid not in (tids[0], tids[1])]
    bamtools.write_bam(p, hdr, refs, second + first + rest)
    rc, msg, _ = _summary(p, ext.fasta)
    assert rc == -3 and "not sorted" in msg, (rc, msg)


def test_mutated_bam_and_bai_bytes_never_crash(ext, tmp_path):
    """fuzz: header fields, block sizes, record fields and index bytes; any status is fine, returning is the test"""
    rng = random.Random(1234)
    raw = bytearray(bamtools._bgzf_blocks(ext.bam1))
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    p = str(tmp_path / "fz.bam")
    for it in range(40):
        d = bytearray(raw)
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(0, min(len(d), 200_000))
            d[i] = rng.randrange(256)
        # re-frame the mutated inflated stream into well-formed BGZF blocks: the record parser sees the damage
        with open(p, "wb") as f:
            for o in range(0, len(d), 60000):
                f.write(bamtools._bgzf_block(bytes(d[o:o + 60000])))
            f.write(bamtools._bgzf_block(b""))
        shutil.copy(ext.bam1 + ".bai", p + ".bai")
        _summary(p, ext.fasta)
    bai = bytearray(open(ext.bam1 + ".bai", "rb").read())
    shutil.copy(ext.bam1, p)
    for it in range(60):
        d = bytearray(bai)
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(4, len(d))
            d[i] = rng.randrange(256)
        if it % 5 == 0:
            d = d[: rng.randrange(8, len(d))]
        open(p + ".bai", "wb").write(d)
        _summary(p, ext.fasta, region="SSR3:100-400")


def test_negative_l_seq_is_rejected(ext, tmp_path):
    raw = bytearray(bamtools._bgzf_blocks(ext.bam1))
    # first record: after magic(4) l_text(4) text n_ref(4) refs
    l_text = int.from_bytes(raw[4:8], "little")
    o = 8 + l_text
    n_ref = int.from_bytes(raw[o:o + 4], "little")
    o += 4
    for _ in range(n_ref):
        ln = int.from_bytes(raw[o:o + 4], "little")
        o += 8 + ln
    raw[o + 4 + 16:o + 4 + 20] = (-5 & 0xFFFFFFFF).to_bytes(4, "little")  # l_seq of the first record
    p = str(tmp_path / "neg.bam")
    with open(p, "wb") as f:
        for k in range(0, len(raw), 60000):
            f.write(bamtools._bgzf_block(bytes(raw[k:k + 60000])))
        f.write(bamtools._bgzf_block(b""))
    shutil.copy(ext.bam1 + ".bai", p + ".bai")
    rc, msg, _ = _summary(p, ext.fasta)
    assert rc == -3 and "corrupt" in msg, (rc, msg)

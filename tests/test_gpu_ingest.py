"""Device ingest: BGZF blocks inflated by a kernel, records parsed / prefiltered / packed / paired on the
device. Checked against zlib (the inflate) and against the host ingest of the same library (identical result columns),
which in turn is checked against the oracle and the golden vectors elsewhere in this suite.
"""
import os
import zlib

import numpy as np
import pytest

import bamtools
from raer_b200 import FilterParam, _cabi, synth
from raer_b200.glue import Sites, prepare_pileup_call

pytestmark = pytest.mark.gpu


def _raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(data) + co.flush()


def test_inflate_matches_zlib():
    rng = np.random.default_rng(5)
    blocks = []
    text = (b"ACGTTGCA" * 9000)[:65280]
    blocks.append((text, 6, zlib.Z_DEFAULT_STRATEGY))                      # long matches, dynamic Huffman
    blocks.append((rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(), 6, zlib.Z_DEFAULT_STRATEGY))  # incompressible
    blocks.append((rng.integers(0, 256, 40000, dtype=np.uint8).tobytes(), 0, zlib.Z_DEFAULT_STRATEGY))  # stored blocks
    blocks.append((rng.integers(0, 4, 30000, dtype=np.uint8).tobytes(), 6, zlib.Z_FIXED))              # fixed Huffman
    blocks.append((b"", 6, zlib.Z_DEFAULT_STRATEGY))                       # the BGZF EOF block's payload
    blocks.append((b"x", 9, zlib.Z_DEFAULT_STRATEGY))
    blocks.append((bytes(65280), 9, zlib.Z_DEFAULT_STRATEGY))              # one run: distance 1, length 258 chains
    blocks.append((rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(), 1, zlib.Z_HUFFMAN_ONLY))
    for k in range(200):                                                   # BAM-like: skewed bytes with repeats
        n = int(rng.integers(1, 65281))
        a = rng.choice(np.array([0, 1, 2, 17, 33, 65, 255], dtype=np.uint8), n, p=[.3, .2, .15, .1, .1, .1, .05])
        cut = int(rng.integers(0, n))
        a[cut:cut + 300] = a[:300][: max(0, min(300, n - cut))]
        blocks.append((a.tobytes(), int(rng.integers(1, 10)), zlib.Z_DEFAULT_STRATEGY))
    payloads = [_raw_deflate(d, lv, st) for d, lv, st in blocks]
    got, status, ms = _cabi.gpu_inflate(payloads, [len(d) for d, _, _ in blocks])
    assert (status == 0).all(), status[status != 0]
    for (d, _, _), g in zip(blocks, got):
        assert g == d
    assert ms > 0


def test_inflate_reports_corrupt_blocks():
    good = _raw_deflate(b"hello world " * 500)
    bad_type = bytes([0x07]) + good[1:]          # BTYPE 3
    trunc = good[: len(good) // 2]
    flipped = bytearray(good)
    flipped[len(flipped) // 2] ^= 0x5A
    payloads = [good, bad_type, trunc, bytes(flipped), good]
    got, status, _ = _cabi.gpu_inflate(payloads, [6000] * 5)
    assert status[0] == 0 and status[4] == 0 and got[0] == b"hello world " * 500 and got[4] == got[0]
    assert status[1] != 0 and status[2] != 0
    # a flipped byte either fails to decode or decodes to different bytes / another length: never silently equal
    assert status[3] != 0 or got[3] != got[0]
    # wrong expected size
    _, status, _ = _cabi.gpu_inflate([good], [5999])
    assert status[0] != 0


def _columns_equal(a, b):
    assert a["nrows"] == b["nrows"]
    for k in ("tid", "pos", "strand", "REF"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("rpbz", "vdb", "sor"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    for fa, fb in zip(a["files"], b["files"]):
        for k in fa:
            assert np.array_equal(fa[k], fb[k]), k


def _both(pc, monkeypatch, **kw):
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    dev = _cabi.pileup(pc, **kw)
    monkeypatch.setenv("RAER_GPU_INGEST", "0")
    host = _cabi.pileup(pc, **kw)
    return dev, host


CASES = {
    "default": dict(),
    "variants": dict(param=FilterParam(only_keep_variants=True, min_depth=2)),
    "strict": dict(param=FilterParam(min_mapq=30, min_base_quality=25, trim_5p=3, trim_3p=3, indel_dist=4, splice_dist=5,
                                     min_splice_overhang=8, homopolymer_len=5, read_bqual=(0.25, 20.0))),
    "unstranded": dict(param=FilterParam(library_type="unstranded", min_depth=1)),
    "fr-second": dict(param=FilterParam(library_type="fr-second-strand", min_depth=1)),
    "no-overlap-removal": dict(param=FilterParam(remove_overlaps=False)),
    # the MD verdict of parse_mismatches, one bit per read (the bundled BAMs carry MD tags)
    "mismatch-filter": dict(param=FilterParam(max_mismatch_type=(1, 2), min_depth=1)),
    "mismatch-filter-strict": dict(param=FilterParam(max_mismatch_type=(2, 2), only_keep_variants=True, library_type="unstranded")),
}


@pytest.mark.parametrize("inflate", ["host-for-small-ranges", "device-only"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_device_ingest_matches_host_ingest(ext, case, inflate, monkeypatch):
    # file ranges of a few blocks are inflated on the host (one block costs the device decoder 4 ms however few there are);
    # RAER_GI_DEVICE_INFLATE=1 sends them through the kernel as well
    if inflate == "device-only":
        monkeypatch.setenv("RAER_GI_DEVICE_INFLATE", "1")
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta, **CASES[case])
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] >= 1 and dev["timing"]["n_host_contigs"] == 0
    assert host["timing"]["n_device_contigs"] == 0
    assert dev["nrows"] > 0
    _columns_equal(dev, host)
    assert dev["timing"]["n_aligned_bases"] == host["timing"]["n_aligned_bases"]


def test_device_ingest_site_list(ext, monkeypatch):
    base = _cabi.pileup(prepare_pileup_call(ext.bam1, ext.fasta))
    pick = np.arange(0, base["nrows"], 7)
    sites = Sites([base["seqname"][i] for i in pick], [int(base["pos"][i]) for i in pick])
    pc = prepare_pileup_call(ext.bam1, ext.fasta, sites=sites)
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] >= 1
    _columns_equal(dev, host)
    assert dev["nrows"] > 0


def test_device_ingest_synthetic_two_files(tmp_path, monkeypatch):
    """deeper, paired, spliced synthetic reads over two contigs and two files: the pairing through the name-hash sort,
    kept / dropped reads and the per-file layout of the batch"""
    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=2, n_contigs=2, contig_len=60_000, pairs_per_contig=20_000, seed=11)
    pc = prepare_pileup_call(ds["bams"], ds["fasta"], param=FilterParam(min_depth=1, library_type="fr-first-strand"))
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 2 and dev["timing"]["n_host_contigs"] == 0
    _columns_equal(dev, host)
    assert dev["nrows"] > 10000


def test_max_depth_within_reach(ext, tmp_path, monkeypatch):
    """bam_plp_push drops reads that start where the previous one did once the buffer holds max_depth reads, which depends
    on what was dropped before: the device ingest replays that rule over (position, end) of the passing reads on the host
    and keeps the contig on the device"""
    pc = prepare_pileup_call(ext.bam1, ext.fasta, param=FilterParam(max_depth=5))
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] >= 1 and dev["timing"]["n_host_contigs"] == 0
    _columns_equal(dev, host)
    full = _cabi.pileup(prepare_pileup_call(ext.bam1, ext.fasta))
    assert dev["nrows"] > 0 and (dev["nrows"] != full["nrows"] or not np.array_equal(dev["files"][0]["nRef"], full["files"][0]["nRef"]))
    # deep synthetic data, two files, several depth caps (coverage is about 190 per file)
    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=2, n_contigs=2, contig_len=20_000, pairs_per_contig=20_000, seed=23)
    for md in (30, 150, 400):
        pc = prepare_pileup_call(ds["bams"], ds["fasta"], param=FilterParam(max_depth=md, min_depth=1))
        dev, host = _both(pc, monkeypatch)
        assert dev["timing"]["n_device_contigs"] == 2 and dev["timing"]["n_host_contigs"] == 0
        _columns_equal(dev, host)


def test_records_straddling_bgzf_blocks(ext, tmp_path, monkeypatch):
    """the record walk first assumes every BGZF block starts with a record (htslib writers) and verifies it; a file
    whose records straddle blocks fails that check and is walked from the linear index of the .bai instead"""
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    p = str(tmp_path / "straddle.bam")
    bamtools.write_bam_straddling(p, hdr, refs, recs, block_payload=3000)
    pc = prepare_pileup_call(p, ext.fasta)
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] >= 1 and dev["timing"]["n_host_contigs"] == 0
    _columns_equal(dev, host)
    base = _cabi.pileup(prepare_pileup_call(ext.bam1, ext.fasta))
    _columns_equal(dev, base)


def test_linear_index_walk(ext, monkeypatch):
    monkeypatch.setenv("RAER_GI_LINEAR_WALK", "1")
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta)
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] >= 1
    _columns_equal(dev, host)


def test_whole_contig_region_uses_the_device_ingest(ext, monkeypatch):
    """pileup_sites(chroms = ...) / BiocParallel over chromosomes call with region = a contig name (R/pileup.R:208-243)"""
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta, region="SPCS3")
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 1 and dev["timing"]["n_host_contigs"] == 0
    assert dev["nrows"] > 0 and set(dev["seqname"]) == {"SPCS3"}
    _columns_equal(dev, host)
    # a coordinate range stays with the host stream and the index
    pc = prepare_pileup_call(ext.bam1, ext.fasta, region="SPCS3:100-300")
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 0
    _columns_equal(dev, host)


# ---- a BAM that cannot be read to its end fails the call on the device path too (src/plp.c:1195; tests/test_ingest_errors.py
# holds the host path to the same rules)
def _device_call_fails(bam, fasta, monkeypatch, match):
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    pc = prepare_pileup_call(bam, fasta)
    with pytest.raises(_cabi.RaerError, match=match):
        _cabi.pileup(pc)


def _last_mapped_coffset(bai_path):
    """file offset of the block in which the last indexed (mapped) record ends"""
    import struct

    d = open(bai_path, "rb").read()
    n_ref = struct.unpack_from("<i", d, 4)[0]
    o, last = 8, 0
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", d, o)[0]
        o += 4
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", d, o)
            o += 8
            for _ in range(n_chunk):
                vb, ve = struct.unpack_from("<QQ", d, o)
                o += 16
                if b != 37450:
                    last = max(last, ve >> 16)
        n_intv = struct.unpack_from("<i", d, o)[0]
        o += 4 + 8 * n_intv
    return last


@pytest.mark.parametrize("frac", [0.3, 0.8])
def test_device_ingest_truncated_file_fails(ext, tmp_path, monkeypatch, frac):
    """cuts inside the mapped records (a cut in the unmapped tail is not seen by an indexed read, here or in htslib)"""
    import shutil

    p = str(tmp_path / "cut.bam")
    data = open(ext.bam1, "rb").read()
    frac = frac * _last_mapped_coffset(ext.bam1 + ".bai") / len(data)
    open(p, "wb").write(data[: int(len(data) * frac)])
    shutil.copy(ext.bam1 + ".bai", p + ".bai")
    # without the BGZF end-of-file block the call goes to the host stream, which reads to the cut and reports it
    _device_call_fails(p, ext.fasta, monkeypatch, "truncated|corrupt|inflate failed")
    # a file that keeps its end-of-file block but lost blocks in front of it: the index points beyond the data
    if frac < 0.5:
        open(p, "wb").write(data[: int(len(data) * frac)] + bamtools._EOF)
        _device_call_fails(p, ext.fasta, monkeypatch, "truncated|corrupt|inflate failed|not sorted")


def test_device_ingest_flipped_payload_bytes_fail(tmp_path, monkeypatch):
    """a flipped bit in a block's payload either breaks the DEFLATE stream or the block's CRC-32 (verified by the inflate
    kernel): never silently other bytes. The file is large enough that the header read does not touch the damaged block."""
    import random
    import shutil

    ds = synth.make_bulk_dataset(str(tmp_path / "d"), n_bams=1, n_contigs=1, contig_len=400_000, pairs_per_contig=45_000, seed=5)
    src = ds["bams"][0]
    data = open(src, "rb").read()
    offs, o = [], 0
    while o + 18 <= len(data):
        bsize = int.from_bytes(data[o + 16:o + 18], "little")
        offs.append((o, bsize + 1))
        o += bsize + 1
    last = _last_mapped_coffset(src + ".bai")
    inner = [(bo, bl) for bo, bl in offs if bo > (1 << 20) and bo + bl < last]
    assert len(inner) > 20
    rng = random.Random(11)
    p = str(tmp_path / "flip.bam")
    shutil.copy(src + ".bai", p + ".bai")
    for _ in range(6):
        bo, bl = inner[rng.randrange(len(inner))]
        d = bytearray(data)
        d[bo + 18 + rng.randrange(0, bl - 18 - 8)] ^= 1 << rng.randrange(8)
        open(p, "wb").write(d)
        _device_call_fails(p, ds["fasta"], monkeypatch, "inflate failed")
    # the CRC itself: a damaged footer is noticed as well
    bo, bl = inner[3]
    d = bytearray(data)
    d[bo + bl - 8] ^= 0x10
    open(p, "wb").write(d)
    _device_call_fails(p, ds["fasta"], monkeypatch, r"inflate failed .*-21")
    monkeypatch.setenv("RAER_SKIP_CRC", "1")  # the check can be switched off like on the host
    assert _cabi.pileup(prepare_pileup_call(p, ds["fasta"]))["nrows"] > 0


def test_device_ingest_unsorted_input_fails(ext, tmp_path, monkeypatch):
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    mapped = [i for i, r in enumerate(recs) if r.tid == 0 and not (r.flag & (4 | 256 | 512 | 1024 | 2048))
              and (not (r.flag & 1) or (r.flag & 2))]
    i, j = mapped[10], mapped[len(mapped) // 2]
    assert recs[i].pos != recs[j].pos
    recs[i], recs[j] = recs[j], recs[i]
    p = str(tmp_path / "unsorted.bam")
    bamtools.write_bam(p, hdr, refs, recs)
    _device_call_fails(p, ext.fasta, monkeypatch, "not sorted")
    # chromosomes out of order
    hdr, refs, recs = bamtools.read_bam(ext.bam1)
    tids = sorted({r.tid for r in recs if r.tid >= 0})
    first = [r for r in recs if r.tid == tids[0]]
    second = [r for r in recs if r.tid == tids[1]]
    rest = [r for r in recs if r.tid not in (tids[0], tids[1])]
    half = len(first) // 2
    bamtools.write_bam(p, hdr, refs, first[:half] + second + first[half:] + rest)
    _device_call_fails(p, ext.fasta, monkeypatch, "not sorted")


# ---- single cell: screadaln on the device (cell-barcode whitelist, UMI codes)
def _sc_triplets(sce):
    o = np.lexsort((sce.col, sce.row))
    return (sce.shape, list(sce.barcodes), sce.row[o].tolist(), sce.col[o].tolist(), sce.nRef[o].tolist(), sce.nAlt[o].tolist())


@pytest.mark.parametrize("umi_tag", ["UB", None])
def test_single_cell_device_ingest_matches_host_ingest(tmp_path, monkeypatch, umi_tag):
    from raer_b200 import api

    ds = synth.make_sc_dataset(str(tmp_path / "sc"))
    gr = Sites(**ds["sites"])
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    dev = api.pileup_cells_triplets(ds["bam"], gr, ds["barcodes"], umi_tag=umi_tag, param=fp)
    assert _cabi.sc_last_device_contigs() >= 1
    monkeypatch.setenv("RAER_GPU_INGEST", "0")
    host = api.pileup_cells_triplets(ds["bam"], gr, ds["barcodes"], umi_tag=umi_tag, param=fp)
    assert _cabi.sc_last_device_contigs() == 0
    assert _sc_triplets(dev) == _sc_triplets(host)
    assert sum(_sc_triplets(dev)[4]) > 1000


def test_single_cell_device_ingest_on_the_bundled_10x_bam(ext, monkeypatch):
    """the reference's own 10x example (known answers in test_sc_synth_parity / the golden suite): sparse sites there, so
    the site list is widened until the contig qualifies for the device ingest"""
    from raer_b200 import api

    _, refs, recs = bamtools.read_bam(ext.scbam)
    cbs = list(dict.fromkeys(r.tag_z("CB") for r in recs if r.tag_z("CB") is not None))
    length = dict(refs)["2"]
    pos = sorted(set([579, 589, 601, 625, 645] + list(range(1, length, 997))))
    gr = Sites(["2"] * len(pos), pos, strand=["-"] * len(pos), ref=["A"] * len(pos), alt=["G"] * len(pos))
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    dev = api.pileup_cells_triplets(ext.scbam, gr, cbs, param=fp)
    assert _cabi.sc_last_device_contigs() == 1
    monkeypatch.setenv("RAER_GPU_INGEST", "0")
    host = api.pileup_cells_triplets(ext.scbam, gr, cbs, param=fp)
    assert _sc_triplets(dev) == _sc_triplets(host)
    assert sum(_sc_triplets(dev)[4]) > 100


# ---- position windows: a contig whose reads exceed what one batch's 32-bit offsets describe is served in several batches
@pytest.mark.parametrize("window_bytes", ["4096", "20000", "150000"])
def test_contig_served_in_position_windows(tmp_path, monkeypatch, window_bytes):
    """RAER_GI_WINDOW_BYTES forces many small windows: halo reads, mates kept with their partners across a boundary,
    pairs, per-window offsets; rows must not depend on where the boundaries fall"""
    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=2, n_contigs=2, contig_len=50_000, pairs_per_contig=12_000, seed=19)
    pc = prepare_pileup_call(ds["bams"], ds["fasta"], param=FilterParam(min_depth=1, library_type="fr-first-strand"))
    monkeypatch.setenv("RAER_GI_WINDOW_BYTES", window_bytes)
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 2 and dev["timing"]["n_host_contigs"] == 0
    _columns_equal(dev, host)
    monkeypatch.delenv("RAER_GI_WINDOW_BYTES")
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    one = _cabi.pileup(pc)
    _columns_equal(dev, one)
    assert dev["nrows"] > 10000


def test_position_windows_on_the_golden_files_and_single_cell(ext, tmp_path, monkeypatch):
    from raer_b200 import api

    monkeypatch.setenv("RAER_GI_WINDOW_BYTES", "4096")
    for kw in (dict(), dict(param=FilterParam(only_keep_variants=True, min_depth=2)),
               dict(param=FilterParam(library_type="unstranded", trim_5p=4, splice_dist=3))):
        pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta, **kw)
        dev, host = _both(pc, monkeypatch)
        assert dev["timing"]["n_device_contigs"] >= 1
        _columns_equal(dev, host)
    ds = synth.make_sc_dataset(str(tmp_path / "sc"))
    gr = Sites(**ds["sites"])
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    dev = api.pileup_cells_triplets(ds["bam"], gr, ds["barcodes"], param=fp)
    assert _cabi.sc_last_device_contigs() >= 1
    monkeypatch.setenv("RAER_GPU_INGEST", "0")
    host = api.pileup_cells_triplets(ds["bam"], gr, ds["barcodes"], param=fp)
    assert _sc_triplets(dev) == _sc_triplets(host)


# ---- coordinate ranges ("contig:beg-end": tiles of a contig cut for several GPUs, user regions)
@pytest.mark.parametrize("region", ["chr1:1-60000", "chr1:20001-40000", "chr2:35000-60000", "chr1:1-17000", "chr2:52000-60000"])
def test_coordinate_range_through_the_device_ingest(tmp_path, monkeypatch, region):
    """the byte range of a coordinate range comes from the linear index (first record that reaches the window the range
    begins in .. an entry behind its end, verified by looking at the first record behind it); rows equal the host
    stream's for the same region string"""
    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=2, n_contigs=2, contig_len=60_000, pairs_per_contig=15_000, seed=29)
    monkeypatch.setenv("RAER_GI_REGION_MIN", "0")
    pc = prepare_pileup_call(ds["bams"], ds["fasta"], region=region, param=FilterParam(min_depth=1, library_type="fr-first-strand"))
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 1 and dev["timing"]["n_host_contigs"] == 0
    assert host["timing"]["n_device_contigs"] == 0
    _columns_equal(dev, host)
    assert dev["nrows"] > 100
    assert dev["timing"]["n_aligned_bases"] == host["timing"]["n_aligned_bases"]
    monkeypatch.setenv("RAER_GI_WINDOW_BYTES", "20000")
    monkeypatch.setenv("RAER_GPU_INGEST", "1")
    _columns_equal(_cabi.pileup(pc), host)


def test_narrow_coordinate_ranges_stay_with_the_host_stream(ext, monkeypatch):
    pc = prepare_pileup_call(ext.bam1, ext.fasta, region="SSR3:203-205")
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_device_contigs"] == 0
    _columns_equal(dev, host)

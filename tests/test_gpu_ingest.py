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
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_device_ingest_matches_host_ingest(ext, case, monkeypatch):
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


def test_device_ingest_falls_back_where_it_must(ext, monkeypatch):
    """max_depth within reach: the stateful drop of bam_plp_push needs the stream replayed on the host"""
    pc = prepare_pileup_call(ext.bam1, ext.fasta, param=FilterParam(max_depth=5))
    dev, host = _both(pc, monkeypatch)
    assert dev["timing"]["n_host_contigs"] >= 1
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

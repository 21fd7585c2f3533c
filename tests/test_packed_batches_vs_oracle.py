"""The batches bench.py measures are laid out by synth.pack_device_batch (mate indices, overlap coin, maxend, padding),
not by the BAM ingest. This test pins that packer -- and with it the benched workload -- to the CPU oracle: the same
generated records go (a) through pack_device_batch + raer_batch_pileup_host and (b) through write_bam + the oracle's
own BAM reader and htslib restatement. Rows must be bit-exact, rpbz / vdb / sor within 1e-6 relative (the north
star's tolerance). 2 contigs x 2 files at the C2 depth (100x per file), mate-overlap removal on and off, all three
tile kernels."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_backend
from test_batch_layer import _rows_to_np

pytestmark = pytest.mark.gpu

CONTIG_LEN = 30_000
PAIRS = 15_000  # 2 x 100 bp x 15000 / 30 kb = 100x per file
PREFIX = ["b0_", "b1_"]


@pytest.fixture(scope="module")
def packed(tmp_path_factory):
    import torch

    from raer_b200 import synth

    d = str(tmp_path_factory.mktemp("packed"))
    names = ["chr1", "chr2"]
    lens = [CONTIG_LEN, CONTIG_LEN]
    refs = synth.make_reference(lens, 9001)
    fa = os.path.join(d, "ref.fa")
    synth.write_fasta(fa, names, refs)
    per = [[synth.gen_contig(refs[t], PAIRS, 9100 + 1000 * b + t) for t in range(2)] for b in range(2)]
    bams = []
    for b in range(2):
        p = os.path.join(d, f"s{b + 1}.bam")
        synth.write_bam(p, names, lens, per[b], qname_prefix=PREFIX[b])
        bams.append(p)
    return dict(fasta=fa, bams=bams, names=names, refs=refs, per=per, torch=torch)


def _device_rows(packed, remove_overlaps, okv):
    """rows of both contigs through the batch layer of the C ABI, decoded to the columns of the .Call result"""
    from raer_b200 import _cabi, synth

    L = _cabi.lib()
    L.raer_altcode_to_string.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.c_char_p]
    L.raer_altcode_to_string.restype = None
    nf = 2
    conf = _cabi.BulkConf()
    conf.n_files = nf
    conf.min_depth, conf.min_bq = 1, 20
    conf.report_multiallelic, conf.want_stats = 1, 1
    conf.remove_overlaps = 1 if remove_overlaps else 0
    for f in range(nf):
        conf.min_mapq[f] = 0
        conf.only_keep_variants[f] = okv
    conf.reg_beg, conf.reg_end = 0, 2**31 - 1
    kh = (C.c_int32 * (2 * nf))(*([4] * (2 * nf)))  # khash state of every (file, strand), carried across the call
    out = dict(seq=[], pos=[], strand=[], REF=[], stats=[], counts=[], alt=[])
    ctx = L.raer_ctx_create(0)
    assert ctx, _cabi.last_error()
    try:
        for t in range(2):
            per = [packed["per"][b][t] for b in range(nf)]
            T, rb = synth.pack_device_batch(per, t, 0, CONTIG_LEN, remove_overlaps=remove_overlaps, qname_prefixes=PREFIX)
            ref = synth.reference_ascii(packed["refs"][t]).contiguous()
            assert L.raer_ctx_set_reference(ctx, t, C.c_char_p(ref.data_ptr()), ref.numel()) == 0
            rows = _cabi.Rows()
            rc = L.raer_batch_pileup_host(ctx, C.byref(conf), C.byref(rb), C.byref(rows))
            assert rc == 0, _cabi.last_error()
            grow = [rows.grow_pos[k] for k in range(2 * nf)]
            r = _rows_to_np(rows, nf)
            L.raer_rows_free(C.byref(rows))
            buf = C.create_string_buffer(16)
            for i in range(r["pos"].size):
                s = int(r["meta"][i] & 1)
                refc = int((r["meta"][i] >> 8) & 0xFF)
                alts = []
                for f in range(nf):
                    k = f * 2 + s
                    if kh[k] == 4 and grow[k] < r["pos"][i]:
                        kh[k] = 8
                    one = (C.c_int32 * 1)(kh[k])
                    L.raer_altcode_to_string(int(r["alt"][i, f]), refc, one, buf)
                    kh[k] = one[0]
                    alts.append(buf.value.decode())
                out["alt"].append(alts)
                out["strand"].append("-" if s else "+")
                out["REF"].append(chr(refc))
            for k in range(2 * nf):
                if kh[k] == 4 and grow[k] != 2**31 - 1:
                    kh[k] = 8
            out["seq"] += [packed["names"][t]] * r["pos"].size
            out["pos"].append(r["pos"] + 1)
            out["stats"].append(r["stats"])
            out["counts"].append(r["counts"])
    finally:
        L.raer_ctx_destroy(ctx)
    out["pos"] = np.concatenate(out["pos"])
    out["stats"] = np.concatenate(out["stats"])
    out["counts"] = np.concatenate(out["counts"])
    return out


@pytest.mark.parametrize("kernel", ["lane", "fast", "generic"])
@pytest.mark.parametrize("okv", [1, 0])
@pytest.mark.parametrize("remove_overlaps", [True, False])
def test_packed_batches_equal_the_oracle_on_the_same_records(packed, remove_overlaps, okv, kernel, monkeypatch):
    from raer_b200 import FilterParam

    if kernel != "lane":
        monkeypatch.setenv("RAER_KERNEL", kernel)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=bool(okv), remove_overlaps=remove_overlaps)
    want = oracle_backend.pileup_sites(packed["bams"], packed["fasta"], param=fp)
    got = _device_rows(packed, remove_overlaps, okv)
    assert len(want) == got["pos"].size and len(want) > (3000 if okv else 50_000)
    assert list(want.seqnames) == got["seq"]
    assert np.array_equal(want.pos, got["pos"])
    assert list(want.strand) == got["strand"]
    assert list(want.REF) == got["REF"]
    cols = ["nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"]
    for k, nm in enumerate(cols):
        m = want.assays[nm] if nm in want.assays else want.extra[nm]
        assert np.array_equal(np.asarray(m), got["counts"][:, :, k]), nm
    alt = want.assays["ALT"] if "ALT" in want.assays else want.extra["ALT"]
    assert [list(a) for a in np.asarray(alt)] == got["alt"]
    for j, nm in enumerate(("rpbz", "vdb", "sor")):
        a, b = getattr(want, nm), got["stats"][:, j]
        fin = np.isfinite(a)
        assert np.array_equal(fin, np.isfinite(b)), nm
        assert np.array_equal(a[~fin], b[~fin]), nm
        assert np.allclose(a[fin], b[fin], rtol=1e-6, atol=0), nm

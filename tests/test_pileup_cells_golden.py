"""Known-answer tests of the single-cell pileup engine.

Expectations come from the reference's tests/testthat/test_pileup_cells.R and README.md:116-175.
Run against the CPU oracle (not gpu) and the CUDA product through the C-ABI (gpu).
"""
import os

import numpy as np
import pytest

import bamtools
from raer_b200 import FilterParam, scan_bam_flag
from raer_b200.glue import Sites


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def eng(request):
    if request.param == "oracle":
        import oracle_backend

        return oracle_backend
    from raer_b200 import api

    return api


@pytest.fixture(scope="module")
def cbs(ext):
    _, _, recs = bamtools.read_bam(ext.scbam)
    seen, out = set(), []
    for r in recs:
        cb = r.tag_z("CB")
        if cb is not None and cb not in seen:
            seen.add(cb)
            out.append(cb)
    assert len(out) == 556
    return out


def gr5():
    # sort(GRanges(c("2:579:-","2:625:-","2:645:-","2:589:-","2:601:-"))) with REF/ALT as in the test
    pos = [579, 589, 601, 625, 645]
    ref = ["A", "A", "T", "A", "A"]
    alt = ["G", "G", "C", "G", "G"]
    return Sites(["2"] * 5, pos, strand=["-"] * 5, ref=ref, alt=alt)


def test_basic(eng, ext, cbs, tmp_path):  # test_pileup_cells.R:27-63
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    sce = eng.pileup_cells(ext.scbam, gr5(), cbs, str(tmp_path), param=fp)
    assert sce.shape == (5, 556)
    assert sce.barcodes == cbs
    for f in ("barcodes.txt.gz", "sites.txt.gz", "counts.mtx.gz"):
        assert os.path.exists(tmp_path / f)
    assert sce.nRef.sum() > 0 and sce.nAlt.sum() > 0
    # per-site totals reproduced by the survey probe; tests pin 230 and 7
    assert list(zip(sce.row_sums("nRef"), sce.row_sums("nAlt"))) == [(103, 127), (273, 7), (3, 277), (149, 135), (0, 0)]
    sce = eng.pileup_cells(ext.scbam, gr5(), ["non", "existent", "barcodes"], str(tmp_path), param=fp)
    assert sce.nRef.sum() + sce.nAlt.sum() == 0
    with pytest.raises(Exception):
        eng.pileup_cells(ext.scbam, {"not": "granges"}, cbs, str(tmp_path), param=fp)
    with pytest.raises(Exception):
        eng.pileup_cells(ext.scbam, gr5(), list(range(10)), str(tmp_path), param=fp)


def test_ranges(eng, ext, cbs, tmp_path):  # :65-69
    sce = eng.pileup_cells(ext.scbam, gr5(), cbs, str(tmp_path),
                           param=FilterParam(library_type="fr-second-strand", min_depth=0))
    assert list(sce.site_pos) == [579, 589, 601, 625, 645]
    assert sce.site_strand == ["-"] * 5


@pytest.mark.parametrize("kw,nrow", [
    (dict(min_variant_reads=1), 4), (dict(min_depth=1), 4), (dict(min_variant_reads=10), 3), (dict(min_depth=231), 3),
    (dict(min_depth=231, min_variant_reads=8), 2)])
def test_depth_filters(eng, ext, cbs, tmp_path, kw, nrow):  # :71-109
    sce = eng.pileup_cells(ext.scbam, gr5(), cbs, str(tmp_path), param=FilterParam(library_type="fr-second-strand", **kw))
    assert sce.shape[0] == nrow


def test_multiple_bams_as_cells(eng, ext, tmp_path):  # :112-151
    letters = list("ABCDEFGHIJ")
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    sce = eng.pileup_cells([ext.scbam] * 10, gr5(), letters, str(tmp_path), cb_tag=None, umi_tag=None, param=fp)
    assert sce.shape == (5, 10) and sce.barcodes == letters
    assert list(sce.site_pos) == [579, 589, 601, 625, 645]
    fp = FilterParam(library_type="fr-second-strand", min_depth=1)
    sce = eng.pileup_cells([ext.scbam] * 10, gr5(), letters, str(tmp_path), cb_tag=None, umi_tag=None, param=fp)
    assert sce.shape == (4, 10)
    d = sce.dense("nRef")
    assert np.all(d == d[:, :1])
    with pytest.raises(Exception):
        eng.pileup_cells([ext.scbam] * 10, gr5(), letters[:2], str(tmp_path), cb_tag=None, umi_tag=None, param=fp)


def test_output_files_and_read_sparray(eng, ext, cbs, tmp_path):  # :153-190
    from raer_b200.api import read_sparray

    sce = eng.pileup_cells(ext.scbam, gr5(), cbs, str(tmp_path), param=FilterParam(min_depth=0))
    fns = [str(tmp_path / f) for f in ("counts.mtx.gz", "sites.txt.gz", "barcodes.txt.gz")]
    for f in fns:
        assert open(f, "rb").read(2) == b"\x1f\x8b"
    sp = read_sparray(*fns, site_format="coordinate")
    assert sp.shape == sce.shape and sp.rownames() == sce.rownames()
    assert np.array_equal(sp.dense("nRef"), sce.dense("nRef")) and np.array_equal(sp.dense("nAlt"), sce.dense("nAlt"))
    sp = read_sparray(*fns, site_format="index")
    assert sp.index == [1, 2, 3, 4, 5]


def test_default_param_dims(eng, ext, cbs, tmp_path):  # :192-203
    sce = eng.pileup_cells(ext.scbam, gr5(), cbs, str(tmp_path))
    assert sce.shape == (4, 556)


def test_custom_indexes(eng, ext, tmp_path):  # :205-223
    import shutil

    b2 = str(tmp_path / "idx2")
    shutil.copy(ext.scbam + ".bai", b2)
    sce = eng.pileup_cells([ext.scbam] * 2, gr5(), ["A", "B"], str(tmp_path), cb_tag=None, umi_tag=None,
                           param=FilterParam(min_depth=0), indexes=[b2, b2])
    assert list(sce.site_pos) == [579, 589, 601, 625, 645]


def test_multiple_alleles_per_site(eng, ext, cbs, tmp_path):  # :249-260
    sites = Sites(["2"] * 6, [261, 261, 262, 262, 279, 279], strand=["+", "-"] * 3,
                  ref=["G", "C", "C", "G", "G", "C"], alt=["T", "A", "T", "A", "T", "A"])
    sce = eng.pileup_cells(ext.scbam, sites, cbs, str(tmp_path),
                           param=FilterParam(library_type="fr-second-strand", min_depth=0))
    assert sce.shape[0] == 6
    assert np.all(sce.row_sums("nAlt") > 0) and np.all(sce.row_sums("nRef") > 0)
    assert len(set(sce.rownames())) == 6


def test_readme_three_sites_three_barcodes(eng, ext, tmp_path):  # README.md:116-175
    sites = Sites(["2"] * 3, [579, 625, 589], strand=["-"] * 3, ref=["A"] * 3, alt=["G"] * 3)
    cbs3 = ["CACCAAACAACAACAA-1", "TATTCCACACCCTCTA-1", "GACCTTCAGTTGTAAG-1"]
    sce = eng.pileup_cells(ext.scbam, sites, cbs3, str(tmp_path),
                           param=FilterParam(only_keep_variants=True, library_type="fr-first-strand", min_depth=2))
    # rows follow the input order of `sites` (579, 625, 589); README prints them sorted by rownames
    order = {int(p): i for i, p in enumerate(sce.site_pos)}
    nref, nalt = sce.dense("nRef"), sce.dense("nAlt")
    assert nref[order[579]].tolist() == [0, 0, 1] and nalt[order[579]].tolist() == [1, 1, 1]
    assert nref[order[589]].tolist() == [1, 1, 2] and nalt[order[589]].tolist() == [0, 0, 0]
    assert nref[order[625]].tolist() == [0, 0, 0] and nalt[order[625]].tolist() == [1, 1, 1]


MOUSE_HDR = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:2\tLN:1115", "@SQ\tSN:6\tLN:400", "@SQ\tSN:11\tLN:493",
             "@SQ\tSN:8\tLN:396"]
ALN = ("A00836:50:HKF7VDSXX:4:1361:12337:2503\t16\t2\t247\t255\t14M315N77M\t*\t0\t0\t"
       "TTATTTATTTATTTTTTTGAGACAGGGTTTCTCTCTGTAGCCCTGGCTGTCCTGGAACTCACTCTGTAGACCAGGCTGGCCTCGAACTCAG\t"
       + "F" * 91 +
       "\tCB:Z:AGGGAGTAGGCTATCT-1\tUB:Z:CTATGTTCGTGG\tRE:A:E\tNH:i:1\tHI:i:1\tnM:i:7\tCR:Z:AGGGAGTAGGCTATCT\t"
       "UR:Z:CTATGTTCGTGG\tAS:i:67\tCY:Z:FFFFFFFFFFFFFFFF\tUY:Z:FFFFFFFFFFFF\txf:i:0")


def _change(aln, field, pos, new):
    rec = aln.split("\t")
    s = rec[field]
    rec[field] = s[:pos - 1] + new + s[pos:]
    return "\t".join(rec)


def test_umi_consensus(eng, tmp_path):  # :287-361
    ref_aln = _change(ALN, 9, 84, "A")
    ol = Sites(["2"] * 3, [20, 645, 646], strand=["+"] * 3, ref=["A"] * 3, alt=["G"] * 3)
    cb = ["AGGGAGTAGGCTATCT-1"]
    out = str(tmp_path / "o")
    fp = FilterParam(library_type="fr-first-strand", min_depth=0)

    def run(alns, **kw):
        bam = bamtools.sam_to_bam(str(tmp_path / "t.bam"), MOUSE_HDR, alns)
        sce = eng.pileup_cells(bam, ol, cb, out, **kw)
        return sce.dense("nRef")[:, 0].tolist(), sce.dense("nAlt")[:, 0].tolist()

    assert run([ref_aln, ALN], param=fp) == ([0, 0, 1], [0, 1, 0])
    assert run([ref_aln, ref_aln, ALN], param=fp) == ([0, 1, 1], [0, 0, 0])
    low = _change(ref_aln, 10, 84, chr(10 + 33))
    assert run([low, low, ALN], param=fp) == ([0, 0, 1], [0, 1, 0])
    fp2 = FilterParam(library_type="fr-first-strand", min_base_quality=10, min_depth=0)
    assert run([low, low, ALN], param=fp2, umi_tag=None) == ([0, 2, 3], [0, 1, 0])
    other = _change(ALN, 9, 84, "C")
    assert run([other, other, ALN], param=fp2) == ([0, 0, 1], [0, 0, 0])


def test_rownames_not_duplicated_bulk_then_cells(eng, ext, cbs, tmp_path):  # :226-247
    import oracle_backend

    bulkfp = FilterParam(min_mapq=255, min_variant_reads=1, min_allelic_freq=0.01, only_keep_variants=True,
                         report_multiallelic=False, library_type="fr-second-strand",
                         bam_flags=scan_bam_flag(isSecondaryAlignment=False, isSupplementaryAlignment=False,
                                                 isNotPassingQualityControls=False))
    rse = oracle_backend.pileup_sites(ext.scbam, ext.mouse_fa, chroms="2", param=bulkfp)
    assert len(rse) > 0
    sites = Sites(list(rse.seqnames), list(rse.pos), strand=list(rse.strand), ref=list(rse.REF),
                  alt=list(rse.assay("ALT")[:, 0]))
    sce = eng.pileup_cells(ext.scbam, sites, cbs, str(tmp_path),
                           param=FilterParam(library_type="fr-second-strand", min_depth=0))
    assert len(set(sce.rownames())) == len(sce.rownames())

"""Known-answer tests of the bulk pileup engine.

Every expectation is taken from the reference's own suite (tests/testthat/test_pileup_sites.R) or
its README (README.md:62-109); line numbers are cited per test. The same assertions run against
the CPU oracle (``engine == "oracle"``; pins the oracle, no GPU needed) and against the CUDA
product through the C-ABI (``engine == "cuda"``; marked gpu).
"""
import os

import numpy as np
import pytest

import bamtools
from raer_b200 import FilterParam, scan_bam_flag
from raer_b200.glue import Sites


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def plp(request):
    if request.param == "oracle":
        import oracle_backend

        return oracle_backend.pileup_sites
    from raer_b200 import api

    return api.pileup_sites


def vec(res, assay, col=0):
    return res.assay(assay)[:, col]


def test_pileup_works(plp, ext):  # test_pileup_sites.R:15-20
    res = plp(ext.bam1, ext.fasta, Sites.from_bed(ext.bed))
    assert len(res.REF) == 182
    assert len(res.assays) == 7


def test_filtering_for_variants(plp, ext):  # :22-38
    sites = Sites.from_bed(ext.bed)
    res = plp(ext.bam1, ext.fasta, sites)
    vars_ = res.subset(vec(res, "ALT") != "-")
    res_all = plp(ext.bam1, ext.fasta, sites, param=FilterParam(only_keep_variants=True))
    assert len(res_all) == 2
    assert vars_.identical(res_all)
    assert list(res_all.seqnames) == ["SSR3", "SPCS3"] and list(res_all.pos) == [228, 227]
    assert list(res_all.strand) == ["-", "+"] and list(vec(res_all, "ALT")) == ["G", "A"]
    assert vec(res_all, "nAlt")[1] == 2
    res_2 = plp(ext.bam1, ext.fasta, sites, param=FilterParam(min_variant_reads=2))
    assert len(res_2) == 1
    assert vars_.subset(vec(vars_, "nAlt") >= 2).identical(res_2)


def test_nbam_pileup(plp, ext):  # :40-98
    sites = Sites.from_bed(ext.bed)
    res = plp([ext.bam1, ext.bam1], ext.fasta, sites,
              param=FilterParam(library_type=["fr-first-strand", "fr-first-strand"]))
    for k in res.assays:
        assert np.array_equal(res.assays[k][:, 0], res.assays[k][:, 1])
    assert (res.nrow, res.ncol) == (182, 2)
    res = plp([ext.bam1, ext.bam2], ext.fasta, sites)
    assert (res.nrow, res.ncol) == (182, 2)
    assert not np.array_equal(res.assays["nRef"][:, 0], res.assays["nRef"][:, 1])
    res = plp([ext.bam1, ext.bam1], ext.fasta, sites, param=FilterParam(library_type=["fr-first-strand", "unstranded"]))
    assert (res.nrow, res.ncol) == (214, 2)
    res = plp([ext.bam1, ext.bam2], ext.fasta, sites,
              param=FilterParam(library_type="fr-first-strand", only_keep_variants=True))
    assert np.all((res.assays["ALT"] != "-").sum(axis=1) > 0)
    res = plp([ext.bam1] * 4, ext.fasta, sites, param=FilterParam(library_type="fr-first-strand"))
    assert res.ncol == 4
    res = plp([ext.bam1, ext.bam2], ext.fasta, sites,
              param=FilterParam(library_type="fr-first-strand", only_keep_variants=[True, False]))
    assert np.all(res.assays["ALT"][:, 0] != "-")
    res = plp([ext.bam1, ext.bam2], ext.fasta, sites,
              param=FilterParam(library_type="fr-first-strand", only_keep_variants=[False, True]))
    assert np.all(res.assays["ALT"][:, 1] != "-")


def test_regional_query(plp, ext):  # :100-111
    res = plp(ext.bam1, ext.fasta, region="SSR3:203-205")
    assert len(res) == 3 and list(res.pos) == [203, 204, 205]
    with pytest.raises(Exception):
        plp(ext.bam1, ext.fasta, region="chr1")
    res = plp(ext.bam1, ext.fasta, chroms="SSR3")
    assert len(res) == 529


def test_incorrect_region_and_missing_files(plp, ext):  # :113-121
    with pytest.raises(Exception):
        plp(ext.bam1, ext.fasta, region="chrHello")
    with pytest.raises(Exception):
        plp("hello.bam", ext.fasta)
    with pytest.raises(Exception):
        plp(ext.bam1, "hello.fasta")


def test_library_types(plp, ext):  # :123-144
    res = plp(ext.bam1, ext.fasta, param=FilterParam(library_type="fr-first-strand"))
    assert (res.strand == "+").sum() == 619 and (res.strand == "-").sum() == 1047
    assert np.all(res.strand[res.seqnames == "DHFR"] == "-")
    assert np.all(res.strand[res.seqnames == "SPCS3"] == "+")
    res = plp(ext.bam1, ext.fasta, param=FilterParam(library_type="fr-second-strand"))
    assert (res.strand == "+").sum() == 1047 and (res.strand == "-").sum() == 619
    assert np.all(res.strand[res.seqnames == "DHFR"] == "+")
    assert np.all(res.strand[res.seqnames == "SPCS3"] == "-")
    res = plp(ext.bam1, ext.fasta, param=FilterParam(library_type="unstranded"))
    assert np.all(res.strand == "+")


@pytest.mark.parametrize("lt", ["fr-first-strand", "fr-second-strand", "unstranded"])
def test_ref_alt_identities(plp, ext, lt):  # :173-201
    res = plp(ext.bam1, ext.fasta, param=FilterParam(library_type=lt))
    bases = "ATCG"
    mat = {b: vec(res, "n" + b) for b in bases}
    nref = np.array([mat[r][i] for i, r in enumerate(res.REF)])
    nalt = np.array([sum(mat[b][i] for b in bases if b != r) for i, r in enumerate(res.REF)])
    assert np.array_equal(nref, vec(res, "nRef"))
    assert np.array_equal(nalt, vec(res, "nAlt"))


def test_mapq_filter_equals_prefiltered_bam(plp, ext, tmp_path):  # :205-216
    bout = str(tmp_path / "mq.bam")
    bamtools.filter_bam(ext.bam1, bout, lambda r: r.mapq >= 255)
    a = plp(ext.bam1, ext.fasta, param=FilterParam(min_mapq=255))
    b = plp(bout, ext.fasta)
    b.sample_names = a.sample_names
    assert a.identical(b)


@pytest.mark.parametrize("minus", [False, True])
def test_flag_filter_equals_prefiltered_bam(plp, ext, tmp_path, minus):  # :218-240
    sites = Sites.from_bed(ext.bed)
    bout = str(tmp_path / "fl.bam")
    bamtools.filter_bam(ext.bam1, bout, lambda r: bool(r.flag & 16) == minus)
    a = plp(ext.bam1, ext.fasta, sites, param=FilterParam(bam_flags=scan_bam_flag(isMinusStrand=minus)))
    b = plp(bout, ext.fasta, sites)
    b.sample_names = a.sample_names
    assert a.identical(b)


def test_read_trimming(plp, ext):  # :244-267
    def run(region, **kw):
        return vec(plp(ext.bam1, ext.fasta, region=region, param=FilterParam(**kw)), "nRef")

    r = "SSR3:440-450"
    a, b, d, e = run(r), run(r, trim_5p=6), run(r, trim_5p=6, trim_3p=6), run(r, trim_3p=6)
    ex1 = [0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0]
    ex2 = [1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    assert list(a - b) == ex1 and list(b - d) == ex2 and list(a - e) == ex2
    r = "SSR3:128-130"
    a, b, d, e = run(r), run(r, trim_5p=6), run(r, trim_5p=6, trim_3p=6), run(r, trim_3p=6)
    assert list(a - b) == [3, 4, 2] and list(b - d) == [3, 2, 0] and list(a - e) == [3, 2, 0]


def test_fractional_trimming(plp, ext):  # :269-296
    def run(**kw):
        return plp(ext.bam1, ext.fasta, region="SSR3:440-450", param=FilterParam(min_base_quality=0, **kw))

    a, b = vec(run(), "nRef"), vec(run(ftrim_5p=0.10), "nRef")
    d, e = vec(run(ftrim_5p=0.10, ftrim_3p=0.10), "nRef"), vec(run(ftrim_3p=0.10), "nRef")
    ex1 = [0, 0, 1, 1, 1, 1, 1, 1, 2, 1, 1]
    ex2 = [1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0]
    assert list(a - b) == ex1 and list(b - d) == ex2 and list(a - e) == ex2
    assert len(run(ftrim_5p=0.99, ftrim_3p=0.99)) == 0
    with pytest.raises(ValueError):
        run(ftrim_3p=1.10)
    with pytest.raises(ValueError):
        run(ftrim_5p=-0.10)


def _hp_positions(fasta, n=6):
    pos = set()
    name, seq = None, []
    seqs = {}
    for line in open(fasta):
        if line.startswith(">"):
            if name:
                seqs[name] = "".join(seq)
            name, seq = line[1:].split()[0], []
        else:
            seq.append(line.strip())
    seqs[name] = "".join(seq)
    for nm, s in seqs.items():
        i = s.find("A" * n)
        while i >= 0:
            for k in range(n):
                pos.add((nm, i + k + 1))
            i = s.find("A" * n, i + 1)
    return pos


def test_homopolymer_filter(plp, ext):  # :298-308 (findOverlaps counts hits per (match, row) pair)
    a = plp(ext.bam1, ext.fasta, param=FilterParam(homopolymer_len=6))
    b = plp(ext.bam1, ext.fasta)
    assert not a.identical(b)
    hp = _hp_positions(ext.fasta)
    assert sum((s, int(p)) in hp for s, p in zip(a.seqnames, a.pos)) == 0
    # vmatchPattern returns overlapping matches; count (match,row) overlaps like findOverlaps
    seqs = {}
    name = None
    for line in open(ext.fasta):
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = ""
        else:
            seqs[name] += line.strip()
    n_hits = 0
    for nm, s in seqs.items():
        starts = [i + 1 for i in range(len(s) - 5) if s[i:i + 6] == "AAAAAA"]
        for st in starts:
            n_hits += int(((b.seqnames == nm) & (b.pos >= st) & (b.pos <= st + 5)).sum())
    assert n_hits == 120


def test_splice_dist(plp, ext):  # :310-329
    def variant_sites(**kw):
        r = plp(ext.bam1, ext.fasta, param=FilterParam(**kw))
        r = r.subset(vec(r, "ALT") != "-")
        return {(s, int(p)) for s, p in zip(r.seqnames, r.pos)}

    near = {("SPCS3", 227), ("SPCS3", 347), ("SPCS3", 348)}
    assert near <= variant_sites()
    assert variant_sites(splice_dist=5) & near == {("SPCS3", 227)}


def test_indel_dist(plp, ext):  # :332-365
    a = plp(ext.bam1, ext.fasta, param=FilterParam(min_base_quality=1, only_keep_variants=True))
    b = plp(ext.bam1, ext.fasta, param=FilterParam(min_base_quality=1, indel_dist=5, only_keep_variants=True))
    ka = {(s, int(p)): int(n) for s, p, n in zip(a.seqnames, a.pos, vec(a, "nAlt"))}
    kb = {(s, int(p)): int(n) for s, p, n in zip(b.seqnames, b.pos, vec(b, "nAlt"))}
    assert ("SSR3", 387) in ka and ("SSR3", 387) not in kb
    assert ka[("SSR3", 388)] - kb[("SSR3", 388)] == 1
    assert ka[("SSR3", 388)] == 7


@pytest.mark.parametrize("mm,expected", [((0, 0), 2), ((1, 1), 0), ((0, 10), 2), ((0, 1), 0), ((8, 0), 2), ((1, 0), 0)])
def test_mismatch_filter(plp, ext, mm, expected):  # :367-429
    res = plp(ext.bam1, ext.fasta, region="SSR3:244-247",
              param=FilterParam(min_base_quality=10, only_keep_variants=True, max_mismatch_type=mm))
    assert len(res) == expected


def test_serial_equals_per_chromosome(ext):  # :432-444 (BiocParallel branch = one .Call per contig)
    import oracle_backend

    rse = oracle_backend.pileup_sites(ext.bam1, ext.fasta)
    rse_mc = oracle_backend.pileup_sites(ext.bam1, ext.fasta, parallel_chroms=True)
    assert rse.identical(rse_mc)


def test_limiting_chromosomes(plp, ext):  # :447-465
    rse = plp(ext.bam1, ext.fasta)
    rse2 = plp(ext.bam1, ext.fasta, chroms=["SSR3", "SPCS3", "DHFR"])
    assert rse.identical(rse2)
    rse2 = plp(ext.bam1, ext.fasta, chroms="SSR3")
    assert rse2.identical(rse.subset(rse.seqnames == "SSR3"))


def test_read_bqual(plp, ext):  # :469-509
    res_no = plp(ext.bam1, ext.fasta, param=FilterParam(min_base_quality=0, library_type="unstranded"))
    _, refs, recs = bamtools.read_bam(ext.bam1)
    names = [n for n, _ in refs]
    bad = []
    for r in recs:
        if r.tid < 0 or r.flag & 4:
            continue
        q = r.qual
        if sum(1 for x in q if x < 20) / len(q) >= 0.25:
            bad.append(r)
    res = plp(ext.bam1, ext.fasta,
              param=FilterParam(read_bqual=(0.25, 20.0), min_base_quality=0, library_type="unstranded"))
    cov = {}
    for r in bad:
        p = r.pos
        for op, l in r.cigar:
            if op in (0, 7, 8):
                for k in range(l):
                    cov[(names[r.tid], p + k + 1)] = cov.get((names[r.tid], p + k + 1), 0) + 1
                p += l
            elif op in (2, 3):
                p += l
    tot_no = {(s, int(p)): int(a + b) for s, p, a, b in zip(res_no.seqnames, res_no.pos, vec(res_no, "nRef"), vec(res_no, "nAlt"))}
    tot = {(s, int(p)): int(a + b) for s, p, a, b in zip(res.seqnames, res.pos, vec(res, "nRef"), vec(res, "nAlt"))}
    assert len(bad) > 0
    # the R test compares rows present in both results that overlap the bad reads
    n_checked = 0
    for k, c in cov.items():
        if k in tot_no and k in tot:
            assert tot_no[k] - tot[k] == c
            n_checked += 1
    assert n_checked > 0


@pytest.mark.parametrize("region,overhang,nref", [
    ("SPCS3:220-220", 0, 7), ("SPCS3:220-220", 52, 6), ("SPCS3:350-350", 0, 5), ("SPCS3:350-350", 79, 4),
    ("SPCS3:350-350", 101, None)])
def test_splice_overhang(plp, ext, region, overhang, nref):  # :511-559
    res = plp(ext.bam1, ext.fasta, region=region,
              param=FilterParam(library_type="unstranded", min_splice_overhang=overhang, min_base_quality=0))
    if nref is None:
        assert len(res) == 0
    else:
        assert len(res) == 1 and int(vec(res, "nRef")[0]) == nref


def test_chroms_missing_from_fasta(plp, ext):  # :561-564
    with pytest.raises(Exception):
        with pytest.warns(UserWarning):
            plp(ext.bam1, ext.mouse_fa, Sites.from_bed(ext.bed))


def test_multiallelics(plp, ext):  # :566-602
    r = "DHFR:299-299"
    res_no = plp(ext.bam2, ext.fasta, region=r)
    assert res_no.assay("ALT")[0, 0] == "T,G"
    assert len(plp(ext.bam2, ext.fasta, region=r, param=FilterParam(report_multiallelic=False))) == 0
    n_og = len(plp(ext.bam2, ext.fasta))
    assert len(plp(ext.bam2, ext.fasta, param=FilterParam(report_multiallelic=False))) == n_og - 1
    res = plp(ext.bam2, ext.fasta, region=r, param=FilterParam(min_allelic_freq=0.05))
    assert len(res) == 1 and res.assay("ALT")[0, 0] == "G"
    for k in res.assays:
        if k != "ALT":
            assert np.array_equal(res.assays[k], res_no.assays[k])
    assert len(plp(ext.bam2, ext.fasta, param=FilterParam(min_allelic_freq=0.05))) == n_og
    res = plp(ext.bam2, ext.fasta, region=r, param=FilterParam(min_allelic_freq=0.05, report_multiallelic=False))
    assert len(res) == 1 and res.assay("ALT")[0, 0] == "G"
    res = plp(ext.bam2, ext.fasta, region=r, param=FilterParam(min_allelic_freq=0.0001))
    assert len(res) == 1 and res.assay("ALT")[0, 0] == "T,G"


def test_custom_indexes(plp, ext, tmp_path):  # :621-635
    import shutil

    b1, b2 = str(tmp_path / "i1"), str(tmp_path / "i2")
    shutil.copy(ext.bam1 + ".bai", b1)
    shutil.copy(ext.bam2 + ".bai", b2)
    res = plp([ext.bam1, ext.bam2], ext.fasta, Sites.from_bed(ext.bed), indexes=[b1, b2])
    assert len(res.REF) == 182 and res.ncol == 2


def test_readme_two_bam_default(plp, ext):  # README.md:62-93
    res = plp([ext.bam1, ext.bam2], ext.fasta)
    assert len(res) == 1695
    assert list(res.seqnames[:4]) == ["SSR3"] * 4 and list(res.pos[:4]) == [1, 2, 3, 4]
    assert list(res.strand[:4]) == ["-"] * 4
    assert res.assays["nRef"][:4].tolist() == [[13, 12], [14, 12], [14, 12], [15, 12]]
    assert res.assays["nAlt"][:4].sum() == 0


def test_readme_variant_filter(plp, ext):  # README.md:100-109
    res = plp([ext.bam1, ext.bam2], ext.fasta,
              param=FilterParam(only_keep_variants=True, library_type="fr-first-strand", min_depth=2))
    assert len(res) == 74


def test_stats_columns_present(plp, ext):  # :604-609
    res = plp(ext.bam1, ext.fasta, Sites.from_bed(ext.bed))
    assert res.rpbz.shape == res.vdb.shape == res.sor.shape == (182,)
    has_var = vec(res, "ALT") != "-"
    assert np.all(res.sor[~has_var] == 0.0) and np.all(res.rpbz[~has_var] == 0.0)
    assert np.all(res.sor[has_var] != 0.0)

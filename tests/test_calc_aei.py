"""calc_AEI() (reference R/calc_AEI.R:68-323) on the bundled BAMs: the assertions of the reference's
tests/testthat/test_aei.R:22-68 on the oracle (CPU), and CUDA == oracle for every one of the 12 indices
and their per-file (alt, ref) sums (BASELINE.json config C4 as a parity case)."""
import math

import pytest

import bamtools  # noqa: F401  (tests/ on sys.path)
import oracle_backend
from raer_b200 import FilterParam
from raer_b200.glue import Sites


def _whole_contigs(fasta):
    names, lens = [], []
    for line in open(fasta + ".fai"):
        f = line.split("\t")
        names.append(f[0])
        lens.append(int(f[1]))
    return Sites(names, [1] * len(names), lens)


def _mock_snps(fasta, chrom="SPCS3"):
    """every A of one contig, as test_aei.R:16 builds it"""
    seq, cur = [], None
    for line in open(fasta):
        if line.startswith(">"):
            cur = line[1:].split()[0]
        elif cur == chrom:
            seq.append(line.strip())
    s = "".join(seq)
    pos = [i + 1 for i, c in enumerate(s) if c == "A"]
    return Sites([chrom] * len(pos), pos)


def test_calc_aei_reference_assertions_on_the_oracle(ext):  # test_aei.R:22-68
    bams = [ext.bam1, ext.bam2]  # ko, wt
    alu = _whole_contigs(ext.fasta)
    fp = FilterParam(library_type="fr-first-strand")
    aei = oracle_backend.calc_AEI(bams, ext.fasta, alu, param=fp)
    assert len(aei) == 2 and len(aei["AEI"]) == 2
    ko, wt = (aei["AEI"][k] for k in aei["AEI"])
    assert len(ko) == 12 and len(wt) == 12
    assert wt["A_G"] > ko["A_G"]
    # the index is 100 * alt / (alt + ref) of the per-file sums
    for name, idx in aei["AEI"].items():
        for pair, v in idx.items():
            alt, ref = aei["AEI_per_chrom"][name][pair]
            assert (math.isnan(v) and alt + ref == 0) or v == 100.0 * alt / (alt + ref)
    # SNP mask: with every A of SPCS3 masked the A_* sums shrink, the other pairs are untouched
    snp = oracle_backend.calc_AEI(bams, ext.fasta, alu, snp_db=_mock_snps(ext.fasta), param=fp)
    for name in aei["AEI"]:
        a0, r0 = aei["AEI_per_chrom"][name]["A_G"]
        a1, r1 = snp["AEI_per_chrom"][name]["A_G"]
        assert r1 < r0 and a1 <= a0
        assert snp["AEI_per_chrom"][name]["C_T"] == aei["AEI_per_chrom"][name]["C_T"]
    # unstranded data needs gene annotation
    with pytest.raises(ValueError):
        oracle_backend.calc_AEI(bams, ext.fasta, alu, param=FilterParam(library_type="unstranded"))
    genes = _whole_contigs(ext.fasta)
    genes.strand = ["-", "+", "-"]
    us = oracle_backend.calc_AEI(bams, ext.fasta, alu, genes, param=FilterParam(library_type="unstranded"))
    for idx in us["AEI"].values():
        assert idx["A_G"] > idx["T_C"]


@pytest.mark.gpu
@pytest.mark.parametrize("lt", ["fr-first-strand", "unstranded"])
def test_calc_aei_cuda_equals_oracle(ext, lt):
    from raer_b200 import api

    bams = [ext.bam1, ext.bam2]
    alu = _whole_contigs(ext.fasta)
    genes = None
    if lt == "unstranded":
        genes = _whole_contigs(ext.fasta)
        genes.strand = ["-", "+", "-"]
    fp = FilterParam(library_type=lt)
    want = oracle_backend.calc_AEI(bams, ext.fasta, alu, genes, snp_db=_mock_snps(ext.fasta, "DHFR"), param=fp)
    got = api.calc_AEI(bams, ext.fasta, alu, genes, snp_db=_mock_snps(ext.fasta, "DHFR"), param=fp)
    assert got["AEI_per_chrom"] == want["AEI_per_chrom"]
    for name in want["AEI"]:
        for pair, v in want["AEI"][name].items():
            g = got["AEI"][name][pair]
            assert (math.isnan(v) and math.isnan(g)) or g == v


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(library_type="fr-first-strand"),
    dict(library_type="fr-second-strand", min_base_quality=10, trim_5p=3, min_mapq=3),
])
def test_fused_device_reduction_equals_row_based_calc_aei(tmp_path, kw):
    """The device-side 4 x 4 reduction (raer_pileup_args.aei_mode) against calc_AEI computed from the
    rows, on synthetic BAMs with interval sites and a SNP list."""
    import numpy as np

    from raer_b200 import _cabi, api, synth

    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=2, n_contigs=3, contig_len=60_000, pairs_per_contig=8000, seed=31)
    rng = np.random.default_rng(3)
    names, starts, ends = [], [], []
    for c in ds["names"][:2]:  # the third contig has no interval: it must not contribute
        for s in sorted(rng.choice(59_000, size=40, replace=False)):
            names.append(c); starts.append(int(s) + 1); ends.append(int(s) + int(rng.integers(100, 400)))
    alu = Sites(names, starts, ends)
    snp_pos = sorted(set(int(x) for x in rng.choice(60_000, size=3000, replace=False)))
    snps = Sites([ds["names"][0]] * len(snp_pos), [p + 1 for p in snp_pos])
    fp = FilterParam(**kw)
    fused = api.calc_AEI(ds["bams"], ds["fasta"], alu, snp_db=snps, param=fp)
    rows = api._calc_aei_impl(_cabi.pileup, ds["bams"], ds["fasta"], alu, None, snps, fp)
    want = oracle_backend.calc_AEI(ds["bams"], ds["fasta"], alu, snp_db=snps, param=fp)
    assert fused["AEI_per_chrom"] == rows["AEI_per_chrom"] == want["AEI_per_chrom"]
    assert sum(a + r for v in fused["AEI_per_chrom"].values() for a, r in v.values()) > 10_000


@pytest.mark.gpu
def test_fused_reduction_for_unstranded_libraries(tmp_path):
    """correct_strand() as interval arithmetic in front of the device reduction: genes of both strands, overlapping
    each other in places (those sites are dropped), sites outside every gene (dropped), a SNP list"""
    import numpy as np

    from raer_b200 import _cabi, api, synth

    ds = synth.make_bulk_dataset(str(tmp_path), n_bams=1, n_contigs=2, contig_len=60_000, pairs_per_contig=8000, seed=37)
    rng = np.random.default_rng(5)
    names, starts, ends = [], [], []
    for c in ds["names"]:
        for s in sorted(rng.choice(59_000, size=50, replace=False)):
            names.append(c); starts.append(int(s) + 1); ends.append(int(s) + int(rng.integers(100, 400)))
    alu = Sites(names, starts, ends)
    gn, gs, ge, gst = [], [], [], []
    for c in ds["names"]:
        for k in range(12):
            a = int(rng.integers(1, 55_000))
            gn.append(c); gs.append(a); ge.append(a + int(rng.integers(2_000, 9_000))); gst.append("+-"[k % 2])
    genes = Sites(gn, gs, ge, strand=gst)
    snp_pos = sorted(set(int(x) for x in rng.choice(60_000, size=2000, replace=False)))
    snps = Sites([ds["names"][0]] * len(snp_pos), [p + 1 for p in snp_pos])
    fp = FilterParam(library_type="unstranded")
    fused = api.calc_AEI(ds["bams"], ds["fasta"], alu, genes, snp_db=snps, param=fp)
    rows = api._calc_aei_impl(_cabi.pileup, ds["bams"], ds["fasta"], alu, genes, snps, fp)
    want = oracle_backend.calc_AEI(ds["bams"], ds["fasta"], alu, genes, snp_db=snps, param=fp)
    assert fused["AEI_per_chrom"] == rows["AEI_per_chrom"] == want["AEI_per_chrom"]
    assert sum(a + r for v in fused["AEI_per_chrom"].values() for a, r in v.values()) > 5_000


def test_gene_strand_split_of_site_intervals():
    """correct_strand() as interval arithmetic (api._sites_by_gene_strand) against the per-position definition: a site
    takes the strand of the genes over it when they agree, and is dropped otherwise"""
    import random

    from raer_b200 import api

    rng = random.Random(4)
    for _ in range(60):
        sites = Sites(["c"] * 4 + ["d"], [rng.randrange(1, 80) for _ in range(5)], None)
        sites.end = [s + rng.randrange(0, 15) for s in sites.start]
        gn, gs, ge, gst = [], [], [], []
        for _ in range(rng.randrange(0, 7)):
            a = rng.randrange(1, 90)
            gn.append(rng.choice("cd")); gs.append(a); ge.append(a + rng.randrange(0, 30)); gst.append(rng.choice("+-*"))
        genes = Sites(gn, gs, ge, strand=gst)
        plus, minus = api._sites_by_gene_strand(sites, genes)
        want = {"+": set(), "-": set()}
        for sn, a, b in zip(sites.seqnames, sites.start, sites.end):
            for p in range(a, b + 1):
                st = {g for n, x, y, g in zip(gn, gs, ge, gst) if n == sn and x <= p <= y}
                if len(st) == 1 and "*" not in st:
                    want[st.pop()].add((sn, p))
        for got, key in ((plus, "+"), (minus, "-")):
            have = [(sn, p) for sn, a, b in zip(got.seqnames, got.start, got.end) for p in range(a, b + 1)]
            assert len(have) == len(set(have)) and set(have) == want[key]

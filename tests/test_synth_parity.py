"""Parity of the CUDA bulk path against the CPU oracle on seeded synthetic BAMs of the C2 shape
 and on adversarial hand-built records. Integer outputs must be bit-exact, rpbz/vdb/sor
within 1e-6 relative (BASELINE.json north_star), Inf matching exactly.
"""
import os

import numpy as np
import pytest

import bamtools
import oracle_backend
from raer_b200 import FilterParam, scan_bam_flag
from raer_b200.glue import Sites

STATS_RTOL = 1e-6


@pytest.fixture(scope="module")
def syn(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("syn")
    return synth.make_bulk_dataset(str(d), n_bams=2, n_contigs=3, contig_len=120_000, pairs_per_contig=6000, seed=1001)


def _cuda():
    from raer_b200 import api

    return api.pileup_sites


def _assert_same(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert np.array_equal(a.seqnames, b.seqnames) and np.array_equal(a.pos, b.pos)
    assert np.array_equal(a.strand, b.strand) and np.array_equal(a.REF, b.REF)
    for k in list(a.assays) + list(a.extra):
        x, y = a.assay(k), b.assay(k)
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)[:5]
            raise AssertionError(f"{k} differs at rows {bad.tolist()}: {x[bad[:, 0]].tolist()} vs {y[bad[:, 0]].tolist()} "
                                 f"pos {a.pos[bad[:, 0]].tolist()} {a.seqnames[bad[:, 0]].tolist()}")
    assert a.identical(b, stats_rtol=STATS_RTOL), "statistics differ"


CONFIGS = [
    dict(),
    dict(library_type="fr-first-strand", only_keep_variants=True),  # the C2 call
    dict(library_type="unstranded", min_base_quality=0),
    dict(library_type="fr-second-strand", min_mapq=[255, 3]),
    dict(remove_overlaps=False, min_depth=5),
    dict(trim_5p=5, trim_3p=7, indel_dist=4, splice_dist=6),
    dict(ftrim_5p=0.1, ftrim_3p=0.15, min_splice_overhang=30),
    dict(homopolymer_len=4, min_variant_reads=2, only_keep_variants=[False, True]),
    dict(min_allelic_freq=0.2, report_multiallelic=False),
    dict(min_depth=0, min_base_quality=30),
    dict(read_bqual=(0.2, 20.0), max_depth=40),
    dict(bam_flags=scan_bam_flag(isMinusStrand=True)),
    dict(min_base_quality=0, only_keep_variants=True, min_allelic_freq=0.01),
    dict(library_type="fr-first-strand", max_mismatch_type=(1, 2)),
    dict(library_type="unstranded", max_mismatch_type=(2, 3), trim_5p=3, min_base_quality=10),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CONFIGS, ids=[str(i) for i in range(len(CONFIGS))])
def test_two_bam_parity(syn, kw):
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(syn["bams"], syn["fasta"], param=fp)
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    _assert_same(got, want)


# Every configuration runs k_pileup_fast by default; RAER_KERNEL=generic sends the
# same call through k_pileup_tiles, so both kernels are held to the oracle.
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [CONFIGS[i] for i in (0, 1, 2, 3, 4, 5, 6, 9, 12, 13, 14)],
                         ids=[str(i) for i in (0, 1, 2, 3, 4, 5, 6, 9, 12, 13, 14)])
def test_two_bam_parity_generic_kernel(syn, kw, monkeypatch):
    monkeypatch.setenv("RAER_KERNEL", "generic")
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(syn["bams"], syn["fasta"], param=fp)
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    _assert_same(got, want)


# Calls with only the mapq / base-quality filters run k_pileup_lane by default; RAER_KERNEL=fast keeps them on
# k_pileup_fast<false>, so that kernel stays held to the oracle as well.
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [CONFIGS[i] for i in (0, 1, 2, 3, 4, 7, 8, 9, 10, 12)],
                         ids=[str(i) for i in (0, 1, 2, 3, 4, 7, 8, 9, 10, 12)])
def test_two_bam_parity_item_kernel(syn, kw, monkeypatch):
    monkeypatch.setenv("RAER_KERNEL", "fast")
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(syn["bams"], syn["fasta"], param=fp)
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    _assert_same(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("fast_w", [32, 160, 512])
def test_fast_kernel_tile_width_does_not_change_results(syn, fast_w, monkeypatch):
    fp = FilterParam(library_type="fr-first-strand", min_depth=0)
    want = oracle_backend.pileup_sites(syn["bams"], syn["fasta"], param=fp)
    monkeypatch.setenv("RAER_FAST_W", str(fast_w))
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    _assert_same(got, want)


@pytest.fixture(scope="module")
def syn_deep(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("syn_deep")
    # 200x per BAM like the C2 workload: tiles with more flagged sites than one second-pass round holds, dense
    # mate overlaps, many reads per tile list chunk
    return synth.make_bulk_dataset(str(d), n_bams=2, n_contigs=2, contig_len=40_000, pairs_per_contig=40_000, seed=4711)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(library_type="fr-first-strand", only_keep_variants=True),  # the C2 call
    dict(library_type="fr-second-strand", min_base_quality=30, min_depth=0),
    dict(library_type="unstranded", trim_5p=6, trim_3p=6, splice_dist=5, indel_dist=4, min_splice_overhang=12,
         only_keep_variants=True),
])
def test_c2_depth_parity(syn_deep, kw):
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(syn_deep["bams"], syn_deep["fasta"], param=fp)
    got = _cuda()(syn_deep["bams"], syn_deep["fasta"], param=fp)
    _assert_same(got, want)
    assert len(want) > 5000


@pytest.fixture(scope="module")
def syn5(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("syn5")
    return synth.make_bulk_dataset(str(d), n_bams=5, n_contigs=2, contig_len=60_000, pairs_per_contig=2500, seed=77)


# More files = narrower tiles (the counters of all files share the CTA's shared memory) and more (file, strand)
# counter blocks: 5 BAMs with mixed per-file options.
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["default", "generic"])
@pytest.mark.parametrize("kw", [
    dict(library_type="fr-first-strand", only_keep_variants=[True, False, False, True, False], min_mapq=[0, 255, 3, 0, 0]),
    dict(library_type=["unstranded", "fr-first-strand", "fr-second-strand", "unstranded", "fr-first-strand"], min_depth=3,
         trim_5p=4, min_allelic_freq=0.05),
])
def test_five_bam_parity(syn5, kw, kernel, monkeypatch):
    if kernel == "generic":
        monkeypatch.setenv("RAER_KERNEL", "generic")
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(syn5["bams"], syn5["fasta"], param=fp)
    got = _cuda()(syn5["bams"], syn5["fasta"], param=fp)
    _assert_same(got, want)
    assert len(want) > 300


@pytest.mark.gpu
@pytest.mark.parametrize("batch_reads", [40, 333, 5000])
def test_batch_seams_do_not_change_results(syn, batch_reads, monkeypatch):
    fp = FilterParam(library_type="fr-first-strand")
    want = oracle_backend.pileup_sites(syn["bams"], syn["fasta"], param=fp)
    monkeypatch.setenv("RAER_BATCH_READS", str(batch_reads))
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    _assert_same(got, want)


@pytest.mark.gpu
def test_region_and_sites(syn):
    fp = FilterParam(only_keep_variants=True)
    for region in ("chr2:5000-90000", "chr3"):
        _assert_same(_cuda()(syn["bams"], syn["fasta"], region=region, param=fp),
                     oracle_backend.pileup_sites(syn["bams"], syn["fasta"], region=region, param=fp))
    rng = np.random.default_rng(5)
    pos = np.sort(rng.choice(np.arange(1, 119_000), size=4000, replace=False))
    sites = Sites(["chr1"] * 2000 + ["chr3"] * 2000, pos.tolist())
    fp = FilterParam(min_depth=2)
    _assert_same(_cuda()(syn["bams"], syn["fasta"], sites, param=fp),
                 oracle_backend.pileup_sites(syn["bams"], syn["fasta"], sites, param=fp))


@pytest.mark.gpu
def test_single_bam_and_linearity(syn):
    """Columns of a two-file call equal the single-file calls wherever a row exists in both."""
    fp = FilterParam(library_type="fr-first-strand")
    both = _cuda()(syn["bams"], syn["fasta"], param=fp)
    for f in range(2):
        one = _cuda()(syn["bams"][f], syn["fasta"], param=fp)
        _assert_same(one, oracle_backend.pileup_sites(syn["bams"][f], syn["fasta"], param=fp))
        key_b = {(s, int(p), st): i for i, (s, p, st) in enumerate(zip(both.seqnames, both.pos, both.strand))}
        idx = np.array([key_b[(s, int(p), st)] for s, p, st in zip(one.seqnames, one.pos, one.strand)])
        for k in ("nRef", "nAlt", "nA", "nT", "nC", "nG"):
            assert np.array_equal(one.assay(k)[:, 0], both.assay(k)[idx, f])


@pytest.mark.gpu
def test_checksum_of_counts_equals_aligned_bases(syn):
    """With every filter open, every aligned non-N base of every kept read is counted exactly once."""
    fp = FilterParam(min_base_quality=0, remove_overlaps=False, min_depth=1, library_type="unstranded",
                     bam_flags=scan_bam_flag())
    got = _cuda()(syn["bams"], syn["fasta"], param=fp)
    total = int(got.assay("nRef").sum() + got.assay("nAlt").sum())
    assert total == syn["aligned_bases"]


HDR = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:c1\tLN:400", "@SQ\tSN:c2\tLN:300"]


def _write_ref(path, seqs):
    with open(path, "w") as fa, open(path + ".fai", "w") as fai:
        off = 0
        for nm, s in seqs.items():
            h = f">{nm}\n"
            fa.write(h + s + "\n")
            off += len(h)
            fai.write(f"{nm}\t{len(s)}\t{off}\t{len(s)}\t{len(s) + 1}\n")
            off += len(s) + 1


def _sam(q, flag, rname, pos, mapq, cigar, mpos, tlen, seq, qual, extra=""):
    s = f"{q}\t{flag}\t{rname}\t{pos}\t{mapq}\t{cigar}\t=\t{mpos}\t{tlen}\t{seq}\t{qual}"
    return s + ("\t" + extra if extra else "")


@pytest.fixture(scope="module")
def adversarial(tmp_path_factory):
    """Every CIGAR op, clips, overlapping mates with equal / unequal qualities and a deletion inside the
    overlap, IUPAC read bases, N and lower case in the reference, duplicate qnames, depth > max_depth."""
    d = tmp_path_factory.mktemp("adv")
    rng = np.random.default_rng(11)
    c1 = "".join(rng.choice(list("ACGT"), size=400))
    c1 = c1[:50] + "NNNN" + c1[54:120].lower() + "AAAAAAAA" + c1[128:]
    c2 = "".join(rng.choice(list("ACGT"), size=300))
    fa = str(d / "adv.fa")
    _write_ref(fa, {"c1": c1, "c2": c2})
    up = c1.upper()

    def rd(start0, cigar):
        """read sequence following the reference through `cigar` from 0-based start"""
        import re

        out, p = "", start0
        for ln, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
            ln = int(ln)
            if op in "M=X":
                out += up[p:p + ln].replace("N", "A")
                p += ln
            elif op in "DN":
                p += ln
            elif op in "IS":
                out += "".join(rng.choice(list("ACGT"), size=ln))
        return out

    def mut(s, i, b):
        return s[:i] + b + s[i + 1:]

    recs = []
    q40 = lambda n: "I" * n  # noqa: E731
    # overlapping proper pairs: equal bases, mismatching bases with unequal and equal qualities
    for k, (p1, p2) in enumerate([(10, 40), (12, 50), (15, 15), (20, 95)]):
        s1, s2 = rd(p1, "60M"), rd(p2, "60M")
        if k == 0:
            s2 = mut(s2, 5, "A" if s2[5] != "A" else "C")  # mismatch inside the overlap, equal quality
        if k == 1:
            s2 = mut(s2, 8, "G" if s2[8] != "G" else "T")
        qa, qb = q40(60), (q40(60) if k != 1 else "5" * 60)
        recs.append((p1, _sam(f"ov{k}", 99, "c1", p1 + 1, 255, "60M", p2 + 1, p2 + 60 - p1, s1, qa, "MD:Z:60")))
        recs.append((p2, _sam(f"ov{k}", 147, "c1", p2 + 1, 255, "60M", p1 + 1, -(p2 + 60 - p1), s2, qb, "MD:Z:60")))
    # deletion inside the mate overlap, insertion, soft + hard clips, spliced read, = and X ops, padding
    s1, s2 = rd(130, "30M2D30M"), rd(150, "60M")
    recs.append((130, _sam("del", 163, "c1", 131, 255, "30M2D30M", 151, 80, s1, q40(60))))
    recs.append((150, _sam("del", 83, "c1", 151, 255, "60M", 131, -80, s2, q40(60))))
    recs.append((135, _sam("ins", 0, "c1", 136, 60, "20M3I20M5S", 0, 0, rd(135, "20M3I20M5S"), q40(48))))
    recs.append((140, _sam("clip", 16, "c1", 141, 60, "3H4S30M2S", 0, 0, rd(140, "4S30M2S"), "".join(chr(33 + q) for q in rng.integers(0, 41, 36)))))
    recs.append((160, _sam("spl", 0, "c1", 161, 3, "25M60N20M1D15M", 0, 0, rd(160, "25M60N20M1D15M"), q40(60))))
    recs.append((170, _sam("eqx", 16, "c1", 171, 60, "10=1X10=2P5M", 0, 0, mut(rd(170, "26M"), 10, "T"), q40(26))))
    # IUPAC and N read bases
    s = rd(200, "40M")
    s = mut(mut(mut(s, 3, "N"), 7, "R"), 20, "M")
    recs.append((200, _sam("iupac", 0, "c1", 201, 60, "40M", 0, 0, s, q40(40))))
    # duplicate qname triple (secondary flagged) + QC fail + duplicate + supplementary records
    for i, fl in enumerate([0, 256, 512, 1024, 2048, 4]):
        recs.append((210 + i, _sam("dupname", fl, "c1", 211 + i, 60, "30M", 0, 0, rd(210 + i, "30M"), q40(30))))
    # paired but not proper (always dropped), mate unmapped
    recs.append((220, _sam("notproper", 65, "c1", 221, 60, "30M", 300, 0, rd(220, "30M"), q40(30))))
    recs.append((222, _sam("mateunmapped", 73, "c1", 223, 60, "30M", 223, 0, rd(222, "30M"), q40(30))))
    # a tall stack of identical starts to trip max_depth, with variants at one position
    for i in range(60):
        s = rd(260, "50M")
        if i % 7 == 0:
            s = mut(s, 25, "C" if s[25] != "C" else "G")
        if i % 11 == 0:
            s = mut(s, 25, "G" if s[25] not in "G" else "T")
        if i % 13 == 0:
            s = mut(s, 25, "T" if s[25] != "T" else "A")
        recs.append((260 + (i // 20), _sam(f"st{i}", 16 if i % 2 else 0, "c1", 261 + (i // 20), 60, "50M", 0, 0, s[: 50], q40(50))))
    recs.sort(key=lambda x: x[0])
    lines = [r[1] for r in recs]
    # second contig: plain reads
    for i in range(30):
        p = 5 + 7 * i
        s = "".join(c2[p:p + 40])
        lines.append(_sam(f"c2r{i}", 16 if i % 3 == 0 else 0, "c2", p + 1, 60, "40M", 0, 0, s, q40(40)))
    bam = bamtools.sam_to_bam(str(d / "adv.bam"), HDR, lines)
    return bam, fa


ADV_CONFIGS = [
    dict(library_type="unstranded", min_base_quality=0),
    dict(library_type="fr-first-strand"),
    dict(library_type="fr-second-strand", min_base_quality=10, only_keep_variants=True),
    dict(library_type="unstranded", max_depth=25, min_base_quality=0),
    dict(library_type="unstranded", remove_overlaps=False, min_depth=0),
    dict(library_type="unstranded", bam_flags=scan_bam_flag(), min_base_quality=0),
    dict(library_type="unstranded", indel_dist=3, splice_dist=4, trim_5p=2, min_splice_overhang=22),
    dict(library_type="unstranded", min_allelic_freq=0.1),
    dict(library_type="unstranded", homopolymer_len=5, min_mapq=10),
]


ADV_CONFIGS += [
    dict(library_type="fr-first-strand", ftrim_5p=0.15, ftrim_3p=0.2, min_base_quality=0),
    dict(library_type="unstranded", trim_3p=6, indel_dist=5, min_base_quality=0, min_depth=0),
    dict(library_type="fr-second-strand", splice_dist=8, min_splice_overhang=60, only_keep_variants=True),
    dict(library_type="fr-first-strand", max_mismatch_type=(1, 1), min_base_quality=0),
    dict(library_type="unstranded", max_mismatch_type=(2, 2)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["default", "generic"])
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("kw", ADV_CONFIGS, ids=[str(i) for i in range(len(ADV_CONFIGS))])
def test_adversarial_parity(adversarial, kw, mode, kernel, monkeypatch):
    bam, fa = adversarial
    from raer_b200 import _cabi, api

    if kernel == "generic":
        monkeypatch.setenv("RAER_KERNEL", "generic")
    fp = FilterParam(**kw)
    monkeypatch.setattr(oracle_backend, "OVERLAP_MODE", mode)
    want = oracle_backend.pileup_sites(bam, fa, param=fp)
    got = api._pileup_sites_impl(lambda pc: _cabi.pileup(pc, overlap_mode=mode), bam, fa, param=fp)
    _assert_same(got, want)


def test_adversarial_oracle_runs(adversarial):
    """CPU-only sanity: the oracle digests every adversarial record and both htslib overlap rules give the
    same base counts while being allowed to differ in strand statistics."""
    bam, fa = adversarial
    fp = FilterParam(library_type="unstranded", min_base_quality=0)
    res = {}
    for mode in (1, 2):
        oracle_backend.OVERLAP_MODE = mode
        res[mode] = oracle_backend.pileup_sites(bam, fa, param=fp)
    oracle_backend.OVERLAP_MODE = 1
    assert len(res[1]) > 300
    assert len(res[1]) == len(res[2])


# ---- khash growth at a site that writes no row (src/plp.c:487: every variant read is a kh_put) ------------------
@pytest.fixture(scope="module")
def growth(tmp_path_factory):
    """Two BAMs. Site 30 (REF A) shows C, G, T in both; in the second file a fourth variant read follows, which
    resizes that file's variant table from 4 to 8 buckets. min_variant_reads = 5 keeps the site out of the result.
    Site 80 (REF C) shows T and A six times: the printed ALT order tells the table size (T,A with 4 buckets,
    A,T with 8)."""
    d = tmp_path_factory.mktemp("growth")
    rng = np.random.default_rng(5)
    ref = list("".join(rng.choice(list("ACGT"), size=400)))
    ref[30], ref[80] = "A", "C"
    ref = "".join(ref)
    fa = str(d / "g.fa")
    _write_ref(fa, {"c1": ref, "c2": "ACGT" * 75})
    bams = []
    for fno, first in enumerate([["C", "G", "T"], ["C", "G", "T", "C"]]):
        lines = []
        for i in range(10):
            s = ref[10:50]
            if i < len(first):
                s = s[:20] + first[i] + s[21:]
            lines.append(_sam(f"a{i}", 0, "c1", 11, 60, "40M", 0, 0, s, "I" * 40))
        for i in range(10):
            s = ref[60:100]
            if i < 6:
                s = s[:20] + ("T" if i % 2 == 0 else "A") + s[21:]
            lines.append(_sam(f"b{i}", 0, "c1", 61, 60, "40M", 0, 0, s, "I" * 40))
        bams.append(bamtools.sam_to_bam(str(d / f"g{fno}.bam"), HDR, lines))
    return bams, fa


def _growth_alt(res):
    i = int(np.flatnonzero(res.pos == 81)[0])
    return [str(x) for x in res.assay("ALT")[i]]


def test_khash_growth_oracle(growth):
    bams, fa = growth
    res = oracle_backend.pileup_sites(bams, fa, param=FilterParam(library_type="unstranded", min_variant_reads=5))
    assert list(res.pos) == [81]
    assert _growth_alt(res) == ["T,A", "A,T"]


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["default", "generic"])
def test_khash_growth_at_unwritten_site(growth, kernel, monkeypatch):
    bams, fa = growth
    if kernel == "generic":
        monkeypatch.setenv("RAER_KERNEL", "generic")
    for kw in (dict(min_variant_reads=5), dict(min_variant_reads=5, only_keep_variants=True), dict(min_depth=11)):
        fp = FilterParam(library_type="unstranded", **kw)
        want = oracle_backend.pileup_sites(bams, fa, param=fp)
        got = _cuda()(bams, fa, param=fp)
        _assert_same(got, want)
    fp = FilterParam(library_type="unstranded", min_variant_reads=5)
    assert _growth_alt(_cuda()(bams, fa, param=fp)) == ["T,A", "A,T"]


# ---- C3 shape: eight BAMs, a list of A (+) / T (-) reference positions as width-1 sites, min_depth = 2,
# as one eight-file call and as eight one-file calls
@pytest.fixture(scope="module")
def syn8(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("syn8")
    ds = synth.make_bulk_dataset(str(d), n_bams=8, n_contigs=2, contig_len=50_000, pairs_per_contig=3000, seed=3001)
    rng = np.random.default_rng(3000)
    names, pos = [], []
    for t, nm in enumerate(ds["names"]):
        ref = synth.make_reference(ds["lens"], 3001)[t].numpy()
        cand = np.flatnonzero((ref == 0) | (ref == 3))
        pick = np.sort(rng.choice(cand, size=5000, replace=False))
        names += [nm] * len(pick)
        pos += (pick + 1).tolist()
    ds["sites"] = Sites(names, pos)
    return ds


@pytest.mark.gpu
def test_known_sites_eight_bams(syn8):
    fp = FilterParam(min_depth=2)
    want = oracle_backend.pileup_sites(syn8["bams"], syn8["fasta"], syn8["sites"], param=fp)
    got = _cuda()(syn8["bams"], syn8["fasta"], syn8["sites"], param=fp)
    _assert_same(got, want)
    assert 1000 < len(want) <= 2 * len(syn8["sites"])
    assert set(np.unique(want.REF)) <= {"A", "T"}
    # one file per call (one BAM per GPU in the C3 layout): the same counts as that file's column wherever both have the row
    one = _cuda()(syn8["bams"][3], syn8["fasta"], syn8["sites"], param=fp)
    _assert_same(one, oracle_backend.pileup_sites(syn8["bams"][3], syn8["fasta"], syn8["sites"], param=fp))
    key = {(s, p, st): i for i, (s, p, st) in enumerate(zip(want.seqnames, want.pos, want.strand))}
    idx = np.array([key[(s, p, st)] for s, p, st in zip(one.seqnames, one.pos, one.strand)])
    for k in ("nRef", "nAlt", "nA", "nT", "nC", "nG"):
        assert np.array_equal(one.assay(k)[:, 0], want.assay(k)[idx, 3])


# raer_pileup streams the contigs of a call through several workers (BAI seeks) when the host has the cores for it;
# RAER_SERIAL_INGEST keeps the single-stream path. Both must give the oracle's result, rows in contig order.
@pytest.mark.gpu
@pytest.mark.parametrize("serial", [False, True])
@pytest.mark.parametrize("batch_reads", [None, 700])
def test_serial_and_contig_parallel_ingest(syn5, serial, batch_reads, monkeypatch):
    if serial:
        monkeypatch.setenv("RAER_SERIAL_INGEST", "1")
    if batch_reads:
        monkeypatch.setenv("RAER_BATCH_READS", str(batch_reads))
    fp = FilterParam(library_type="fr-first-strand", min_variant_reads=2)
    want = oracle_backend.pileup_sites(syn5["bams"][:3], syn5["fasta"], param=fp)
    got = _cuda()(syn5["bams"][:3], syn5["fasta"], param=fp)
    _assert_same(got, want)
    assert list(got.seqnames) == sorted(got.seqnames, key=lambda s: int(s[3:]))


@pytest.mark.gpu
def test_unusable_index_falls_back_to_the_single_stream(syn5, tmp_path):
    """Without `region` the reference never opens the index; a call whose .bai is unreadable must still work."""
    import shutil

    bams = []
    for b in syn5["bams"][:2]:
        dst = str(tmp_path / os.path.basename(b))  # sample names come from the file names
        shutil.copy(b, dst)
        with open(dst + ".bai", "wb") as fh:
            fh.write(b"not an index")
        bams.append(dst)
    fp = FilterParam(library_type="fr-first-strand")
    want = oracle_backend.pileup_sites(syn5["bams"][:2], syn5["fasta"], param=fp)
    got = _cuda()(bams, syn5["fasta"], param=fp)
    _assert_same(got, want)

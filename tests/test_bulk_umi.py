"""pileup_sites(umi_tag = ...): only the first read per UMI, site, strand and file is processed
(reference src/plp.c:622-646, consulted at :738 before the filter verdict is used). On the device the host
turns that into one number per read (hdr.aux) plus a list of exceptions (umi_holes). Compared with the
oracle's literal per-site string set on hand-built duplicates (identical copies, shifted copies, a duplicate
whose first copy shows an N or a deletion / ref-skip at the site, reads without the tag, both strands) and on
the bundled 10x BAM."""
import numpy as np
import pytest

import bamtools
import oracle_backend
from raer_b200 import FilterParam
from test_synth_parity import HDR, _assert_same, _sam, _write_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def umi_bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("umi")
    rng = np.random.default_rng(5)
    c1 = "".join(rng.choice(list("ACGT"), size=400))
    c2 = "".join(rng.choice(list("ACGT"), size=300))
    fa = str(d / "umi.fa")
    _write_ref(fa, {"c1": c1, "c2": c2})

    def rd(start0, cigar):
        import re

        out, p = "", start0
        for ln, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
            ln = int(ln)
            if op in "M=X":
                out += c1[p:p + ln]
                p += ln
            elif op in "DN":
                p += ln
            elif op in "IS":
                out += "".join(rng.choice(list("ACGT"), size=ln))
        return out

    def mut(s, i, b):
        return s[:i] + b + s[i + 1:]

    q = lambda n: "I" * n  # noqa: E731
    recs = []

    def add(name, flag, pos0, cigar, seq, umi, qual=None):
        recs.append((pos0, _sam(name, flag, "c1", pos0 + 1, 60, cigar, 0, 0, seq, qual or q(len(seq)),
                                f"UB:Z:{umi}" if umi else "")))

    # identical copies, a variant only in the later copies (must disappear), another UMI at the same place
    for i in range(4):
        s = rd(10, "50M")
        if i > 0:
            s = mut(s, 20, "A" if s[20] != "A" else "C")
        add(f"dupA{i}", 0, 10, "50M", s, "AAAA")
    add("otherA", 0, 10, "50M", mut(rd(10, "50M"), 30, "G" if c1[40] != "G" else "T"), "CCCC")
    # shifted copies: the later read is only shadowed up to the earlier read's end
    add("shift0", 0, 70, "40M", rd(70, "40M"), "GGGG")
    add("shift1", 0, 85, "40M", mut(rd(85, "40M"), 35, "T" if c1[120] != "T" else "A"), "GGGG")
    add("shift2", 0, 90, "60M", mut(mut(rd(90, "60M"), 3, "C" if c1[93] != "C" else "G"), 50, "A" if c1[140] != "A" else "C"), "GGGG")
    # the first copy shows N at two positions: the second copy is counted there (holes), and only there
    s0 = mut(mut(rd(160, "40M"), 5, "N"), 17, "N")
    add("n0", 0, 160, "40M", s0, "TTTT")
    add("n1", 0, 160, "40M", mut(mut(rd(160, "40M"), 5, "G" if c1[165] != "G" else "C"), 9, "A" if c1[169] != "A" else "T"), "TTTT")
    add("n2", 0, 162, "40M", rd(162, "40M"), "TTTT")
    # the first copy has a deletion and a ref-skip: it still takes the UMI there (its entry shows the next base)
    add("gap0", 0, 220, "15M3D10M20N15M", rd(220, "15M3D10M20N15M"), "ACAC")
    add("gap1", 0, 222, "60M", mut(rd(222, "60M"), 14, "T" if c1[236] != "T" else "G"), "ACAC")
    # the first copy fails its filters (low quality everywhere): it still takes the UMI
    add("lowq0", 0, 300, "30M", rd(300, "30M"), "GTGT", qual="#" * 30)
    add("lowq1", 0, 300, "30M", mut(rd(300, "30M"), 10, "A" if c1[310] != "A" else "C"), "GTGT")
    # reads without the tag are skipped everywhere; same UMI on the other strand is its own group
    add("notag", 0, 300, "30M", mut(rd(300, "30M"), 12, "G" if c1[312] != "G" else "T"), None)
    add("rev0", 16, 340, "40M", rd(340, "40M"), "AAAA")
    add("rev1", 16, 341, "40M", mut(rd(341, "40M"), 7, "C" if c1[348] != "C" else "A"), "AAAA")
    add("fwd_same_umi", 0, 342, "40M", mut(rd(342, "40M"), 9, "G" if c1[351] != "G" else "T"), "AAAA")
    recs.sort(key=lambda x: x[0])
    lines = [r[1] for r in recs]
    for i in range(20):  # second contig: a few duplicates too
        p = 5 + 9 * (i // 2)
        lines.append(_sam(f"c2r{i}", 0, "c2", p + 1, 60, "40M", 0, 0, c2[p:p + 40], q(40), f"UB:Z:U{i // 2}"))
    bam = bamtools.sam_to_bam(str(d / "umi.bam"), HDR, lines)
    return bam, fa


@pytest.mark.parametrize("kw", [
    dict(library_type="fr-first-strand", min_depth=1),
    dict(library_type="unstranded", min_depth=1, min_base_quality=0),
    dict(library_type="fr-second-strand", min_depth=0, only_keep_variants=True),
    dict(library_type="fr-first-strand", min_depth=1, trim_5p=3, indel_dist=2),
])
def test_bulk_umi_parity(umi_bam, kw):
    from raer_b200 import api

    bam, fa = umi_bam
    fp = FilterParam(**kw)
    want = oracle_backend.pileup_sites(bam, fa, param=fp, umi_tag="UB")
    got = api.pileup_sites(bam, fa, param=fp, umi_tag="UB")
    _assert_same(got, want)
    # the tag changes the answer on this input (otherwise the test pins nothing)
    plain = oracle_backend.pileup_sites(bam, fa, param=fp)
    assert not (len(plain) == len(want) and np.array_equal(plain.assay("nRef"), want.assay("nRef"))
                and np.array_equal(plain.assay("nAlt"), want.assay("nAlt")))


def test_bulk_umi_on_the_bundled_10x_bam(ext):
    from raer_b200 import api

    fp = FilterParam(library_type="fr-second-strand", min_depth=1)
    want = oracle_backend.pileup_sites(ext.scbam, ext.mouse_fa, param=fp, umi_tag="UB")
    got = api.pileup_sites(ext.scbam, ext.mouse_fa, param=fp, umi_tag="UB")
    _assert_same(got, want)
    assert len(want) > 0

"""pileup_cells on a synthetic 10x-shaped BAM of the C5 shape at test size: the CUDA engine through the
C-ABI against the CPU oracle, triplet for triplet (integer outputs, bit-exact)."""
import numpy as np
import pytest

import oracle_backend
from raer_b200 import FilterParam, scan_bam_flag
from raer_b200.glue import Sites


@pytest.fixture(scope="module")
def sc(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("sc_syn")
    ds = synth.make_sc_dataset(str(d))
    ds["gr"] = Sites(**ds["sites"])
    return ds


def _triplets(sce):
    if sce is None:
        return None
    o = np.lexsort((sce.col, sce.row))
    return (sce.shape, list(sce.barcodes), sce.row[o].tolist(), sce.col[o].tolist(), sce.nRef[o].tolist(),
            sce.nAlt[o].tolist(), sce.rownames())


CONFIGS = [
    # the C5 call (vignettes/raer.Rmd:155-159)
    (dict(library_type="fr-second-strand", min_mapq=255, min_variant_reads=1), dict()),
    (dict(library_type="fr-second-strand", min_depth=0), dict(umi_tag=None)),
    (dict(library_type="fr-second-strand", min_depth=3, min_variant_reads=2, min_base_quality=30), dict()),
    (dict(library_type="fr-first-strand", min_depth=0), dict()),
    (dict(library_type="unstranded", min_depth=0, trim_5p=6, trim_3p=4, indel_dist=3, splice_dist=5), dict()),
    (dict(library_type="fr-second-strand", min_depth=0, bam_flags=scan_bam_flag(isDuplicate=False)), dict(umi_tag=None)),
    (dict(library_type="fr-second-strand", min_depth=0, min_splice_overhang=20, homopolymer_len=4), dict()),
]


def test_sc_dataset_is_informative(sc, tmp_path):
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    sce = oracle_backend.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path), param=fp)
    assert sce.shape == (len(sc["gr"]), len(sc["barcodes"]))
    assert sce.nRef.sum() > 1000 and sce.nAlt.sum() > 200
    noumi = oracle_backend.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path), umi_tag=None, param=fp)
    assert noumi.nRef.sum() > 1.2 * sce.nRef.sum()  # UMI consensus collapses duplicates
    # the C5 call keeps the sites with a variant read only
    fp = FilterParam(library_type="fr-second-strand", min_mapq=255, min_variant_reads=1)
    c5 = oracle_backend.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path), param=fp)
    assert 100 < c5.shape[0] < len(sc["gr"])


@pytest.mark.gpu
@pytest.mark.parametrize("kw,ckw", CONFIGS, ids=[str(i) for i in range(len(CONFIGS))])
def test_sc_synth_parity(sc, kw, ckw, tmp_path):
    from raer_b200 import api

    fp = FilterParam(**kw)
    want = oracle_backend.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path / "o"), param=fp, **ckw)
    got = api.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path / "c"), param=fp, **ckw)
    assert _triplets(got) == _triplets(want)


@pytest.mark.gpu
def test_sc_synth_barcode_subset_and_chroms(sc, tmp_path):
    from raer_b200 import api

    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    sub = sc["barcodes"][::3]
    want = oracle_backend.pileup_cells(sc["bam"], sc["gr"], sub, str(tmp_path / "o"), chroms=["chr2"], param=fp)
    got = api.pileup_cells(sc["bam"], sc["gr"], sub, str(tmp_path / "c"), chroms=["chr2"], param=fp)
    assert _triplets(got) == _triplets(want)


@pytest.mark.gpu
@pytest.mark.parametrize("kw,ckw", CONFIGS[:3], ids=["0", "1", "2"])
def test_triplets_in_memory_equal_the_file_round_trip(sc, kw, ckw, tmp_path):
    """raer_scpileup_triplets: same cells as the three text files read back."""
    from raer_b200 import api

    fp = FilterParam(**kw)
    files = api.pileup_cells(sc["bam"], sc["gr"], sc["barcodes"], str(tmp_path / "c"), param=fp, **ckw)
    mem = api.pileup_cells_triplets(sc["bam"], sc["gr"], sc["barcodes"], param=fp, **ckw)
    assert _triplets(mem) == _triplets(files)


@pytest.mark.gpu
def test_triplets_on_the_bundled_10x_bam(ext, tmp_path):
    import bamtools
    from raer_b200 import api

    _, _, recs = bamtools.read_bam(ext.scbam)
    cbs = list(dict.fromkeys(r.tag_z("CB") for r in recs if r.tag_z("CB") is not None))
    gr = Sites(["2"] * 5, [579, 589, 601, 625, 645], strand=["-"] * 5, ref=["A", "A", "T", "A", "A"], alt=["G", "G", "C", "G", "G"])
    fp = FilterParam(library_type="fr-second-strand", min_depth=0)
    files = api.pileup_cells(ext.scbam, gr, cbs, str(tmp_path), param=fp)
    mem = api.pileup_cells_triplets(ext.scbam, gr, cbs, param=fp)
    assert _triplets(mem) == _triplets(files)
    assert list(zip(mem.row_sums("nRef"), mem.row_sums("nAlt"))) == [(103, 127), (273, 7), (3, 277), (149, 135), (0, 0)]

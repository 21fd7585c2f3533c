"""Site lists reach the reads through the BAI index (the reference queries the index per region,
src/plp.c:856-882): the ingest skips what lies between the intervals of the list, forward only. Same kept reads, same
batches, same rows as streaming the whole file; fewer bytes inflated. Host-only checks through raer_ingest_summary,
and the rows through the device under -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_backend
from raer_b200 import FilterParam, _cabi
from raer_b200.glue import Sites, prepare_pileup_call


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    from raer_b200 import synth

    d = str(tmp_path_factory.mktemp("jumps"))
    # two contigs of 6 Mb, about 25 MB of BGZF each: many 16 kb index windows, seeks that pay
    return synth.make_bulk_dataset(d, n_bams=1, n_contigs=2, contig_len=6_000_000, pairs_per_contig=150_000, seed=91)


def _sites(ds, n, seed):
    rng = np.random.default_rng(seed)
    sn, sp = [], []
    for nm, L in zip(ds["names"], ds["lens"]):
        pos = np.sort(rng.choice(np.arange(1000, L - 1000), size=n // len(ds["names"]), replace=False))
        sn += [nm] * pos.size
        sp += [int(p) for p in pos]
    return Sites(sn, sp)


def _summary(ds, sites):
    L = _cabi.lib()
    L.raer_ingest_summary.argtypes = [C.POINTER(_cabi.PileupArgs), C.c_int64, C.POINTER(C.c_int64)]
    pc = prepare_pileup_call(ds["bams"], ds["fasta"], sites)
    a = _cabi.PileupArgs()
    keep = [_cabi._strs(pc.bampaths), _cabi._strs(pc.indexes), _cabi._strs(list(pc.lst[0])), _cabi._ints(pc.lst[1]),
            _cabi._ints(pc.lst[2])]
    a.nbam, a.bampaths, a.indexes, a.fapath = pc.n, keep[0], keep[1], pc.fapath.encode()
    a.n_regions, a.reg_seqnames, a.reg_start, a.reg_end = len(pc.lst[0]), keep[2], keep[3], keep[4]
    a.int_args = (C.c_int32 * 14)(*pc.int_args)
    a.dbl_args = (C.c_double * 5)(*pc.dbl_args)
    a.lgl_args = (C.c_int32 * 2)(*[int(x) for x in pc.lgl_args])
    keep += [_cabi._ints(pc.libtype), _cabi._ints([int(x) for x in pc.only_keep_variants]), _cabi._ints(pc.min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3:]
    out = (C.c_int64 * 16)()
    rc = L.raer_ingest_summary(C.byref(a), 0, out)
    assert rc == 0, _cabi.last_error()
    return dict(zip(("reads", "pairs", "cigar", "sq", "batches", "aligned_bases", "records", "distinct", "inflated", "seeks"), out))


@pytest.mark.parametrize("n_sites", [10, 100, 1000, 100_000])
def test_jumps_keep_the_same_reads_and_inflate_less(big, n_sites, monkeypatch):
    sites = _sites(big, n_sites, 5)
    jump = _summary(big, sites)
    monkeypatch.setenv("RAER_SERIAL_INGEST", "1")
    stream = _summary(big, sites)
    for k in ("reads", "pairs", "cigar", "sq", "batches", "distinct"):
        assert jump[k] == stream[k], (k, jump, stream)
    assert stream["seeks"] == 0 and stream["reads"] > 0
    if n_sites <= 100:
        assert jump["seeks"] > 0
        assert jump["inflated"] < stream["inflated"] / (20 if n_sites == 10 else 2), (jump, stream)
        assert jump["records"] < stream["records"]
    else:
        # sites closer than a few 16 kb index windows: nothing to skip but the two ends of each contig; never more
        # work than streaming
        assert jump["inflated"] <= stream["inflated"] * 1.05


@pytest.mark.gpu
@pytest.mark.parametrize("n_sites", [10, 100, 100_000])
def test_rows_through_the_index_equal_streaming_and_the_oracle(big, n_sites, monkeypatch):
    from raer_b200 import api

    sites = _sites(big, n_sites, 6)
    fp = FilterParam(library_type="fr-first-strand", min_depth=2)
    got = api.pileup_sites(big["bams"], big["fasta"], sites=sites, param=fp)
    t_jump = dict(_cabi.LAST_TIMING)
    monkeypatch.setenv("RAER_SERIAL_INGEST", "1")
    ref = api.pileup_sites(big["bams"], big["fasta"], sites=sites, param=fp)
    t_stream = dict(_cabi.LAST_TIMING)
    assert got.identical(ref) and len(got) > 0
    if n_sites <= 100:
        assert t_jump["n_seeks"] > 0 and t_jump["n_inflated_bytes"] < t_stream["n_inflated_bytes"] / 2
        want = oracle_backend.pileup_sites(big["bams"], big["fasta"], sites=sites, param=fp)
        assert got.identical(want, stats_rtol=1e-6)

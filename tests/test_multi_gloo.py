"""world_size-2 tests of the multi-GPU host logic: work units (contigs, tiles of contigs, BAM files), ordered
concatenation on rank 0. CPU: gloo with the oracle as the per-rank .Call target. -m gpu: the same two ranks on one GPU
with the C-ABI (CUDA) call, against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _save(out, res):
    np.savez(out, pos=res.pos, seq=res.seqnames, strand=res.strand, nref=res.assay("nRef"), nalt=res.assay("nAlt"),
             alt=res.assay("ALT"), sor=res.sor, rpbz=res.rpbz, vdb=res.vdb)


def _worker(rank, world, port, bams, fasta, out, cuda=False, min_units=1, per_bam=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_backend
    from raer_b200 import FilterParam, multi

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if cuda:
        os.environ["LOCAL_RANK"] = "0"  # both ranks on the one GPU of the test box
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True)
    call = None if cuda else oracle_backend.oracle_pileup_call
    if per_bam:
        res = multi.pileup_sites_per_bam_distributed(bams, fasta, param=fp, _call=call)
        if rank == 0:
            for i, r in enumerate(res):
                _save(f"{out}.{i}.npz", r)
    else:
        res = multi.pileup_sites_distributed(bams, fasta, param=fp, _call=call, min_units_per_rank=min_units)
        if rank == 0:
            _save(out, res)
    dist.barrier()
    dist.destroy_process_group()


def _same(got, want, stats_rtol=None):
    assert np.array_equal(got["pos"], want.pos) and np.array_equal(got["seq"], want.seqnames)
    assert np.array_equal(got["strand"], want.strand)
    assert np.array_equal(got["nref"], want.assay("nRef")) and np.array_equal(got["nalt"], want.assay("nAlt"))
    assert np.array_equal(got["alt"], want.assay("ALT"))
    for k, w in (("sor", want.sor), ("rpbz", want.rpbz), ("vdb", want.vdb)):
        if stats_rtol is None:
            assert np.array_equal(got[k], w), k
        else:
            fin = np.isfinite(w)
            assert np.array_equal(fin, np.isfinite(got[k])) and np.array_equal(got[k][~fin], w[~fin]), k
            assert np.allclose(got[k][fin], w[fin], rtol=stats_rtol, atol=0), k


def test_sharded_equals_single_process(ext, tmp_path):
    import oracle_backend
    from raer_b200 import FilterParam

    out = str(tmp_path / "merged.npz")
    bams = [ext.bam1, ext.bam2]
    mp.spawn(_worker, args=(2, _free_port(), bams, ext.fasta, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True)
    want = oracle_backend.pileup_sites(bams, ext.fasta, param=FilterParam(library_type="fr-first-strand",
                                                                         only_keep_variants=True))
    _same(got, want)


@pytest.fixture(scope="module")
def syn_multi(tmp_path_factory):
    from raer_b200 import synth

    d = tmp_path_factory.mktemp("syn_multi")
    return synth.make_bulk_dataset(str(d), n_bams=2, n_contigs=3, contig_len=60_000, pairs_per_contig=3000, seed=77)


def _want(ds, bams=None):
    import oracle_backend
    from raer_b200 import FilterParam

    return oracle_backend.pileup_sites(bams or ds["bams"], ds["fasta"],
                                       param=FilterParam(library_type="fr-first-strand", only_keep_variants=True))


def test_tiles_of_contigs_equal_single_process(syn_multi, tmp_path):
    # 3 contigs on 2 ranks with 4 units per rank wanted: every contig is cut into 3 tiles, each one .Call with a region
    from raer_b200.multi import plan_units

    plan = plan_units(syn_multi["names"], dict(zip(syn_multi["names"], syn_multi["lens"])), 2, 4)
    assert sum(len(p) for p in plan) == 9 and all(u[2] is not None for p in plan for u in p)
    out = str(tmp_path / "tiles.npz")
    mp.spawn(_worker, args=(2, _free_port(), syn_multi["bams"], syn_multi["fasta"], out, False, 4), nprocs=2, join=True)
    _same(np.load(out, allow_pickle=True), _want(syn_multi))


def test_one_bam_per_rank(syn_multi, tmp_path):
    out = str(tmp_path / "perbam")
    mp.spawn(_worker, args=(2, _free_port(), syn_multi["bams"], syn_multi["fasta"], out, False, 2, True), nprocs=2, join=True)
    for i, bam in enumerate(syn_multi["bams"]):
        _same(np.load(f"{out}.{i}.npz", allow_pickle=True), _want(syn_multi, [bam]))


@pytest.mark.gpu
@pytest.mark.parametrize("min_units", [1, 4])
def test_sharded_cuda_equals_oracle(syn_multi, tmp_path, min_units):
    """the partitioned path with the CUDA call: whole contigs (min_units 1) and tiles with halo reads (min_units 4)"""
    out = str(tmp_path / "cuda.npz")
    mp.spawn(_worker, args=(2, _free_port(), syn_multi["bams"], syn_multi["fasta"], out, True, min_units), nprocs=2, join=True)
    _same(np.load(out, allow_pickle=True), _want(syn_multi), stats_rtol=1e-6)


@pytest.mark.gpu
def test_one_bam_per_rank_cuda(syn_multi, tmp_path):
    out = str(tmp_path / "perbam_cuda")
    mp.spawn(_worker, args=(2, _free_port(), syn_multi["bams"], syn_multi["fasta"], out, True, 2, True), nprocs=2, join=True)
    for i, bam in enumerate(syn_multi["bams"]):
        _same(np.load(f"{out}.{i}.npz", allow_pickle=True), _want(syn_multi, [bam]), stats_rtol=1e-6)


def test_unit_plan_cuts_contigs_when_there_are_too_few():
    from raer_b200.multi import plan_units, unit_region

    chroms = [f"c{i}" for i in range(20)]
    lens = {c: 5_000_000 for c in chroms}
    for world in (1, 2, 4):
        plan = plan_units(chroms, lens, world)
        assert sorted(u[1] for p in plan for u in p) == sorted(chroms) and all(u[2] is None for p in plan for u in p)
    plan = plan_units(chroms, lens, 8)  # 20 contigs < 2 x 8 ranks: two tiles per contig, 5 units per rank
    assert [len(p) for p in plan] == [5] * 8
    covered = {}
    for p in plan:
        for _, c, b, e in p:
            covered.setdefault(c, []).append((b, e))
    assert all(sorted(v) == [(0, 2_500_000), (2_500_000, 5_000_000)] for v in covered.values())
    assert unit_region("c1", 0, 2_500_000) == "c1:1-2500000" and unit_region("c1", None, None) == "c1"


def test_shard_plan_is_balanced_and_complete():
    from raer_b200.multi import shard_contigs

    chroms = [f"c{i}" for i in range(20)]
    lens = {c: 5_000_000 for c in chroms}
    for world in (1, 2, 4, 8):
        plan = shard_contigs(chroms, lens, world)
        assert sorted(sum(plan, [])) == sorted(chroms)
        sizes = [len(p) for p in plan]
        assert max(sizes) - min(sizes) <= 1
    lens = {c: (i + 1) * 1000 for i, c in enumerate(chroms)}
    plan = shard_contigs(chroms, lens, 4)
    loads = [sum(lens[c] for c in p) for p in plan]
    assert max(loads) / min(loads) < 1.1

"""world_size-2 gloo test of the multi-GPU host logic (contig sharding + ordered concatenation).
The per-rank .Call target is the CPU oracle here; on GPUs the same code path binds the C-ABI."""
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


def _worker(rank, world, port, bams, fasta, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_backend
    from raer_b200 import FilterParam, multi

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True)
    res = multi.pileup_sites_distributed(bams, fasta, param=fp, _call=oracle_backend.oracle_pileup_call)
    if rank == 0:
        np.savez(out, pos=res.pos, seq=res.seqnames, strand=res.strand, nref=res.assay("nRef"), nalt=res.assay("nAlt"),
                 alt=res.assay("ALT"), sor=res.sor)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process(ext, tmp_path):
    import oracle_backend
    from raer_b200 import FilterParam

    out = str(tmp_path / "merged.npz")
    bams = [ext.bam1, ext.bam2]
    mp.spawn(_worker, args=(2, _free_port(), bams, ext.fasta, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True)
    want = oracle_backend.pileup_sites(bams, ext.fasta, param=FilterParam(library_type="fr-first-strand",
                                                                         only_keep_variants=True))
    assert np.array_equal(got["pos"], want.pos) and np.array_equal(got["seq"], want.seqnames)
    assert np.array_equal(got["strand"], want.strand)
    assert np.array_equal(got["nref"], want.assay("nRef")) and np.array_equal(got["nalt"], want.assay("nAlt"))
    assert np.array_equal(got["alt"], want.assay("ALT"))
    assert np.array_equal(got["sor"], want.sor)


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

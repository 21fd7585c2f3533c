"""Multi-GPU execution of the bulk pileup: one process per GPU, work sharded by contig, no
collective on the data path; the merge is a host-side ordered concatenation (BASELINE.json
north_star). torch.distributed is only plumbing (rendezvous + gathering the per-rank
row blocks on rank 0): NCCL on GPUs, gloo in the CPU tests.

The reference parallelises the same way one level up: one .Call per chromosome through
BiocParallel, results rbind-ed in R (R/pileup.R:208-243).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

from . import glue
from .filter_param import FilterParam
from .results import PileupResult, rbind, result_from_columns


def shard_contigs(chroms: List[str], lengths: dict, world: int) -> List[List[str]]:
    """Greedy longest-first assignment of contigs to ranks (balances aligned bases when coverage is
    roughly uniform). Deterministic, so every rank computes the same plan without communication."""
    order = sorted(chroms, key=lambda c: (-lengths[c], chroms.index(c)))
    load = [0] * world
    plan: List[List[str]] = [[] for _ in range(world)]
    for c in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(c)
        load[r] += lengths[c]
    for p in plan:
        p.sort(key=chroms.index)
    return plan


def _pileup_sites_sharded_impl(call: Callable, rank: int, world: int, bamfiles, fasta, sites=None, chroms=None,
                               param: Optional[FilterParam] = None, umi_tag=None, indexes=None) -> Optional[PileupResult]:
    """This rank's contigs -> one .Call per contig (region = contig), rbind in contig order."""
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    todo, contigs, _ = glue.setup_valid_regions(bamfiles[0], chroms, None, fasta)
    plan = shard_contigs(todo, dict(contigs), world)
    parts = []
    for ctig in plan[rank]:
        pc = glue.prepare_pileup_call(bamfiles, fasta, sites, ctig, None, param, umi_tag, indexes)
        parts.append((todo.index(ctig), result_from_columns(call(pc), pc.sample_ids)))
    return parts


def merge_parts(all_parts) -> Optional[PileupResult]:
    """Ordered concatenation of (contig order, rows) blocks gathered from every rank."""
    flat = [p for parts in all_parts for p in parts]
    if not flat:
        return None
    flat.sort(key=lambda t: t[0])
    return rbind([p for _, p in flat])


def pileup_sites_distributed(bamfiles, fasta, sites=None, chroms=None, param: Optional[FilterParam] = None,
                             umi_tag=None, indexes=None, _call: Optional[Callable] = None) -> Optional[PileupResult]:
    """Call from every rank of an initialised torch.distributed group; rank 0 returns the merged result."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if _call is None:
        from . import _cabi

        local = int(os.environ.get("LOCAL_RANK", rank))
        _call = lambda pc: _cabi.pileup(pc, device=local)  # noqa: E731
    parts = _pileup_sites_sharded_impl(_call, rank, world, bamfiles, fasta, sites, chroms, param, umi_tag, indexes)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(parts, gathered, dst=0)
    return merge_parts(gathered) if rank == 0 else None

"""Multi-GPU execution of the bulk pileup: one process per GPU, work sharded by genomic region (whole contigs, or
tiles of a contig when there are fewer than two contigs per rank) and -- for independent calls -- by BAM file; no
collective on the data path; the merge is a host-side ordered concatenation (BASELINE.json north_star).
torch.distributed is only plumbing (rendezvous + gathering the per-rank row blocks on rank 0): NCCL on GPUs, gloo in
the CPU tests.

The reference parallelises the same way one level up: one .Call per chromosome through BiocParallel, results rbind-ed
in R (R/pileup.R:208-243); calc_AEI and Smart-seq2 runs fan out per BAM file (R/calc_AEI.R:176-236, R/sc-pileup.R:185-230).
A tile is one .Call with region = "contig:beg-end": the engine fetches the reads that overlap it through the index (the
halo of a tile) and writes the rows of its positions only (src/plp.c:918), so the tiles' rows concatenate to the
contig's rows. Three pieces of state do not cross a tile boundary, exactly as they do not cross a region boundary
in the reference: the max_depth buffer count, the khash size behind the ALT column's order (relevant only after a site
with four distinct variant keys), and the mate-overlap correction of a read whose earlier mate ends before the tile
(htslib's overlap pass can change qualities of the later mate beyond the earlier mate's end; a region query does not
fetch that mate, here or in the reference).
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


def plan_units(chroms: List[str], lengths: dict, world: int, min_units_per_rank: int = 1, min_balance: float = 0.9):
    """Work units (order key, contig, beg0, end0) per rank; beg0 / end0 are None for a whole contig. Every contig is
    cut into k equal tiles, k being the smallest number that gives every rank at least `min_units_per_rank` units and
    a longest-first assignment whose busiest rank carries at most 1 / min_balance of the mean load (20 equal contigs on
    8 GPUs: whole contigs would leave ranks with 3 and with 2, 83 %; two tiles per contig give 5 units each).
    Deterministic: every rank computes the same plan without communication."""
    n = len(chroms)

    def build(k):
        units = []
        for ci, c in enumerate(chroms):
            L = int(lengths[c])
            if k == 1 or L < 2 * k:
                units.append(((ci, 0), c, None, None, L))
                continue
            step = -(-L // k)
            for j in range(k):
                b, e = j * step, min(L, (j + 1) * step)
                if b < e:
                    units.append(((ci, j), c, b, e, e - b))
        order = sorted(range(len(units)), key=lambda i: (-units[i][4], i))
        load = [0] * world
        plan = [[] for _ in range(world)]
        for i in order:
            r = min(range(world), key=lambda q: (load[q], q))
            plan[r].append(units[i][:4])
            load[r] += units[i][4]
        for p in plan:
            p.sort(key=lambda u: u[0])
        return plan, load

    best = None
    for k in range(1, 65):
        plan, load = build(k)
        best = plan
        if world <= 1 or n == 0:
            break
        if n * k < min_units_per_rank * world:
            continue
        if max(load) > 0 and sum(load) / (world * max(load)) >= min_balance:
            break
    return best


def unit_region(contig: str, beg0, end0) -> str:
    """region string of a unit: 1-based inclusive, as hts_parse_reg reads it"""
    return contig if beg0 is None else f"{contig}:{beg0 + 1}-{end0}"


def _pileup_sites_tiled_impl(call: Callable, rank: int, world: int, bamfiles, fasta, sites=None, chroms=None,
                             param: Optional[FilterParam] = None, umi_tag=None, indexes=None, min_units_per_rank: int = 1):
    """This rank's units -> one .Call per unit (region = contig or contig:beg-end)."""
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    todo, contigs, _ = glue.setup_valid_regions(bamfiles[0], chroms, None, fasta)
    plan = plan_units(todo, dict(contigs), world, min_units_per_rank)
    parts = []
    for key, ctig, b, e in plan[rank]:
        pc = glue.prepare_pileup_call(bamfiles, fasta, sites, unit_region(ctig, b, e), None, param, umi_tag, indexes)
        parts.append((key, result_from_columns(call(pc), pc.sample_ids)))
    return parts


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


def _default_call():
    import torch.distributed as dist

    from . import _cabi

    local = int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    return lambda pc: _cabi.pileup(pc, device=local)


def pileup_sites_distributed(bamfiles, fasta, sites=None, chroms=None, param: Optional[FilterParam] = None,
                             umi_tag=None, indexes=None, _call: Optional[Callable] = None,
                             min_units_per_rank: int = 1) -> Optional[PileupResult]:
    """Call from every rank of an initialised torch.distributed group; rank 0 returns the merged result. Work is cut by
    contig, and into tiles of a contig when the contigs alone would leave ranks idle (plan_units)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if _call is None:
        _call = _default_call()
    parts = _pileup_sites_tiled_impl(_call, rank, world, bamfiles, fasta, sites, chroms, param, umi_tag, indexes,
                                     min_units_per_rank)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(parts, gathered, dst=0)
    return merge_parts(gathered) if rank == 0 else None


def pileup_sites_per_bam_distributed(bamfiles, fasta, sites=None, chroms=None, param: Optional[FilterParam] = None,
                                     umi_tag=None, indexes=None, _call: Optional[Callable] = None):
    """Independent one-BAM calls (the C3 shape "one BAM per GPU"; calc_AEI's loop over BAM files, R/calc_AEI.R:176-183):
    BAM i goes to rank i % world. Rank 0 returns the list of per-BAM results in the order of `bamfiles`. Inside ONE
    multi-BAM call the files cannot be split: the write decision and the pooled statistics couple them (src/plp.c:374-443)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if _call is None:
        _call = _default_call()
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    mine = []
    for i, bam in enumerate(bamfiles):
        if i % world != rank:
            continue
        ix = None if indexes is None else [indexes[i]]
        pc = glue.prepare_pileup_call([bam], fasta, sites, None, chroms, param, umi_tag, ix)
        mine.append((i, result_from_columns(_call(pc), pc.sample_ids)))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    flat = sorted((p for parts in gathered for p in parts), key=lambda t: t[0])
    return [r for _, r in flat]

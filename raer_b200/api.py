"""User-facing entry points: pileup_sites(), pileup_cells(), calc_AEI(), read_sparray().

Same names, argument meaning and error behaviour as the reference's R API (R/pileup.R:106,
R/sc-pileup.R:114, R/calc_AEI.R:68), in Python because no R toolchain exists in the build image.
Everything below the ``.Call`` line runs in libraer_b200.so (CUDA, sm_100a); there is no CPU
fallback: if the library or a GPU is missing these functions raise.

The glue is written once as ``_*_impl(call, ...)`` taking the ``.Call`` target as its first
argument so that tests can drive the CPU oracle through the identical glue; the public functions
bind it to the C-ABI only.
"""
from __future__ import annotations

import gzip
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import glue
from .filter_param import FilterParam
from .glue import Sites
from .results import PileupResult, SparseCounts, rbind, result_from_columns


# ----------------------------------------------------------------------------- pileup_sites
def _pileup_sites_impl(call: Callable, bamfiles, fasta, sites=None, region=None, chroms=None, param=None,
                       umi_tag=None, indexes=None, parallel_chroms: bool = False) -> PileupResult:
    pc = glue.prepare_pileup_call(bamfiles, fasta, sites, region, chroms, param, umi_tag, indexes)
    if not parallel_chroms:
        cols = call(pc)
        return result_from_columns(cols, pc.sample_ids)
    # BPPARAM branch (R/pileup.R:208-243): one .Call per contig with region = contig, results rbind-ed
    param = param or FilterParam()
    chroms_to_process, _, reg = glue.setup_valid_regions(pc.bampaths[0], chroms, region, fasta)
    if reg is not None:
        chroms_to_process = [reg]
    parts = []
    for ctig in chroms_to_process:
        sub = glue.prepare_pileup_call(bamfiles, fasta, sites, ctig, None, param, umi_tag, indexes)
        parts.append(result_from_columns(call(sub), pc.sample_ids))
    return rbind(parts)


def pileup_sites(bamfiles, fasta, sites: Optional[Sites] = None, region: Optional[str] = None, chroms=None,
                 param: Optional[FilterParam] = None, umi_tag: Optional[str] = None, indexes=None) -> PileupResult:
    """Mirror of ``pileup_sites()`` (R/pileup.R:106-243)."""
    from . import _cabi

    return _pileup_sites_impl(_cabi.pileup, bamfiles, fasta, sites, region, chroms, param, umi_tag, indexes)


# ----------------------------------------------------------------------------- sparse array IO
def read_sparray(mtx_fn: str, sites_fn: str, bc_fn: str, site_format: str = "coordinate") -> Optional[SparseCounts]:
    """Mirror of ``read_sparray()`` (R/sc-pileup.R:390-467)."""
    if site_format not in ("coordinate", "index"):
        raise ValueError("site_format must be 'coordinate' or 'index'")
    opener = lambda p: gzip.open(p, "rt") if p.endswith(".gz") else open(p, "rt")  # noqa: E731
    if not os.path.getsize(sites_fn) > 0:
        return None
    idx, sn, st, sd, rf, al = [], [], [], [], [], []
    with opener(sites_fn) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 6:
                continue
            idx.append(int(f[0])); sn.append(f[1]); st.append(int(f[2])); sd.append(int(f[3]))
            rf.append(f[4]); al.append(f[5])
    with opener(bc_fn) as fh:
        cnames = [x.rstrip("\n") for x in fh if x.rstrip("\n") != ""]
    rows, cols_, nref, nalt = [], [], [], []
    with opener(mtx_fn) as fh:
        lines = fh.readlines()
    skip = 3 if lines and lines[0].startswith("%%%") else 0
    for line in lines[skip:]:
        f = line.split()
        if not f:
            continue
        if len(f) != 4:
            raise ValueError("malformed sparseMatrix")
        rows.append(int(f[0])); cols_.append(int(f[1])); nref.append(int(f[2])); nalt.append(int(f[3]))
    n = len(idx)
    # the matrix row is the 1-based index column of the triplet; file order of sites names the rows
    sc = SparseCounts(
        n_sites=n, barcodes=cnames, row=np.asarray(rows, dtype=np.int64) - 1, col=np.asarray(cols_, dtype=np.int64) - 1,
        nRef=np.asarray(nref, dtype=np.int64), nAlt=np.asarray(nalt, dtype=np.int64), site_seqnames=sn,
        site_pos=np.asarray(st, dtype=np.int32), site_strand=["+" if s == 1 else "-" for s in sd], site_ref=rf,
        site_alt=al,
    )
    sc.index = idx  # type: ignore[attr-defined]
    return sc


def write_sparray(sc: SparseCounts, mtx_fn: str, sites_fn: str, bc_fn: str) -> None:
    """Mirror of ``write_sparray()`` (R/sc-pileup.R:470-530): gz outputs, MatrixMarket-like header."""
    order = np.lexsort((sc.row, sc.col)) if len(sc.row) else np.zeros(0, dtype=np.int64)
    with gzip.open(mtx_fn, "wt") as fh:
        fh.write("%%% raer MatrixMarket-like matrix coordinate integer general\n")
        fh.write("%%% x y nRef nAlt\n")
        fh.write(f"%%% {sc.n_sites} {len(sc.barcodes)} {len(sc.row)}\n")
        for k in order:
            fh.write(f"{sc.row[k] + 1} {sc.col[k] + 1} {sc.nRef[k]} {sc.nAlt[k]}\n")
    with gzip.open(sites_fn, "wt") as fh:
        for i in range(sc.n_sites):
            fh.write(f"{i + 1}\t{sc.site_seqnames[i]}\t{sc.site_pos[i]}\t{1 if sc.site_strand[i] == '+' else 2}\t"
                     f"{sc.site_ref[i]}\t{sc.site_alt[i]}\n")
    with gzip.open(bc_fn, "wt") as fh:
        for b in sc.barcodes:
            fh.write(b + "\n")


# ----------------------------------------------------------------------------- pileup_cells
def _get_sc_pileup(call, bamfn, index, ident, sites: Sites, barcodes, outfile_prefix, chrom, umi_tag, cb_tag,
                   param) -> Optional[SparseCounts]:
    """get_sc_pileup() (R/sc-pileup.R:289-345)."""
    if chrom is not None:
        sites = sites.subset([s == chrom for s in sites.seqnames])
    if len(sites) == 0:
        return None
    sc_call = glue.prepare_scpileup_call(bamfn, index, chrom, sites, barcodes, outfile_prefix, ident, cb_tag, umi_tag,
                                         param)
    try:
        res = call(sc_call)
        if isinstance(res, tuple):
            # triplets in memory (raer_scpileup_triplets): what read_sparray() would rebuild from the three files; the
            # sites file lists the sites by contig in order of appearance and position (src/sc-plp.c:852-869)
            site_idx, cb_idx, nref, nalt = res
            sc = SparseCounts(n_sites=len(sites), barcodes=list(barcodes), row=site_idx - 1, col=cb_idx - 1,
                              nRef=nref, nAlt=nalt)
        else:
            if res < 0:
                raise RuntimeError("pileup failed")
            sc = read_sparray(*sc_call.outfns, site_format="index")
    finally:
        for f in sc_call.outfns:
            if os.path.exists(f):
                os.unlink(f)
    if sc is None:
        return None
    # Matrix rows are addressed by the 1-based index column of the triplets (src/sc-plp.c:430), which is
    # the position of the site in `sites`; label row i with sites[i]. (The reference labels rows with
    # sites[<index column in file order>] (R/sc-pileup.R:341-342), which agrees whenever `sites` is
    # position sorted, as in its tests.)
    assert sc.n_sites == len(sites)
    sc.site_seqnames = list(sites.seqnames)
    sc.site_pos = np.asarray(sites.start, dtype=np.int32)
    sc.site_strand = list(sites.strand)
    sc.site_ref = list(sites.ref)
    sc.site_alt = list(sites.alt)
    return sc


def _pileup_cells_impl(call, bamfiles, sites: Sites, cell_barcodes, output_directory, chroms=None, umi_tag="UB",
                       cb_tag="CB", param: Optional[FilterParam] = None, return_sce: bool = True, indexes=None,
                       workers: int = 1, write_files: bool = True):
    param = param or FilterParam()
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    bamfiles = [os.path.expanduser(b) for b in bamfiles]
    if not isinstance(cell_barcodes, (list, tuple)) or not all(isinstance(b, str) or b is None for b in cell_barcodes):
        raise TypeError("cell_barcodes must be a character vector")
    process_nbam = len(bamfiles) > 1
    if process_nbam and len(bamfiles) != len(cell_barcodes):
        raise ValueError("multiple bamfiles detected a character vector of equal length to the number of input "
                         "bams must be supplied to cell_barcodes")
    for b in bamfiles:
        if not os.path.exists(b):
            raise FileNotFoundError(f"bamfile(s) not found: {b}")
    if indexes is None:
        indexes = [glue.find_index(b) for b in bamfiles]
    for b, ix in zip(bamfiles, indexes):
        if ix is None or not os.path.exists(ix):
            raise FileNotFoundError(f"index file(s) not found for bam: {b}")
    os.makedirs(output_directory, exist_ok=True)
    if not isinstance(sites, Sites):
        raise TypeError("sites provided must be a Sites (GRanges) object")
    cell_barcodes = [b for b in cell_barcodes if b is not None]
    cb_tag = glue.check_tag(cb_tag)
    umi_tag = glue.check_tag(umi_tag)
    chroms_valid, _, _ = glue.setup_valid_regions(bamfiles[0], chroms)
    site_chroms = set(sites.seqnames)
    chroms_to_process = [c for c in chroms_valid if c in site_chroms]
    sites = sites.subset([s in chroms_to_process for s in sites.seqnames])

    sces: List[SparseCounts] = []
    if process_nbam:
        n_chunks = max(1, min(workers, len(bamfiles)))
        bounds = np.linspace(0, len(bamfiles), n_chunks + 1).astype(int)
        for k in range(n_chunks):
            lo, hi = bounds[k], bounds[k + 1]
            if hi <= lo:
                continue
            sc = _get_sc_pileup(call, bamfiles[lo:hi], list(indexes[lo:hi]), k + 1, sites, cell_barcodes[lo:hi],
                                output_directory, None, umi_tag, cb_tag, param)
            if sc is not None:
                sces.append(sc)
    else:
        for chrom in chroms_to_process:
            sc = _get_sc_pileup(call, bamfiles[0], indexes[0], chrom, sites, cell_barcodes, output_directory, chrom,
                                umi_tag, cb_tag, param)
            if sc is not None:
                sces.append(sc)

    outfns = {k: os.path.join(output_directory, v) for k, v in
              (("counts", "counts.mtx.gz"), ("sites", "sites.txt.gz"), ("bc", "barcodes.txt.gz"))}
    if not sces:
        return None if return_sce else outfns
    if process_nbam:  # cbind
        base = sces[0]
        col_off = 0
        rows, cols_, nr, na, bcs = [], [], [], [], []
        for sc in sces:
            rows.append(sc.row); cols_.append(sc.col + col_off); nr.append(sc.nRef); na.append(sc.nAlt)
            bcs += sc.barcodes
            col_off += len(sc.barcodes)
        sce = SparseCounts(base.n_sites, bcs, np.concatenate(rows), np.concatenate(cols_), np.concatenate(nr),
                           np.concatenate(na), base.site_seqnames, base.site_pos, base.site_strand, base.site_ref,
                           base.site_alt)
    else:  # rbind
        row_off = 0
        rows, cols_, nr, na = [], [], [], []
        sn, sp, ss, sr, sa = [], [], [], [], []
        for sc in sces:
            rows.append(sc.row + row_off); cols_.append(sc.col); nr.append(sc.nRef); na.append(sc.nAlt)
            sn += sc.site_seqnames; sp += list(sc.site_pos); ss += sc.site_strand; sr += sc.site_ref; sa += sc.site_alt
            row_off += sc.n_sites
        sce = SparseCounts(row_off, sces[0].barcodes, np.concatenate(rows), np.concatenate(cols_), np.concatenate(nr),
                           np.concatenate(na), sn, np.asarray(sp, dtype=np.int32), ss, sr, sa)
    keep = np.ones(sce.n_sites, dtype=bool)
    if param.min_depth > 0:
        keep &= (sce.row_sums("nRef") + sce.row_sums("nAlt")) >= param.min_depth
    if param.min_variant_reads > 0:
        keep &= sce.row_sums("nAlt") >= param.min_variant_reads
    sce = sce.subset_rows(keep)
    if write_files:
        write_sparray(sce, outfns["counts"], outfns["sites"], outfns["bc"])
    return sce if return_sce else outfns


def pileup_cells(bamfiles, sites: Sites, cell_barcodes: Sequence[str], output_directory: str, chroms=None,
                 umi_tag: Optional[str] = "UB", cb_tag: Optional[str] = "CB", param: Optional[FilterParam] = None,
                 return_sce: bool = True, indexes=None):
    """Mirror of ``pileup_cells()`` (R/sc-pileup.R:114-281)."""
    from . import _cabi

    return _pileup_cells_impl(_cabi.scpileup, bamfiles, sites, cell_barcodes, output_directory, chroms, umi_tag,
                              cb_tag, param, return_sce, indexes)


def pileup_cells_triplets(bamfiles, sites: Sites, cell_barcodes: Sequence[str], chroms=None, umi_tag: Optional[str] = "UB",
                          cb_tag: Optional[str] = "CB", param: Optional[FilterParam] = None, indexes=None):
    """``pileup_cells()`` without the text-file round trip of the reference (C writes three files, R reads them back,
    R/sc-pileup.R:334, 390-467): the nonzero cells come back from the engine in memory. Same
    arguments minus ``output_directory``; returns the same object as ``pileup_cells(..., return_sce=True)``."""
    import tempfile

    from . import _cabi

    with tempfile.TemporaryDirectory() as d:  # only names are derived from it; nothing is written
        return _pileup_cells_impl(_cabi.scpileup_triplets, bamfiles, sites, cell_barcodes, d, chroms, umi_tag, cb_tag,
                                  param, True, indexes, write_files=False)


# ----------------------------------------------------------------------------- calc_AEI
_AEI_PAIRS = [(r, a) for r in "ACGT" for a in "ACGT" if r != a]


def _sites_minus_snps(alu: Sites, snp_db: Optional[Sites], chroms) -> Sites:
    """alu_ranges restricted to `chroms`, with the SNP positions cut out (1-based inclusive pieces)."""
    import bisect
    from collections import defaultdict

    snps = defaultdict(list)
    if snp_db is not None:
        for sn, p0, p1 in zip(snp_db.seqnames, snp_db.start, snp_db.end):
            snps[sn].extend(range(int(p0), int(p1) + 1))
        for v in snps.values():
            v.sort()
    names, starts, ends = [], [], []
    keep = set(chroms)
    for sn, a, b in zip(alu.seqnames, alu.start, alu.end):
        if sn not in keep:
            continue
        cur = int(a)
        ps = snps.get(sn, [])
        for p in ps[bisect.bisect_left(ps, cur):bisect.bisect_right(ps, int(b))]:
            if p > cur:
                names.append(sn); starts.append(cur); ends.append(p - 1)
            cur = max(cur, p + 1)
        if cur <= int(b):
            names.append(sn); starts.append(cur); ends.append(int(b))
    return Sites(names, starts, ends)


def _merge_intervals(iv):
    """sorted, disjoint union of closed integer intervals"""
    out = []
    for a, b in sorted(iv):
        if out and a <= out[-1][1] + 1:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def _subtract_intervals(xs, ys):
    """xs minus ys (both sorted, disjoint, closed)"""
    out, j = [], 0
    for a, b in xs:
        cur = a
        while j < len(ys) and ys[j][1] < cur:
            j += 1
        k = j
        while k < len(ys) and ys[k][0] <= b:
            if ys[k][0] > cur:
                out.append([cur, ys[k][0] - 1])
            cur = max(cur, ys[k][1] + 1)
            k += 1
        if cur <= b:
            out.append([cur, b])
    return out


def _intersect_intervals(xs, ys):
    out, i, j = [], 0, 0
    while i < len(xs) and j < len(ys):
        a, b = max(xs[i][0], ys[j][0]), min(xs[i][1], ys[j][1])
        if a <= b:
            out.append([a, b])
        if xs[i][1] < ys[j][1]:
            i += 1
        else:
            j += 1
    return out


def _sites_by_gene_strand(sites: Sites, genes: Sites):
    """correct_strand() as interval arithmetic (R/calc_AEI.R:239-272): a site takes the strand of the genes that overlap
    it when they agree; sites under genes of both strands or under none are dropped. Returns (plus, minus)."""
    from collections import defaultdict

    by = {"+": defaultdict(list), "-": defaultdict(list), "*": defaultdict(list)}
    for sn, a, b, st in zip(genes.seqnames, genes.start, genes.end, genes.strand):
        by[st if st in by else "*"][sn].append((int(a), int(b)))
    per_chr = defaultdict(list)
    for sn, a, b in zip(sites.seqnames, sites.start, sites.end):
        per_chr[sn].append((int(a), int(b)))
    out = {"+": ([], [], []), "-": ([], [], [])}
    for sn in dict.fromkeys(sites.seqnames):
        mine = _merge_intervals(per_chr[sn])
        up, um, us = _merge_intervals(by["+"][sn]), _merge_intervals(by["-"][sn]), _merge_intervals(by["*"][sn])
        # a gene without a strand makes the sites under it ambiguous (or, alone, unstranded): dropped either way
        for st, only in (("+", _subtract_intervals(_subtract_intervals(up, um), us)),
                         ("-", _subtract_intervals(_subtract_intervals(um, up), us))):
            for a, b in _intersect_intervals(mine, only):
                out[st][0].append(sn); out[st][1].append(a); out[st][2].append(b)
    return Sites(*out["+"]), Sites(*out["-"])


def _calc_aei_fused(fused_call, bamfiles, fasta, alu_ranges, snp_db, param, indexes, txdb_genes=None):
    """One device call per BAM returns the 4 x 4 (REF of the row, counted base) sums of the rows pileup_sites() would have
    produced inside alu_ranges minus the SNP positions (raer_pileup_args.aei_mode). Unstranded libraries: every row is
    on '+', and correct_strand() complements the rows that lie under minus-strand genes; here the site intervals are split
    by gene strand up front, the reduction runs once per class, and the minus class is complemented as a matrix."""
    import numpy as np

    aei, per = {}, {}
    idx = {"A": 0, "C": 1, "G": 2, "T": 3}
    unstranded = param.library_type[0] == "unstranded"
    for bi, bam in enumerate(bamfiles):
        chroms, _, _ = glue.setup_valid_regions(bam, None, None, fasta)
        alu_chroms = set(alu_ranges.seqnames)
        chroms = [c for c in chroms if c in alu_chroms]
        sites = _sites_minus_snps(alu_ranges, snp_db, chroms)
        name = os.path.basename(bam)
        tot = {p: (0, 0) for p in _AEI_PAIRS}
        ix = None if indexes is None else [indexes[bi]]

        def reduce(st_sites):
            if len(st_sites) == 0:
                return np.zeros((4, 4), dtype=np.int64)
            pc = glue.prepare_pileup_call([bam], fasta, st_sites, None, None, param, None, ix)
            return np.asarray(fused_call(pc)["aei"][0], dtype=np.int64)

        if len(sites) > 0:
            if unstranded:
                plus, minus = _sites_by_gene_strand(sites, txdb_genes)
                m = reduce(plus) + reduce(minus)[::-1, ::-1]  # A C G T reversed = T G C A: the complement of both axes
            else:
                m = reduce(sites)
            tot = {(r, a): (int(m[idx[r], idx[a]]), int(m[idx[r], idx[r]])) for r, a in _AEI_PAIRS}
        aei[name] = {f"{r}_{a}": (100.0 * alt / (alt + ref) if alt + ref > 0 else float("nan")) for (r, a), (alt, ref) in tot.items()}
        per[name] = {f"{r}_{a}": v for (r, a), v in tot.items()}
    return {"AEI": aei, "AEI_per_chrom": per}


def _calc_aei_impl(call, bamfiles, fasta, alu_ranges: Optional[Sites] = None, txdb_genes: Optional[Sites] = None,
                   snp_db: Optional[Sites] = None, param: Optional[FilterParam] = None, indexes=None, fused_call=None):
    """calc_AEI() (R/calc_AEI.R:68-323): per BAM, per chromosome pileup with min_depth=1 and
    only_keep_variants=FALSE, then per (REF, ALT) pair: sum(n<ALT>) / (sum(n<ALT>) + sum(n<REF>)) * 100.
    """
    import copy

    param = copy.deepcopy(param) if param is not None else FilterParam()
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    if alu_ranges is None:
        raise ValueError("alu_ranges must be supplied")
    param.min_depth = 1
    param.only_keep_variants = [False]
    unstranded = param.library_type[0] == "unstranded"
    if unstranded and txdb_genes is None:
        raise ValueError("txdb required for processing unstranded data")
    if fused_call is not None:
        return _calc_aei_fused(fused_call, bamfiles, fasta, alu_ranges, snp_db, param, indexes, txdb_genes)
    snp_keys = None
    if snp_db is not None:
        snp_keys = set(zip(snp_db.seqnames, snp_db.start))
    aei, aei_per_site_counts = {}, {}
    for bi, bam in enumerate(bamfiles):
        chroms, _, _ = glue.setup_valid_regions(bam, None, None, fasta)
        alu_chroms = set(alu_ranges.seqnames)
        chroms = [c for c in chroms if c in alu_chroms]
        tot_alt = {p: 0 for p in _AEI_PAIRS}
        tot_ref = {p: 0 for p in _AEI_PAIRS}
        for chrom in chroms:
            ix = None if indexes is None else [indexes[bi]]
            res = _pileup_sites_impl(call, [bam], fasta, alu_ranges, None, chrom, param, None, ix)
            if len(res) == 0:
                continue
            keep = np.ones(len(res), dtype=bool)
            if snp_keys is not None:
                keep = np.array([(s, int(p)) not in snp_keys for s, p in zip(res.seqnames, res.pos)], dtype=bool)
            ref = res.REF
            strand = res.strand
            if unstranded:
                # correct_strand(): assign strand from overlapping gene annotation, complement - rows
                gene_strand = _annot_strand(res, txdb_genes)
                keep &= gene_strand != "*"
                comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
                ref = np.array([comp.get(r, r) if gs == "-" else r for r, gs in zip(ref, gene_strand)])
                strand = gene_strand
            for r, a in _AEI_PAIRS:
                m = keep & (ref == r)
                if not m.any():
                    continue
                if unstranded:
                    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
                    plus = m & (strand == "+")
                    minus = m & (strand == "-")
                    tot_alt[(r, a)] += int(res.assay(f"n{a}")[plus, 0].sum() + res.assay(f"n{comp[a]}")[minus, 0].sum())
                    tot_ref[(r, a)] += int(res.assay(f"n{r}")[plus, 0].sum() + res.assay(f"n{comp[r]}")[minus, 0].sum())
                else:
                    tot_alt[(r, a)] += int(res.assay(f"n{a}")[m, 0].sum())
                    tot_ref[(r, a)] += int(res.assay(f"n{r}")[m, 0].sum())
        name = os.path.basename(bam)
        aei[name] = {f"{r}_{a}": (100.0 * tot_alt[(r, a)] / (tot_alt[(r, a)] + tot_ref[(r, a)])
                                  if (tot_alt[(r, a)] + tot_ref[(r, a)]) > 0 else float("nan")) for r, a in _AEI_PAIRS}
        aei_per_site_counts[name] = {f"{r}_{a}": (tot_alt[(r, a)], tot_ref[(r, a)]) for r, a in _AEI_PAIRS}
    return {"AEI": aei, "AEI_per_chrom": aei_per_site_counts}


def _annot_strand(res: PileupResult, genes: Sites) -> np.ndarray:
    out = np.array(["*"] * len(res), dtype=object)
    for i, (s, p) in enumerate(zip(res.seqnames, res.pos)):
        strands = {genes.strand[k] for k in range(len(genes))
                   if genes.seqnames[k] == s and genes.start[k] <= p <= genes.end[k]}
        if len(strands) == 1:
            out[i] = strands.pop()
    return out.astype(str)


def calc_AEI(bamfiles, fasta, alu_ranges=None, txdb=None, snp_db=None, param=None, indexes=None):
    """Mirror of ``calc_AEI()`` (R/calc_AEI.R:68)."""
    from . import _cabi

    # the fused device reduction; for unstranded libraries the site intervals are split by the
    # strand of the overlapping genes first (correct_strand) and the minus class is complemented as a matrix
    return _calc_aei_impl(_cabi.pileup, bamfiles, fasta, alu_ranges, txdb, snp_db, param, indexes,
                          fused_call=lambda pc: _cabi.pileup(pc, aei=True))

"""Argument marshalling above the ``.Call`` boundary.

This is the Python mirror of the R glue the reference keeps in front of its C entry points; it
produces exactly the positional arguments of ``.Call(".pileup", ...)`` (src/plp.c:1143-1145) and
``.Call(".scpileup", ...)`` (src/sc-plp.c:945-948). No computation of the hot path happens here.

Reference: R/pileup.R:106-243 (pileup_sites), :251-306 (setup_valid_regions), :670-678
(gr_to_cregions); R/sc-pileup.R:114-269 (pileup_cells), :289-345 (get_sc_pileup), :532-558 (gr_to_regions).
"""
from __future__ import annotations

import os
import struct
import warnings
import zlib
import numpy as np
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from .filter_param import DEFAULT_BULK_BAM_FLAGS, DEFAULT_SC_BAM_FLAGS, FilterParam


def read_bam_header(path: str) -> List[Tuple[str, int]]:
    """(name, length) of every @SQ in BAM header order (seqinfo_from_header in the R glue)."""
    out = bytearray()
    with open(path, "rb") as fh:
        need = 12
        parsed = None
        while True:
            hdr = fh.read(18)
            if len(hdr) < 18:
                break
            if hdr[:4] != b"\x1f\x8b\x08\x04":
                raise ValueError(f"{path}: not a BGZF file")
            xlen = struct.unpack("<H", hdr[10:12])[0]
            extra = hdr[12:18] + fh.read(xlen - 6)
            bsize = None
            o = 0
            while o + 4 <= xlen:
                slen = struct.unpack("<H", extra[o + 2:o + 4])[0]
                if extra[o:o + 2] == b"BC":
                    bsize = struct.unpack("<H", extra[o + 4:o + 6])[0]
                o += 4 + slen
            cdata = fh.read(bsize + 1 - 12 - xlen)
            out += zlib.decompress(cdata[:-8], -15)
            if len(out) >= need and parsed is None:
                if out[:4] != b"BAM\x01":
                    raise ValueError(f"{path}: not a BAM file")
                l_text = struct.unpack("<i", out[4:8])[0]
                need = 12 + l_text
                parsed = l_text
            if parsed is not None and len(out) >= need:
                # try to parse all references; ask for more data on short reads
                try:
                    o = 8 + parsed
                    n_ref = struct.unpack("<i", out[o:o + 4])[0]
                    o += 4
                    refs = []
                    for _ in range(n_ref):
                        ln = struct.unpack("<i", out[o:o + 4])[0]
                        name = bytes(out[o + 4:o + 4 + ln - 1]).decode()
                        if len(out) < o + 4 + ln + 4:
                            raise struct.error
                        length = struct.unpack("<i", out[o + 4 + ln:o + 8 + ln])[0]
                        refs.append((name, length))
                        o += 8 + ln
                    return refs
                except struct.error:
                    continue
    raise ValueError(f"{path}: truncated BAM header")


def read_fai_names(fasta: str) -> List[str]:
    names = []
    with open(fasta + ".fai") as fh:
        for line in fh:
            if line.strip():
                names.append(line.split("\t", 1)[0])
    return names


def find_index(bam: str) -> Optional[str]:
    for cand in (bam + ".bai", os.path.splitext(bam)[0] + ".bai"):
        if os.path.exists(cand):
            return cand
    return None


def check_tag(tag: Optional[str]) -> Optional[str]:
    """R/utils.R check_tag: NULL -> character(); must be a 2-character string otherwise."""
    if tag is None:
        return None
    if not isinstance(tag, str) or len(tag) != 2:
        raise ValueError("tag must be a character(1) with nchar of 2")
    return tag


def setup_valid_regions(bam: str, chroms=None, region=None, fasta=None):
    """R/pileup.R:251-306."""
    contigs = read_bam_header(bam)
    names = [c[0] for c in contigs]
    chroms_to_process = list(names) if chroms is None else ([chroms] if isinstance(chroms, str) else list(chroms))
    if region is None and chroms is not None and len(chroms_to_process) == 1:
        region = chroms_to_process[0]
    missing = [c for c in chroms_to_process if c not in names]
    if missing:
        warnings.warn("the following chromosomes are not present in the bamfile(s):\n" + "\n".join(missing))
        chroms_to_process = [c for c in chroms_to_process if c not in missing]
    if fasta is not None:
        fa_names = set(read_fai_names(fasta))
        missing = [c for c in chroms_to_process if c not in fa_names]
        if missing:
            warnings.warn("the following chromosomes are not present in the fasta file:\n" + "\n".join(missing))
            chroms_to_process = [c for c in chroms_to_process if c not in missing]
    chroms_to_process.sort(key=names.index)
    if not chroms_to_process:
        raise ValueError("There are no valid chromosomes to process")
    return chroms_to_process, contigs, region


@dataclass
class Sites:
    """Minimal stand-in for the GRanges the R API takes for ``sites``: 1-based inclusive ranges."""

    seqnames: Sequence[str]
    start: Sequence[int]
    end: Optional[Sequence[int]] = None
    strand: Optional[Sequence[str]] = None
    ref: Optional[Sequence[str]] = None  # REF / ALT columns used by pileup_cells
    alt: Optional[Sequence[str]] = None

    def __post_init__(self):
        def ints(v):  # plain Python ints (numpy arrays and lists of numpy scalars included) without a Python-level loop
            return np.asarray(v, dtype=np.int64).tolist() if len(v) else []

        self.seqnames = list(self.seqnames)
        self.start = ints(self.start)
        self.end = list(self.start) if self.end is None else ints(self.end)
        n = len(self.seqnames)
        if self.strand is None:
            self.strand = ["*"] * n
        self.strand = list(self.strand)
        if len(self.start) != n or len(self.end) != n or len(self.strand) != n:
            raise ValueError("seqnames, start, end and strand must have the same length")

    def __len__(self):
        return len(self.seqnames)

    def subset(self, keep: Sequence[bool]) -> "Sites":
        from itertools import compress

        keep = [bool(k) for k in keep] if not isinstance(keep, list) else keep
        pick = lambda v: None if v is None else list(compress(v, keep))  # noqa: E731
        return Sites(pick(self.seqnames), pick(self.start), pick(self.end), pick(self.strand), pick(self.ref),
                     pick(self.alt))

    @staticmethod
    def from_bed(path: str) -> "Sites":
        """rtracklayer::import of a BED file: 0-based half-open -> 1-based inclusive."""
        import gzip

        opener = gzip.open if path.endswith(".gz") else open
        sn, st, en, sd = [], [], [], []
        with opener(path, "rt") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if len(f) < 3 or line.startswith(("#", "track", "browser")):
                    continue
                sn.append(f[0])
                st.append(int(f[1]) + 1)
                en.append(int(f[2]))
                sd.append(f[5] if len(f) > 5 and f[5] in "+-" else "*")
        return Sites(sn, st, en, sd)


@dataclass
class PileupCall:
    """The 13 positional arguments of .Call(".pileup") (src/plp.c:1143-1145)."""

    bampaths: List[str]
    indexes: List[str]
    n: int
    fapath: str
    region: Optional[str]
    lst: Optional[Tuple[List[str], List[int], List[int]]]
    int_args: List[int]
    dbl_args: List[float]
    lgl_args: List[bool]
    libtype: List[int]
    only_keep_variants: List[bool]
    min_mapq: List[int]
    umi: Optional[str]
    contigs: List[Tuple[str, int]] = field(default_factory=list)
    sample_ids: List[str] = field(default_factory=list)


def prepare_pileup_call(bamfiles, fasta, sites: Optional[Sites] = None, region=None, chroms=None,
                        param: Optional[FilterParam] = None, umi_tag=None, indexes=None) -> PileupCall:
    """Everything pileup_sites() does before .Call (R/pileup.R:106-198), serial path."""
    param = param or FilterParam()
    if isinstance(bamfiles, str):
        bamfiles = [bamfiles]
    bamfiles = [os.path.expanduser(b) for b in bamfiles]
    n_files = len(bamfiles)
    if sites is not None and not isinstance(sites, Sites):
        raise TypeError(f"invalid object passed to sites, expecting Sites found {type(sites).__name__}")
    for b in bamfiles:
        if not os.path.exists(b):
            raise FileNotFoundError(f"bamfile(s) not found: {b}")
    if indexes is None:
        indexes = [find_index(b) for b in bamfiles]
    for b, ix in zip(bamfiles, indexes):
        if ix is None or not os.path.exists(ix):
            raise FileNotFoundError(f"index file(s) not found for bam: {b}")
    if not os.path.exists(fasta):
        raise FileNotFoundError(f"fasta file not found: {fasta}")
    fasta = os.path.expanduser(fasta)

    chroms_to_process, contigs, region = setup_valid_regions(bamfiles[0], chroms, region, fasta)

    lst = None
    if sites is not None:
        sites = sites.subset([s in chroms_to_process for s in sites.seqnames])
        if len(sites) == 0:
            raise ValueError("No entries in GRanges")
        lst = (list(sites.seqnames), list(sites.start), list(sites.end))

    umi_tag = check_tag(umi_tag)
    args = param.c_args(n_files, DEFAULT_BULK_BAM_FLAGS)

    if region is not None:
        chroms_to_process = [region]
    if len(chroms_to_process) > 1 and lst is None:
        # many contigs: whole contigs are passed as the region list (R/pileup.R:179-182)
        lens = dict(contigs)
        lst = (list(chroms_to_process), [1] * len(chroms_to_process), [lens[c] for c in chroms_to_process])
    return PileupCall(
        bampaths=bamfiles, indexes=list(indexes), n=n_files, fapath=fasta, region=region, lst=lst,
        int_args=args["int_args"], dbl_args=args["dbl_args"], lgl_args=args["lgl_args"], libtype=args["libtype"],
        only_keep_variants=args["only_keep_variants"], min_mapq=args["min_mapq"], umi=umi_tag, contigs=contigs,
        sample_ids=[os.path.basename(b) for b in bamfiles],
    )


@dataclass
class ScPileupCall:
    """The 14 positional arguments of .Call(".scpileup") (src/sc-plp.c:945-948)."""

    bampaths: List[str]
    indexes: List[str]
    query_region: Optional[str]
    lst: Tuple[List[str], List[int], List[int], List[str], List[str], List[int]]
    barcodes: List[str]
    cbtag: Optional[str]
    int_args: List[int]
    dbl_args: List[float]
    libtype: int
    outfns: Tuple[str, str, str]
    umi: Optional[str]
    remove_overlaps: bool
    min_mapq: int
    min_counts: int


def gr_to_regions(sites: Sites):
    """R/sc-pileup.R:532-558: positions must be width 1, stranded, with single-base REF and ALT."""
    n = len(sites)
    if sites.ref is None or sites.alt is None:
        raise ValueError("sites must have REF and ALT columns")
    if any(s != e for s, e in zip(sites.start, sites.end)):
        raise ValueError("each site must be a single position (start = end)")
    if any(s == "*" for s in sites.strand):
        warnings.warn("missing strand not found in input, coercing strand to '+'")
        sites = Sites(sites.seqnames, sites.start, sites.end, ["+"] * n, sites.ref, sites.alt)
    if any(len(r) != 1 for r in sites.ref) or any(len(a) != 1 for a in sites.alt):
        raise ValueError("REF and ALT must be single bases")
    strand = [1 if s == "+" else 2 for s in sites.strand]
    return (list(sites.seqnames), list(sites.start), strand, list(sites.ref), list(sites.alt), list(range(1, n + 1)))


def prepare_scpileup_call(bamfn, index, chrom, sites: Sites, barcodes, outfile_prefix, ident, cb_tag, umi_tag,
                          param: FilterParam) -> ScPileupCall:
    """get_sc_pileup() up to .Call (R/sc-pileup.R:289-327)."""
    if isinstance(bamfn, str):
        bamfn = [bamfn]
        index = [index]
    plp_outfns = tuple(os.path.join(outfile_prefix, f"{ident}_{x}") for x in ("counts.mtx", "sites.txt", "bcs.txt"))
    args = param.c_args(1, DEFAULT_SC_BAM_FLAGS)
    ia = args["int_args"]
    return ScPileupCall(
        bampaths=list(bamfn), indexes=list(index), query_region=chrom, lst=gr_to_regions(sites),
        barcodes=list(barcodes), cbtag=cb_tag, int_args=ia, dbl_args=args["dbl_args"], libtype=args["libtype"][0],
        outfns=plp_outfns, umi=umi_tag, remove_overlaps=args["lgl_args"][1], min_mapq=args["min_mapq"][0],
        min_counts=max(ia[9], ia[1]),
    )

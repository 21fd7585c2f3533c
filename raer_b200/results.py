"""Result containers: light-weight stand-ins for the Bioconductor objects the R API returns.

``PileupResult`` plays the role of the RangedSummarizedExperiment built by ``lists_to_grs`` +
``merge_pileups`` (R/pileup.R:654-668, 707-762): rowRanges (seqnames, pos, strand), rowData
(REF, rpbz, vdb, sor) and one matrix per assay with a column per BAM file. Every file carries
identical rows at the .Call boundary (src/plp.c:445-459), so no union/NA filling is needed.
``SparseCounts`` plays the role of the SingleCellExperiment returned by ``read_sparray``
(R/sc-pileup.R:390-467): nRef / nAlt sparse matrices, sites x barcodes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

ASSAY_COLS = ("ALT", "nRef", "nAlt", "nA", "nT", "nC", "nG")  # merge_pileups default, R/pileup.R:708
ALL_COLS = ASSAY_COLS + ("nN", "nX")


@dataclass
class PileupResult:
    seqnames: np.ndarray  # str
    pos: np.ndarray       # int32, 1-based
    strand: np.ndarray    # '+' / '-'
    REF: np.ndarray
    rpbz: np.ndarray
    vdb: np.ndarray
    sor: np.ndarray
    assays: Dict[str, np.ndarray]  # name -> [n_rows, n_samples]
    sample_names: List[str] = field(default_factory=list)
    extra: Dict[str, np.ndarray] = field(default_factory=dict)  # nN, nX (returned by C, dropped by merge_pileups)

    def __len__(self):
        return int(self.pos.shape[0])

    @property
    def nrow(self):
        return len(self)

    @property
    def ncol(self):
        return len(self.sample_names)

    def assay(self, name: str) -> np.ndarray:
        if name in self.assays:
            return self.assays[name]
        return self.extra[name]

    def rownames(self) -> List[str]:
        """site_names(): site_<seqname>_<pos>_<1|2> (R/utils.R)."""
        return [f"site_{s}_{p}_{1 if st == '+' else 2}" for s, p, st in zip(self.seqnames, self.pos, self.strand)]

    def subset(self, mask) -> "PileupResult":
        mask = np.asarray(mask)
        return PileupResult(
            self.seqnames[mask], self.pos[mask], self.strand[mask], self.REF[mask], self.rpbz[mask], self.vdb[mask],
            self.sor[mask], {k: v[mask] for k, v in self.assays.items()}, list(self.sample_names),
            {k: v[mask] for k, v in self.extra.items()},
        )

    def identical(self, other: "PileupResult", stats_rtol: Optional[float] = None) -> bool:
        if len(self) != len(other) or self.sample_names != other.sample_names:
            return False
        for a, b in ((self.seqnames, other.seqnames), (self.pos, other.pos), (self.strand, other.strand),
                     (self.REF, other.REF)):
            if not np.array_equal(a, b):
                return False
        for k in self.assays:
            if not np.array_equal(self.assays[k], other.assays[k]):
                return False
        for a, b in ((self.rpbz, other.rpbz), (self.vdb, other.vdb), (self.sor, other.sor)):
            if stats_rtol is None:
                if not np.array_equal(a, b):
                    return False
            else:
                fin = np.isfinite(a)
                if not np.array_equal(fin, np.isfinite(b)) or not np.array_equal(a[~fin], b[~fin]):
                    return False
                if not np.allclose(a[fin], b[fin], rtol=stats_rtol, atol=0):
                    return False
        return True


def result_from_columns(cols: dict, sample_names: List[str]) -> PileupResult:
    """Build from the column layout of the .Call result (src/plp_data.c:285-289, 321-361)."""
    n = len(cols["files"])
    assays = {}
    extra = {}
    for name in ALL_COLS:
        mats = [np.asarray(cols["files"][f][name]) for f in range(n)]
        m = np.stack(mats, axis=1) if len(mats[0]) else np.zeros((0, n), dtype=mats[0].dtype)
        (assays if name in ASSAY_COLS else extra)[name] = m
    return PileupResult(
        seqnames=np.asarray(cols["seqname"], dtype=str) if len(cols["seqname"]) else np.zeros(0, dtype=str),
        pos=np.asarray(cols["pos"], dtype=np.int32), strand=np.asarray(cols["strand"]).astype(str),
        REF=np.asarray(cols["REF"]).astype(str), rpbz=np.asarray(cols["rpbz"], dtype=np.float64),
        vdb=np.asarray(cols["vdb"], dtype=np.float64), sor=np.asarray(cols["sor"], dtype=np.float64),
        assays=assays, sample_names=list(sample_names), extra=extra,
    )


def rbind(results: List[PileupResult]) -> PileupResult:
    r0 = results[0]
    cat = lambda xs: np.concatenate(xs) if len(xs) else xs  # noqa: E731
    return PileupResult(
        cat([r.seqnames for r in results]), cat([r.pos for r in results]), cat([r.strand for r in results]),
        cat([r.REF for r in results]), cat([r.rpbz for r in results]), cat([r.vdb for r in results]),
        cat([r.sor for r in results]), {k: np.concatenate([r.assays[k] for r in results]) for k in r0.assays},
        list(r0.sample_names), {k: np.concatenate([r.extra[k] for r in results]) for k in r0.extra},
    )


@dataclass
class SparseCounts:
    """sites x barcodes nRef/nAlt triplets (dgCMatrix stand-in)."""

    n_sites: int
    barcodes: List[str]
    row: np.ndarray   # 0-based site row
    col: np.ndarray   # 0-based barcode column
    nRef: np.ndarray
    nAlt: np.ndarray
    site_seqnames: List[str] = field(default_factory=list)
    site_pos: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))
    site_strand: List[str] = field(default_factory=list)
    site_ref: List[str] = field(default_factory=list)
    site_alt: List[str] = field(default_factory=list)

    @property
    def shape(self):
        return (self.n_sites, len(self.barcodes))

    def dense(self, which: str) -> np.ndarray:
        m = np.zeros(self.shape, dtype=np.int64)
        np.add.at(m, (self.row, self.col), getattr(self, which))
        return m

    def row_sums(self, which: str) -> np.ndarray:
        s = np.zeros(self.n_sites, dtype=np.int64)
        np.add.at(s, self.row, getattr(self, which))
        return s

    def subset_rows(self, keep: np.ndarray) -> "SparseCounts":
        keep = np.asarray(keep, dtype=bool)
        newidx = np.cumsum(keep) - 1
        m = keep[self.row]
        pick = lambda v: [x for x, k in zip(v, keep) if k]  # noqa: E731
        return SparseCounts(
            int(keep.sum()), list(self.barcodes), newidx[self.row[m]], self.col[m], self.nRef[m], self.nAlt[m],
            pick(self.site_seqnames), np.asarray(self.site_pos)[keep], pick(self.site_strand), pick(self.site_ref),
            pick(self.site_alt),
        )

    def rownames(self) -> List[str]:
        """site_names(gr, allele = TRUE): site_<seq>_<pos>_<1|2>_<REF><ALT>."""
        return [f"site_{s}_{p}_{1 if st == '+' else 2}_{r}{a}" for s, p, st, r, a in
                zip(self.site_seqnames, self.site_pos, self.site_strand, self.site_ref, self.site_alt)]

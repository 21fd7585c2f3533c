"""FilterParam: host-side mirror of raer's ``FilterParam`` S4 class.

Reference: R/pileup.R:334-359 (slots), :538-632 (constructor, defaults and validation),
:434-477 (``c_args_FilterParam``: the positional flattening consumed by ``.Call(".pileup")`` /
``.Call(".scpileup")``), :313-318 and R/sc-pileup.R:283-287 (default BAM flag filters).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

LIBRARY_TYPES = ("unstranded", "fr-first-strand", "fr-second-strand")

_FLAG_BITS = {
    "isPaired": 0x1,
    "isProperPair": 0x2,
    "isUnmappedQuery": 0x4,
    "hasUnmappedMate": 0x8,
    "isMinusStrand": 0x10,
    "isMateMinusStrand": 0x20,
    "isFirstMateRead": 0x40,
    "isSecondMateRead": 0x80,
    "isSecondaryAlignment": 0x100,
    "isNotPassingQualityControls": 0x200,
    "isDuplicate": 0x400,
    "isSupplementaryAlignment": 0x800,
}
_ALL_FLAGS = 4095


def scan_bam_flag(**kwargs) -> Tuple[int, int]:
    """Equivalent of ``Rsamtools::scanBamFlag``: returns ``(keep0, keep1)``.

    ``isX=True`` keeps only reads with bit X set (clears X in keep0); ``isX=False`` keeps only reads
    with bit X unset (clears X in keep1); unspecified bits are not tested.
    """
    keep0 = keep1 = _ALL_FLAGS
    for name, val in kwargs.items():
        if name == "isNotPrimaryRead":  # deprecated alias in Rsamtools
            name = "isSecondaryAlignment"
        if name not in _FLAG_BITS:
            raise ValueError(f"unknown flag name: {name}")
        if val is None:
            continue
        if val:
            keep0 &= ~_FLAG_BITS[name]
        else:
            keep1 &= ~_FLAG_BITS[name]
    return keep0, keep1


# R/pileup.R:313-318
DEFAULT_BULK_BAM_FLAGS = scan_bam_flag(
    isSecondaryAlignment=False, isNotPassingQualityControls=False, isDuplicate=False, isSupplementaryAlignment=False
)
# R/sc-pileup.R:283-287
DEFAULT_SC_BAM_FLAGS = scan_bam_flag(
    isSecondaryAlignment=False, isSupplementaryAlignment=False, isNotPassingQualityControls=False
)


def _is_single_number(x) -> bool:
    return isinstance(x, (int, float)) and not isinstance(x, bool)


@dataclass
class FilterParam:
    """Same slots, defaults and validation as raer's ``FilterParam()`` (R/pileup.R:538-632)."""

    max_depth: int = 10000
    min_depth: int = 1
    min_base_quality: int = 20
    min_mapq: Sequence[int] | int = 0
    library_type: Sequence[str] | str = "fr-first-strand"
    bam_flags: Optional[Tuple[int, int]] = None
    only_keep_variants: Sequence[bool] | bool = False
    trim_5p: int = 0
    trim_3p: int = 0
    ftrim_5p: float = 0.0
    ftrim_3p: float = 0.0
    indel_dist: int = 0
    splice_dist: int = 0
    min_splice_overhang: int = 0
    homopolymer_len: int = 0
    max_mismatch_type: Sequence[int] = (0, 0)
    read_bqual: Sequence[float] = (0.0, 0.0)
    min_variant_reads: int = 0
    min_allelic_freq: float = 0.0
    report_multiallelic: bool = True
    remove_overlaps: bool = True
    _flags_are_default: bool = field(default=False, repr=False)

    def __post_init__(self):
        for nm in ("max_depth", "min_base_quality", "min_depth", "trim_5p", "trim_3p", "indel_dist", "splice_dist",
                   "homopolymer_len", "min_splice_overhang", "min_variant_reads", "ftrim_5p", "ftrim_3p",
                   "min_allelic_freq"):
            if not _is_single_number(getattr(self, nm)):
                raise ValueError(f"{nm} must be a single number")
        for nm in ("max_depth", "min_base_quality", "min_depth", "trim_5p", "trim_3p", "indel_dist", "splice_dist",
                   "homopolymer_len", "min_splice_overhang", "min_variant_reads"):
            setattr(self, nm, int(getattr(self, nm)))
        self.ftrim_5p = float(self.ftrim_5p)
        self.ftrim_3p = float(self.ftrim_3p)
        self.min_allelic_freq = float(self.min_allelic_freq)
        if not (0 <= self.ftrim_5p <= 1):
            raise ValueError("ftrim_5p must be within [0, 1]")
        if not (0 <= self.ftrim_3p <= 1):
            raise ValueError("ftrim_3p must be within [0, 1]")
        self.max_mismatch_type = tuple(int(x) for x in self.max_mismatch_type)
        self.read_bqual = tuple(float(x) for x in self.read_bqual)
        if len(self.max_mismatch_type) != 2:
            raise ValueError("max_mismatch_type must have length 2")
        if len(self.read_bqual) != 2:
            raise ValueError("read_bqual must have length 2")
        if not isinstance(self.report_multiallelic, bool) or not isinstance(self.remove_overlaps, bool):
            raise ValueError("report_multiallelic / remove_overlaps must be TRUE or FALSE")
        if isinstance(self.library_type, str):
            self.library_type = [self.library_type]
        self.library_type = list(self.library_type)
        for lt in self.library_type:
            if lt not in LIBRARY_TYPES:
                raise ValueError("library_type must be one of: " + " ".join(LIBRARY_TYPES))
        if isinstance(self.min_mapq, int):
            self.min_mapq = [self.min_mapq]
        self.min_mapq = [int(x) for x in self.min_mapq]
        if isinstance(self.only_keep_variants, bool):
            self.only_keep_variants = [self.only_keep_variants]
        self.only_keep_variants = [bool(x) for x in self.only_keep_variants]
        if self.bam_flags is None:
            self.bam_flags = scan_bam_flag()
            self._flags_are_default = True
        else:
            if len(self.bam_flags) != 2:
                raise ValueError("bam_flags must be generated using scan_bam_flag()")
            self.bam_flags = (int(self.bam_flags[0]), int(self.bam_flags[1]))
            self._flags_are_default = tuple(self.bam_flags) == scan_bam_flag()

    @staticmethod
    def _adjust(vals, n, name):
        if len(vals) == n:
            return list(vals)
        if len(vals) == 1:
            return list(vals) * n
        raise ValueError(f"{name} requires either 1 value, or individual values for all input bamfiles")

    def c_args(self, n_files: int, default_flags: Tuple[int, int] = DEFAULT_BULK_BAM_FLAGS) -> dict:
        """Positional C argument vectors (R/pileup.R:434-477, 485-489).

        ``default_flags`` replaces an untouched ``bam_flags`` as pileup_sites()/pileup_cells() do
        (R/pileup.R:163-165, R/sc-pileup.R:166-168).
        """
        flags = default_flags if self._flags_are_default else self.bam_flags
        int_args = [
            self.max_depth, self.min_depth, self.min_base_quality, self.trim_5p, self.trim_3p, self.indel_dist,
            self.splice_dist, self.min_splice_overhang, self.homopolymer_len, self.min_variant_reads,
            self.max_mismatch_type[0], self.max_mismatch_type[1], flags[0], flags[1],
        ]
        dbl_args = [self.ftrim_5p, self.ftrim_3p, self.min_allelic_freq, self.read_bqual[0], self.read_bqual[1]]
        lgl_args = [self.report_multiallelic, self.remove_overlaps]
        return {
            "int_args": int_args,
            "dbl_args": dbl_args,
            "lgl_args": lgl_args,
            "libtype": [LIBRARY_TYPES.index(x) for x in self._adjust(self.library_type, n_files, "library_type")],
            "only_keep_variants": self._adjust(self.only_keep_variants, n_files, "only_keep_variants"),
            "min_mapq": self._adjust(self.min_mapq, n_files, "min_mapq"),
        }

"""Drives the CPU oracle through the same glue as the product API (tests only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from raer_b200 import api  # noqa: E402

OVERLAP_MODE = 1


def oracle_pileup_call(pc):
    return pyoracle.pileup(
        pc.bampaths, pc.fapath, region=pc.region, sites=pc.lst, int_args=pc.int_args, dbl_args=pc.dbl_args,
        lgl_args=pc.lgl_args, libtype=pc.libtype, only_keep_variants=pc.only_keep_variants, min_mapq=pc.min_mapq,
        umi_tag=pc.umi, indexes=pc.indexes, overlap_mode=OVERLAP_MODE)


def oracle_sc_call(sc):
    return pyoracle.scpileup(
        sc.bampaths, region=sc.query_region, sites=sc.lst, barcodes=sc.barcodes, cb_tag=sc.cbtag,
        int_args=sc.int_args, dbl_args=sc.dbl_args, libtype=sc.libtype, outfns=sc.outfns, umi_tag=sc.umi,
        remove_overlaps=sc.remove_overlaps, min_mapq=sc.min_mapq, min_counts=sc.min_counts, indexes=sc.indexes,
        overlap_mode=OVERLAP_MODE)


def pileup_sites(*a, **k):
    return api._pileup_sites_impl(oracle_pileup_call, *a, **k)


def pileup_cells(*a, **k):
    return api._pileup_cells_impl(oracle_sc_call, *a, **k)


def calc_AEI(*a, **k):
    return api._calc_aei_impl(oracle_pileup_call, *a, **k)

"""cProfile of one pileup_sites() call on a synthetic sample (where does the Python layer spend its time)."""
import cProfile
import os
import pstats
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raer_b200 import FilterParam, api, synth  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    ds = synth.make_bulk_dataset(d, n_bams=2, n_contigs=2, contig_len=1_000_000, pairs_per_contig=500_000, seed=77)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True, min_depth=1)
    api.pileup_sites(ds["bams"], ds["fasta"], param=fp)
    pr = cProfile.Profile()
    pr.enable()
    r = api.pileup_sites(ds["bams"], ds["fasta"], param=fp)
    pr.disable()
    print(len(r), "rows")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

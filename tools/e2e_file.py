import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RAER_TIMING"] = "1"
from raer_b200 import FilterParam, api, synth
with tempfile.TemporaryDirectory() as d:
    ds = synth.make_bulk_dataset(d, n_bams=2, n_contigs=4, contig_len=250_000, pairs_per_contig=125_000, seed=4242)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True)
    for _ in range(3):
        t = time.perf_counter(); r = api.pileup_sites(ds["bams"], ds["fasta"], param=fp); dt = time.perf_counter() - t
        print("pileup_sites wall %.3f s rows %d  %.3e bases/s" % (dt, len(r), ds["aligned_bases"] / dt))

import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RAER_TIMING"] = "1"
from raer_b200 import FilterParam, api, synth
with tempfile.TemporaryDirectory() as d:
    ds = synth.make_bulk_dataset(d, n_bams=2, n_contigs=int(os.environ.get("E2E_CONTIGS", 4)), contig_len=int(os.environ.get("E2E_LEN", 250_000)),
                                 pairs_per_contig=int(os.environ.get("E2E_PAIRS", 125_000)), seed=4242)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True)
    for mode in ("contig-parallel", "serial", "contig-parallel", "serial"):
        if mode == "serial":
            os.environ["RAER_SERIAL_INGEST"] = "1"
        else:
            os.environ.pop("RAER_SERIAL_INGEST", None)
        t = time.perf_counter(); r = api.pileup_sites(ds["bams"], ds["fasta"], param=fp); dt = time.perf_counter() - t
        print("%s ingest: pileup_sites wall %.3f s rows %d  %.3e bases/s" % (mode, dt, len(r), ds["aligned_bases"] / dt), flush=True)

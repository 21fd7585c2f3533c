"""File-level throughput of pileup_sites() on a synthetic BAM sample, device ingest against host ingest.
usage: python tools/file_level.py [pairs_per_contig] [contig_len] [n_contigs] [n_bams]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raer_b200 import FilterParam, _cabi, api, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
clen = int(sys.argv[2]) if len(sys.argv) > 2 else 2_500_000
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 2
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    ds = synth.make_bulk_dataset(d, n_bams=nb, n_contigs=nc, contig_len=clen, pairs_per_contig=pairs, seed=77)
    size = sum(os.path.getsize(b) for b in ds["bams"])
    print(f"dataset: {nb} BAMs, {size / 1e6:.1f} MB, {ds['aligned_bases']:.3g} aligned bases, written in {time.perf_counter() - t0:.1f} s", flush=True)
    fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True, min_depth=1)
    res = {}
    for mode in ("1", "0", "1", "0"):
        os.environ["RAER_GPU_INGEST"] = mode
        t0 = time.perf_counter()
        r = api.pileup_sites(ds["bams"], ds["fasta"], param=fp)
        dt = time.perf_counter() - t0
        tm = dict(_cabi.LAST_TIMING)
        print(f"RAER_GPU_INGEST={mode}: {dt:.3f} s, {ds['aligned_bases'] / dt / 1e9:.3f} G aligned bases/s, {len(r)} rows, "
              f"device contigs {tm['n_device_contigs']}, host contigs {tm['n_host_contigs']}, kernel {tm['t_kernel']:.3f} s", flush=True)
        res[mode] = r
    print("identical:", res["1"].identical(res["0"]))

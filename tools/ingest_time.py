"""Host-only timing of the BAM ingest (BGZF inflate + parse + pack), no GPU needed:
python tools/ingest_time.py [dir]  -- generates (or reuses) the e2e_file.py sample in dir."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from raer_b200 import FilterParam, _cabi, glue, synth  # noqa: E402
from test_cabi_host import _ingest  # noqa: E402

d = sys.argv[1] if len(sys.argv) > 1 else "/tmp/ing"
os.makedirs(d, exist_ok=True)
if not os.path.exists(os.path.join(d, "s2.bam.bai")):
    synth.make_bulk_dataset(d, n_bams=2, n_contigs=4, contig_len=250_000, pairs_per_contig=125_000, seed=4242)
bams = [os.path.join(d, "s1.bam"), os.path.join(d, "s2.bam")]
fp = FilterParam(library_type="fr-first-strand", only_keep_variants=True, remove_overlaps=os.environ.get("NO_OVERLAP") is None)
pc = glue.prepare_pileup_call(bams, os.path.join(d, "ref.fa"), param=fp)
for _ in range(3):
    t = time.perf_counter()
    s = _ingest(pc)
    dt = time.perf_counter() - t
    print(sorted(s.items()))
    print("ingest %.3f s  %d records  %.3e records/s  %.3e aligned bases/s" % (dt, s["records"], s["records"] / dt, s["aligned_bases"] / dt))

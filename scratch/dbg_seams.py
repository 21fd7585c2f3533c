import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle_backend as ob
from raer_b200 import api, synth, FilterParam
d = synth.make_bulk_dataset('/tmp/syn2', n_bams=2, n_contigs=3, contig_len=120_000, pairs_per_contig=6000, seed=1001)
for kw in (dict(library_type="fr-first-strand"), dict(library_type="fr-first-strand", remove_overlaps=False)):
    fp = FilterParam(**kw)
    want = ob.pileup_sites(d['bams'], d['fasta'], param=fp)
    for br in (40, 60, 100, 200, 333):
        os.environ["RAER_BATCH_READS"] = str(br)
        got = api.pileup_sites(d['bams'], d['fasta'], param=fp)
        ok = len(got) == len(want)
        nbad = -1
        if ok:
            bad = np.argwhere(got.assay("nRef") != want.assay("nRef"))
            nbad = len(bad)
            if nbad:
                r = bad[0][0]
                print("  first bad", want.seqnames[r], want.pos[r], want.strand[r], got.assay("nRef")[r], want.assay("nRef")[r], bad[:6].tolist())
        print(kw.get("remove_overlaps", True), br, len(got), len(want), nbad)

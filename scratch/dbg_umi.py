import os, sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, pytest
import oracle_backend
from raer_b200 import FilterParam, api
import test_bulk_umi as T
class TF:
    def mktemp(self, n):
        import pathlib; return pathlib.Path(tempfile.mkdtemp())
bam, fa = T.umi_bam.__wrapped__(TF())
fp = FilterParam(library_type="fr-first-strand", min_depth=1)
want = oracle_backend.pileup_sites(bam, fa, param=fp, umi_tag="UB")
got = api.pileup_sites(bam, fa, param=fp, umi_tag="UB")
print(len(got), len(want))
def rows(r):
    return {(str(s), int(p), str(st)): (int(a[0]), int(b[0]), int(x[0])) for s, p, st, a, b, x in zip(r.seqnames, r.pos, r.strand, r.assay("nRef"), r.assay("nAlt"), r.assay("nX"))}
g, w = rows(got), rows(want)
for k in sorted(set(g) | set(w)):
    if g.get(k) != w.get(k): print(k, "got", g.get(k), "want", w.get(k))

"""usage: python tools/launch_summary.py <ncu launch list .csv> "<command>" [n_batches] > profiles/<name>.json
Sums gpu__time_duration per kernel of this library over the last n_batches batches (default 4) of the run."""
import csv
import json
import re
import sys

path, cmd = sys.argv[1], sys.argv[2]
n_batches = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 14 and r[0] != "ID"]
ours = [(re.sub(r"^void ", "", r[4]).split("(")[0], float(r[14]) / 1e6) for r in rows
        if re.search(r"\bk_(mate|tile|pileup|sc_|rs_)", r[4])]
# one k_pileup_* launch per batch: keep what belongs to the last n_batches of them
idx = [i for i, (k, _) in enumerate(ours) if k.startswith("k_pileup")]
start = 0
if len(idx) > n_batches:
    prev = idx[-n_batches - 1]
    start = prev + 1
sel = ours[start:]
tot = sum(ms for _, ms in sel)
kern = {}
for k, ms in sel:
    kern.setdefault(k, 0.0)
    kern[k] += ms
out = {
    "command": cmd,
    "note": f"kernels of the last {n_batches} batches of the run. Per-launch times under ncu are cold-cache and serialised: use the "
            "shares. The D2D refresh of the qualities before every batch (cudaMemcpyAsync, stands in for the H2D of a new batch) "
            "is a memcpy, not a kernel, and is not listed.",
    "launches_in_run": len(rows),
    "total_ms": round(tot, 4),
    "kernels": {k: {"ms": round(ms, 4), "share": round(ms / tot, 4)} for k, ms in kern.items()},
}
print(json.dumps(out, indent=1))

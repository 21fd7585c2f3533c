"""Throughput of the device DEFLATE decoder (k_gi_inflate) on the BGZF blocks of a synthetic BAM, next to zlib on one
host core. usage: python tools/inflate_bench.py [pairs] [contig_len]"""
import os
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raer_b200 import _cabi, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
clen = int(sys.argv[2]) if len(sys.argv) > 2 else 1_200_000
with tempfile.TemporaryDirectory() as d:
    ds = synth.make_bulk_dataset(d, n_bams=1, n_contigs=1, contig_len=clen, pairs_per_contig=pairs, seed=5)
    data = open(ds["bams"][0], "rb").read()
payloads, sizes = [], []
o = 0
while o + 18 <= len(data):
    xlen = int.from_bytes(data[o + 10:o + 12], "little")
    bsize = int.from_bytes(data[o + 16:o + 18], "little")
    payloads.append(data[o + 12 + xlen:o + bsize + 1 - 8])
    sizes.append(int.from_bytes(data[o + bsize + 1 - 4:o + bsize + 1], "little"))
    o += bsize + 1
tot = sum(sizes)
print(f"{len(payloads)} blocks, {len(data) / 1e6:.1f} MB -> {tot / 1e6:.1f} MB")
t0 = time.perf_counter()
ref = [zlib.decompress(p, -15) for p in payloads]
dt = time.perf_counter() - t0
print(f"zlib, one core: {tot / dt / 1e9:.3f} GB/s")
for rep in range(3):
    got, status, ms = _cabi.gpu_inflate(payloads, sizes)
    assert (status == 0).all()
    print(f"k_gi_inflate: {ms:.3f} ms, {tot / ms / 1e6:.2f} GB/s (inflated bytes)")
assert got == ref
print("identical to zlib")

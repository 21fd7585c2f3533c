"""Randomised CUDA-vs-oracle parity sweep (development aid): random synthetic datasets and random FilterParam
combinations, all three tile kernels, device and host ingest. usage: python tools/fuzz_parity.py [n_rounds] [seed0]"""
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_backend  # noqa: E402
from raer_b200 import FilterParam, api, synth  # noqa: E402
from test_synth_parity import _assert_same  # noqa: E402


def random_param(rng, n_bams):
    kw = dict(library_type=rng.choice(["unstranded", "fr-first-strand", "fr-second-strand"]))
    if rng.random() < 0.5:
        kw["only_keep_variants"] = [rng.random() < 0.5 for _ in range(n_bams)]
    if rng.random() < 0.4:
        kw["min_base_quality"] = rng.choice([0, 5, 13, 30])
    if rng.random() < 0.3:
        kw["min_mapq"] = [rng.choice([0, 3, 255]) for _ in range(n_bams)]
    if rng.random() < 0.3:
        kw["min_depth"] = rng.choice([0, 2, 5])
    if rng.random() < 0.3:
        kw["trim_5p"], kw["trim_3p"] = rng.randrange(0, 8), rng.randrange(0, 8)
    if rng.random() < 0.2:
        kw["ftrim_5p"], kw["ftrim_3p"] = rng.choice([0.0, 0.05, 0.2]), rng.choice([0.0, 0.1])
    if rng.random() < 0.3:
        kw["splice_dist"] = rng.randrange(1, 8)
    if rng.random() < 0.3:
        kw["indel_dist"] = rng.randrange(1, 8)
    if rng.random() < 0.3:
        kw["min_splice_overhang"] = rng.choice([5, 12, 40])
    if rng.random() < 0.2:
        kw["max_mismatch_type"] = (rng.randrange(1, 3), rng.randrange(1, 4))
    if rng.random() < 0.2:
        kw["homopolymer_len"] = rng.choice([3, 5])
    if rng.random() < 0.2:
        kw["min_allelic_freq"] = rng.choice([0.01, 0.1])
    if rng.random() < 0.2:
        kw["report_multiallelic"] = False
    if rng.random() < 0.2:
        kw["remove_overlaps"] = False
    if rng.random() < 0.2:
        kw["min_variant_reads"] = 2
    if rng.random() < 0.15:
        kw["max_depth"] = rng.choice([8, 40, 200])  # within reach of the coverage: the device ingest hands the contig back
    if rng.random() < 0.15:
        kw["read_bqual"] = (rng.choice([0.1, 0.3]), float(rng.choice([15, 25])))
    return kw


def main():
    n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    bad = 0
    for rnd in range(n_rounds):
        rng = random.Random(seed0 + rnd)
        n_bams = rng.choice([1, 2, 2, 3])
        clen = rng.choice([20_000, 50_000, 90_000])
        pairs = rng.choice([1500, 6000, 25_000])
        with tempfile.TemporaryDirectory() as d:
            ds = synth.make_bulk_dataset(d, n_bams=n_bams, n_contigs=2, contig_len=clen, pairs_per_contig=pairs, seed=seed0 * 100 + rnd)
            for k in range(4):
                kw = random_param(rng, n_bams)
                fp = FilterParam(**kw)
                want = oracle_backend.pileup_sites(ds["bams"], ds["fasta"], param=fp)
                for kern in ("default", "fast", "generic", "host-ingest"):
                    os.environ.pop("RAER_KERNEL", None)
                    os.environ.pop("RAER_GPU_INGEST", None)
                    if kern in ("fast", "generic"):
                        os.environ["RAER_KERNEL"] = kern
                    if kern == "host-ingest":
                        os.environ["RAER_GPU_INGEST"] = "0"
                    try:
                        got = api.pileup_sites(ds["bams"], ds["fasta"], param=fp)
                        _assert_same(got, want)
                    except Exception as e:  # noqa: BLE001
                        bad += 1
                        print(f"FAIL round {rnd} bams {n_bams} len {clen} pairs {pairs} kernel {kern} param {kw}: {str(e)[:300]}")
                os.environ.pop("RAER_KERNEL", None)
                os.environ.pop("RAER_GPU_INGEST", None)
                print(f"round {rnd}.{k}: {len(want)} rows, {n_bams} BAMs, {kw}")
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

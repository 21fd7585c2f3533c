#!/usr/bin/env python
"""bench.py -- aligned bases/s of the bulk pileup hot path on B200, with roofline and CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.

Workload at N=1 = BASELINE.json configs[1]: "all-sites bulk pileup: 2 synthetic BAMs x 100M 2x100bp
reads on a 100 Mb reference, fr-first-strand, only_keep_variants" (20 contigs x 5 Mb,
5e7 proper pairs per BAM, 2e10 aligned bases). One step = one pass of the pileup over that input.
  value  : input resident in HBM (synthetic batches generated on the GPU), kernels only.
  e2e    : the same batches in pinned HOST memory handed to the C-ABI (raer_batch_pileup_host):
           H2D of reads + reference, kernels, D2H of the result rows, host-side ordered concatenation.
  roofline : the tile kernel (k_pileup_fast for this configuration), algorithmic bytes (SURVEY.md 8d constants) / CUDA-event launch time.
  cpu_baseline : the CPU oracle (port of the reference algorithm) on a bounded BAM sample, 1 core.
N>1: strong scaling for C2 / C4: the ONE workload is partitioned by genomic region over the ranks (whole contigs, or
tiles of contigs with their halo reads: raer_b200.multi.plan_units), no collective on the data path, barrier +
max-over-ranks timing; rank 0 checks the tiles' rows against the same contigs run whole. C3 (one BAM per GPU) and
C5 are independent per-GPU inputs: weak scaling.

`--impl reference` times the oracle port on the host cores (one process per contig, the way the
reference parallelises with BiocParallel) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned bases/sec pileup (bulk, all sites, 2 BAMs, fr-first-strand, only_keep_variants)"
UNIT = "aligned bases/s"

# The driver's line is C2 (the default). `--config c3|c4|c5` measure the other BASELINE.json configurations the same way
# (same JSON shape); their lines are committed under profiles/.
BULK = {
    "c2": dict(
        metric=METRIC, n_bams=2, contig_len=5_000_000, contigs=20, pairs=2_500_000,
        workload="C2 all-sites bulk pileup: 2 synthetic BAMs x 1e8 2x100bp reads, 100 Mb reference (20 x 5 Mb), "
                 "fr-first-strand, only_keep_variants",
        param=dict(library_type="fr-first-strand", only_keep_variants=True)),
    "c3": dict(
        metric="aligned bases/sec pileup (bulk, known sites: 1e7 A/T positions, one BAM per GPU, min_depth=2)",
        n_bams=1, contig_len=5_000_000, contigs=20, pairs=1_250_000,
        workload="C3 known-sites pileup: one synthetic BAM per GPU x 5e7 2x100bp records, 100 Mb reference (20 x 5 Mb), "
                 "1e7 sites (A positions of plus genes / T positions of minus genes), fr-first-strand, min_depth=2",
        param=dict(library_type="fr-first-strand", min_depth=2)),
    "c4": dict(
        metric="aligned bases/sec calc_AEI (fused 4x4 reduction over 1e6 Alu-like intervals, one BAM)",
        n_bams=1, contig_len=40_000_000, contigs=25, pairs=4_000_000,
        workload="C4 calc_AEI: one synthetic BAM x 2e8 2x100bp records on a 1 Gb reference (25 x 40 Mb), 70 % of the "
                 "fragments at 1e6 non-overlapping intervals of 280-320 bases, fr-first-strand, fused AEI reduction",
        param=dict(library_type="fr-first-strand")),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configuration (SURVEY.md 8d); c2 is the driver's line")
    ap.add_argument("--scale", type=float, default=float(os.environ.get("RAER_BENCH_SCALE", "1.0")),
                    help="fraction of the contigs of the workload (1.0 = full config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-file-level", action="store_true", help="skip the BAM-file-level run of the public call on a larger sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU (NVML's ideal CPU affinity), so that its pinned host batches are
    allocated on that NUMA node and the H2D copies of the e2e leg do not cross the socket link. Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sample_dataset(outdir, n_contigs, seed=4242, config="c2"):
    """Bounded sample of the workload as real BAM files: same depth per BAM, same read model. Returns the paths plus
    what the call of that configuration needs (sites / intervals)."""
    from raer_b200 import synth
    from raer_b200.glue import Sites

    if config == "c4":
        ds = synth.make_aei_dataset(outdir, n_contigs=n_contigs, contig_len=250_000, pairs_per_contig=25_000, seed=seed)
        ds["bams"] = [ds["bam"]]
        iv = ds["intervals"]
        ds["alu"] = Sites(iv["seqnames"], iv["start"], iv["end"])
        return ds
    n_bams = BULK[config]["n_bams"]
    # 100x per BAM for C2, 50x for the single BAM of C3 (5e7 records on 100 Mb)
    pairs = 125_000 if config == "c2" else 62_500
    ds = synth.make_bulk_dataset(outdir, n_bams=n_bams, n_contigs=n_contigs, contig_len=250_000, pairs_per_contig=pairs, seed=seed)
    if config == "c3":
        refs = synth.make_reference(ds["lens"], seed)
        sn, sp = [], []
        for t, nm in enumerate(ds["names"]):
            pos = synth.known_sites_mask(refs[t], seed + 7 + t).nonzero().flatten().tolist()
            sn += [nm] * len(pos)
            sp += [x + 1 for x in pos]
        ds["sites"] = Sites(sn, sp)
    return ds


C2_PARAM = BULK["c2"]["param"]


def _call_kwargs(ds, config, contig=None):
    """arguments of the public call of a configuration on a sample dataset (optionally one contig of it)"""
    kw = {}
    if config == "c3":
        st = ds["sites"]
        kw["sites"] = st if contig is None else st.subset([s == contig for s in st.seqnames])
    return kw


def _oracle_contig(job):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    from raer_b200 import FilterParam
    from raer_b200.glue import Sites

    bams, fasta, contig, config, extra = job
    fp = FilterParam(**BULK[config]["param"])
    if config == "c4":
        alu = Sites(*extra)
        alu = alu.subset([s == contig for s in alu.seqnames])
        r = oracle_backend.calc_AEI(bams, fasta, alu_ranges=alu, param=fp)
        return sum(a + b for a, b in r["AEI_per_chrom"][os.path.basename(bams[0])].values())
    if config == "c3":
        st = Sites(*extra)
        r = oracle_backend.pileup_sites(bams, fasta, sites=st.subset([s == contig for s in st.seqnames]), region=contig, param=fp)
        return len(r)
    r = oracle_backend.pileup_sites(bams, fasta, region=contig, param=fp)
    return len(r)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "c5":
        return run_reference_sc(args)
    import multiprocessing as mp

    config = args.config
    cores = os.cpu_count() or 1
    n_contigs = max(2, min(cores, 20))
    with tempfile.TemporaryDirectory() as d:
        ds = sample_dataset(d, n_contigs, config=config)
        extra = None
        if config == "c3":
            extra = (ds["sites"].seqnames, ds["sites"].start)
        if config == "c4":
            extra = (ds["alu"].seqnames, ds["alu"].start, ds["alu"].end)
        jobs = [(ds["bams"], ds["fasta"], c, config, extra) for c in ds["names"]]
        workers = min(n_contigs, cores)
        with mp.get_context("spawn").Pool(workers) as pool:
            for _ in range(max(args.warmup, 0)):
                pool.map(_oracle_contig, jobs[:workers])  # bounded warm-up: page cache + imports
            t0 = time.perf_counter()
            rows = 0
            for _ in range(args.steps):
                rows = sum(pool.map(_oracle_contig, jobs))
            dt = (time.perf_counter() - t0) / args.steps
    val = ds["aligned_bases"] / dt
    sample = (f"{len(ds['bams'])} BAM(s) x {n_contigs} contigs x 250 kb ({ds['aligned_bases']:.3g} aligned bases, "
              f"{config.upper()} read model), one process per contig like BiocParallel; {rows} rows / counted bases")
    out = {
        "impl": "reference", "metric": BULK[config]["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": BULK[config]["workload"] + " -- bounded sample (CPU oracle port of raer; R/htslib absent)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    if args.config == "c5":
        return run_ours_sc(args)
    import numpy as np
    import torch
    import torch.distributed as dist

    from raer_b200 import _cabi, synth

    config = args.config
    cfg = BULK[config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: raer_b200 has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = _cabi.lib()

    n_contigs = max(1, int(round(cfg["contigs"] * args.scale)))
    contig_len = cfg["contig_len"]
    pairs = cfg["pairs"]  # per contig per BAM
    n_bams = cfg["n_bams"]
    ctx = L.raer_ctx_create(local)
    if not ctx:
        raise SystemExit(_cabi.last_error())
    stream = torch.cuda.ExternalStream(L.raer_ctx_stream(ctx), device=dev)

    def make_conf(mask_ptr=None):
        conf = _cabi.BulkConf()
        conf.n_files = n_bams
        conf.min_depth, conf.min_bq = (2 if config == "c3" else 1), 20
        conf.report_multiallelic, conf.remove_overlaps, conf.want_stats = 1, 1, 1
        if config == "c2":
            if os.environ.get("RAER_BENCH_NO_STATS"):  # development probe: skip the second pass (not the reference behaviour)
                conf.want_stats = 0
            if os.environ.get("RAER_BENCH_EXT_FILTERS"):  # development probe: a typical RNA-editing FilterParam (not the C2 call)
                conf.trim_5p, conf.trim_3p, conf.splice_dist, conf.indel_dist, conf.min_overhang = 5, 5, 4, 4, 10
            if os.environ.get("RAER_BENCH_NO_OVERLAP"):  # development probe: skip the mate-overlap kernel
                conf.remove_overlaps = 0
        for f in range(n_bams):
            conf.min_mapq[f] = 0
            conf.only_keep_variants[f] = 1 if config == "c2" else 0
        if config == "c2" and (os.environ.get("RAER_BENCH_ALL_SITES") or os.environ.get("RAER_BENCH_AEI")):
            for f in range(n_bams):  # development probes: a row for every covered site / the fused reduction on the C2 reads
                conf.only_keep_variants[f] = 0
            conf.aei_mode = 1 if os.environ.get("RAER_BENCH_AEI") else 0
        if config == "c4":
            conf.aei_mode = 1
        conf.reg_beg, conf.reg_end = 0, 2**31 - 1
        if mask_ptr is not None:
            conf.site_mask = mask_ptr
            conf.mask_beg, conf.mask_bits = 0, contig_len
        return conf

    # ---- work units. One GPU: the contigs of the workload. Several GPUs: C2 and C4 are ONE workload partitioned by
    # genomic region (raer_b200.multi.plan_units: whole contigs, or tiles of contigs with their halo reads when whole
    # contigs would leave ranks idle) -- strong scaling; C3 is one BAM per GPU by definition (independent calls) -- weak.
    from raer_b200 import multi

    partitioned = world > 1 and config in ("c2", "c4")
    if partitioned:
        names = [str(t) for t in range(n_contigs)]
        plan = multi.plan_units(names, {nm: contig_len for nm in names}, world)
        units = [(int(c), b, e) for _, c, b, e in plan[rank]]
    else:
        plan = None
        units = [(t, None, None) for t in range(n_contigs)]

    # ---- generate the workload in HBM
    t_gen = time.perf_counter()
    batches = []
    seed0 = {"c2": 1001, "c3": 3001, "c4": 4001}[config] + (0 if partitioned else 100_000 * rank)
    total_bases = 0
    n_sites = 0
    cache = {}
    keep_full = {}  # rank 0 of a partitioned run: the whole-contig batches of its contigs, for the partition check

    def contig_data(t):
        if cache.get("t") == t:
            return cache["v"]
        ref = synth.make_reference([contig_len], seed0 + t, device=dev)[0]
        starts = [None] * n_bams
        mask = None
        ns = 0
        if config == "c3":
            m = synth.known_sites_mask(ref, seed0 + 7 + t)
            ns = int(m.sum())
            mask = synth.pack_bitmap(m)
            del m
        if config == "c4":
            st, en = synth.alu_intervals(contig_len, seed0 + 10 + t, device=dev)
            m = synth.interval_mask(contig_len, st, en)
            ns = int(m.sum())
            mask = synth.pack_bitmap(m)
            starts = [synth.interval_starts(pairs, contig_len, st, en, seed0 + 100 + t)]
            del m
        per = [synth.gen_contig(ref, pairs, seed0 + 1000 * (b + 1) + t, device=dev, starts=starts[b]) for b in range(n_bams)]
        cache["t"], cache["v"] = t, (ref, per, mask, ns)
        return cache["v"]

    for t, ub, ue in units:
        ref, per, mask, ns = contig_data(t)
        lo, hi = (0, contig_len) if ub is None else (ub, ue)
        for r in per:
            op, ln = r.cigar & 0xF, r.cigar >> 4
            own = (r.pos >= lo) & (r.pos < hi)  # every read counts once, with the unit it starts in
            total_bases += int(((ln * (op == 0)).sum(1) * own).sum())
        n_sites += ns if ub is None else ns * (hi - lo) // contig_len
        sub = per if ub is None else [r.window(lo, hi) for r in per]
        T, rb = synth.pack_device_batch(sub, t, lo, hi)
        T["ref"] = synth.reference_ascii(ref)
        T["qual0"] = T["qual"].clone()
        if mask is not None:
            T["mask"] = mask
        batches.append((T, rb, make_conf(T["mask"].data_ptr() if mask is not None else None)))
        if partitioned and rank == 0 and ub is not None and t not in keep_full and not args.no_e2e:
            Tf, rbf = synth.pack_device_batch(per, t, 0, contig_len)
            Tf["ref"] = T["ref"]
            Tf["qual0"] = Tf["qual"]
            if mask is not None:
                Tf["mask"] = mask
            keep_full[t] = (Tf, rbf)
    cache.clear()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    launches = C.c_int64(0)
    rows_total = [0]

    def step_device():
        rows = 0
        with torch.cuda.stream(stream):
            for T, rb, conf in batches:
                T["qual"].copy_(T["qual0"])  # fresh qualities, stands in for the H2D of a new batch
                L.raer_ctx_set_reference_device(ctx, rb.tid, T["ref"].data_ptr(), T["ref"].numel())
                nr = C.c_int64(0)
                rc = L.raer_batch_pileup_device(ctx, C.byref(conf), C.byref(rb), C.byref(nr), C.byref(launches))
                if rc != 0:
                    raise SystemExit(f"raer_batch_pileup_device failed ({rc}): {_cabi.last_error()}")
                rows += nr.value
        rows_total[0] = rows

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    L.raer_ctx_profile(ctx, 1)
    launches.value = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    k_ms, k_n = C.c_double(0), C.c_int64(0)
    L.raer_ctx_profile_read(ctx, C.byref(k_ms), C.byref(k_n))
    L.raer_ctx_profile(ctx, 0)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    bases_all = torch.tensor([float(total_bases)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bases_all)
    value = float(bases_all.item()) / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (the tile kernel), per launch
    kernel_name = "k_pileup_lane" if not os.environ.get("RAER_KERNEL") and not os.environ.get("RAER_BENCH_EXT_FILTERS") else \
        ("k_pileup_tiles" if os.environ.get("RAER_KERNEL") == "generic" else "k_pileup_fast")
    rows_per_launch = rows_total[0] / max(len(batches), 1)
    alg = [L.raer_batch_algorithmic_bytes(C.byref(rb), int(rows_per_launch), contig_len) for _, rb, _ in batches]
    alg_per_launch = sum(alg) / len(alg)
    launch_ms = k_ms.value / max(k_n.value, 1)
    achieved = alg_per_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel (not measured live)
        tj = json.load(open(os.path.join(ROOT, "profiles", f"ncu_traffic_{kernel_name}.json")))
        if config == "c2":  # the capture is of a C2 launch
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
            traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": launch_ms,
                "kernel_share_of_step": (k_ms.value / args.steps) / (ms / args.steps)}

    # ---- e2e: pinned host batches through the C-ABI (H2D + kernels + D2H rows)
    e2e = None
    partition_check = None
    if not args.no_e2e:
        # Pinned host copies of the batches. All ranks of a node share its RAM: keep this rank's share below 40 % of
        # what is available / world and, if that is less than the whole workload, cycle through the batches that fit
        # (same bytes over PCIe and same kernels per step; the JSON says how many distinct batches were used).
        keys = ["hdr", "maxend", "cigar", "seq", "qual0", "pair_b", "ref"] + (["mask"] if "mask" in batches[0][0] else [])
        per_batch = sum(T[k].numel() * T[k].element_size() for T, _, _ in batches[:1] for k in keys)
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        n_host = max(1, min(len(batches), int(0.4 * avail / max(local_world, 1) / max(per_batch, 1))))
        host = []
        h2d = 0
        for T, rb, _ in batches[:n_host]:
            H = {k: torch.empty(T[k].shape, dtype=T[k].dtype, pin_memory=True) for k in keys}
            for k in H:
                H[k].copy_(T[k])
            hb = _cabi.ReadBatch()
            C.memmove(C.byref(hb), C.byref(rb), C.sizeof(hb))
            hb.hdr, hb.maxend, hb.cigar = H["hdr"].data_ptr(), H["maxend"].data_ptr(), H["cigar"].data_ptr()
            hb.seq, hb.qual, hb.pair_b = H["seq"].data_ptr(), H["qual0"].data_ptr(), H["pair_b"].data_ptr()
            hb.file_off = T["file_off"].ctypes.data
            host.append((H, hb, make_conf(H["mask"].data_ptr() if "mask" in H else None)))
        host = [host[i % n_host] for i in range(len(batches))]
        h2d = sum(H[k].numel() * H[k].element_size() for H, _, _ in host for k in H)
        torch.cuda.synchronize()
        d2h = [0]

        pipe = L.raer_pipe_create(local)
        if not pipe:
            raise SystemExit(_cabi.last_error())

        import hashlib

        import numpy as np

        def rows_digest(rows, lo=None, hi=None):
            """(number of rows, sha1 over every row field) of the rows with lo <= pos < hi"""
            n = rows.n_rows
            if n == 0:
                return (0,) + (hashlib.sha1(b"").hexdigest(),) * 5
            pos = np.ctypeslib.as_array(rows.pos, (n,))
            a, b = (0, n) if lo is None else (int(np.searchsorted(pos, lo, "left")), int(np.searchsorted(pos, hi, "left")))
            parts = (pos[a:b], np.ctypeslib.as_array(rows.meta, (n,))[a:b], np.ctypeslib.as_array(rows.stats, (n, 3))[a:b],
                     np.ctypeslib.as_array(rows.counts, (n, n_bams, 8))[a:b], np.ctypeslib.as_array(rows.altcode, (n, n_bams))[a:b])
            return (b - a,) + tuple(hashlib.sha1(x.tobytes()).hexdigest() for x in parts)

        digests = []  # per batch of the last host step, in submission order

        def collect():
            rows = _cabi.Rows()
            rc = L.raer_pipe_collect(pipe, C.byref(rows))
            if rc != 0:
                raise SystemExit(f"raer_pipe_collect failed ({rc}): {_cabi.last_error()}")
            nb = rows.n_rows * (8 + 24 + n_bams * 36) + (n_bams * 128 if config == "c4" else 0)
            if partitioned and want_digests[0]:
                digests.append(rows_digest(rows))
            L.raer_rows_free(C.byref(rows))
            return nb

        want_digests = [False]

        def step_host():
            # two batches in flight: H2D of batch i + 1 overlaps the kernels of batch i and the rows of batch i - 1
            nbytes = 0
            inflight = 0
            digests.clear()
            for H, hb, conf in host:
                if inflight == 2:
                    nbytes += collect()
                    inflight -= 1
                rc = L.raer_pipe_submit(pipe, C.byref(conf), C.byref(hb), hb.tid, H["ref"].data_ptr(), H["ref"].numel())
                if rc != 0:
                    raise SystemExit(f"raer_pipe_submit failed ({rc}): {_cabi.last_error()}")
                inflight += 1
            while inflight:
                nbytes += collect()
                inflight -= 1
            d2h[0] = nbytes

        want_digests[0] = partitioned and n_host == len(batches)
        step_host()
        unit_digests = list(digests)
        want_digests[0] = False
        barrier()
        n_e2e = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_host()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        t_dt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
        # ---- partition check: the rows of the tiles, gathered on rank 0 in (contig, tile) order, against the rows of the
        # same contigs computed whole (what one GPU computes), for every contig rank 0 holds a tile of
        if partitioned:
            mine = {(t, ub, ue): d for (t, ub, ue), d in zip(units, unit_digests)} if unit_digests else {}
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(mine, gathered, dst=0)
            if rank == 0:
                allu = {}
                for g in gathered:
                    allu.update(g or {})
                checked = bad = 0
                bad_fields = set()
                for t, (Tf, rbf) in keep_full.items():
                    Hf = {k: torch.empty(Tf[k].shape, dtype=Tf[k].dtype, pin_memory=True) for k in keys}
                    for k in Hf:
                        Hf[k].copy_(Tf[k])
                    hb = _cabi.ReadBatch()
                    C.memmove(C.byref(hb), C.byref(rbf), C.sizeof(hb))
                    hb.hdr, hb.maxend, hb.cigar = Hf["hdr"].data_ptr(), Hf["maxend"].data_ptr(), Hf["cigar"].data_ptr()
                    hb.seq, hb.qual, hb.pair_b = Hf["seq"].data_ptr(), Hf["qual0"].data_ptr(), Hf["pair_b"].data_ptr()
                    hb.file_off = Tf["file_off"].ctypes.data
                    conf = make_conf(Hf["mask"].data_ptr() if "mask" in Hf else None)
                    rc = L.raer_pipe_submit(pipe, C.byref(conf), C.byref(hb), hb.tid, Hf["ref"].data_ptr(), Hf["ref"].numel())
                    if rc != 0:
                        raise SystemExit(f"raer_pipe_submit failed ({rc}): {_cabi.last_error()}")
                    rows = _cabi.Rows()
                    if L.raer_pipe_collect(pipe, C.byref(rows)) != 0:
                        raise SystemExit(_cabi.last_error())
                    for (ut, ub, ue), d in sorted(allu.items(), key=lambda kv: (kv[0][0], kv[0][1] or 0)):
                        if ut != t or ub is None:
                            continue
                        checked += 1
                        whole = rows_digest(rows, ub, ue)
                        if whole != tuple(d):
                            bad += 1
                            for nm, x, y in zip(("n_rows", "pos", "meta", "stats", "counts", "altcode"), whole, tuple(d)):
                                if x != y:
                                    bad_fields.add(nm)
                    L.raer_rows_free(C.byref(rows))
                    del Hf
                partition_check = {"tiles_checked": checked, "tiles_total": sum(len(p) for p in plan),
                                   "identical_to_single_gpu_rows": bool(checked > 0 and bad == 0) if checked else None,
                                   "tiles_differing": bad, "fields_differing": sorted(bad_fields),
                                   "how": "sha1 over pos/meta/stats/counts/altcode of every tile's rows vs the same positions of "
                                          "the contig run as one batch, for the contigs rank 0 holds a tile of"}
        e2e = {"value": float(bases_all.item()) / float(t_dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h[0]), "ms_per_step": float(t_dt.item()) * 1e3, "steps": n_e2e,
               "distinct_host_batches": n_host, "batches_per_step": len(batches), "cpus_bound_to_gpu_node": numa,
               "api": "raer_pipe_submit / raer_pipe_collect (C-ABI, pinned host read batches -> rows in host memory, "
                      "two batches in flight)"}
        L.raer_pipe_destroy(pipe)
        del host

    # ---- CPU baseline + file-level breakdown on a bounded BAM sample (rank 0, N=1 only)
    cpu = None
    breakdown = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_backend
        from raer_b200 import FilterParam, api

        with tempfile.TemporaryDirectory() as d:
            ds = sample_dataset(d, 4, config=config)
            fp = FilterParam(**cfg["param"])
            if config == "c4":
                ours = lambda: api.calc_AEI(ds["bam"], ds["fasta"], alu_ranges=ds["alu"], param=fp)  # noqa: E731
                theirs = lambda: oracle_backend.calc_AEI(ds["bam"], ds["fasta"], alu_ranges=ds["alu"], param=fp)  # noqa: E731
            else:
                kw = _call_kwargs(ds, config)
                ours = lambda: api.pileup_sites(ds["bams"], ds["fasta"], param=fp, **kw)  # noqa: E731
                theirs = lambda: oracle_backend.pileup_sites(ds["bams"], ds["fasta"], param=fp, **kw)  # noqa: E731
            t0 = time.perf_counter()
            want = theirs()
            dt_cpu = time.perf_counter() - t0
            ours()  # warm-up: first pinned allocations, page cache
            t0 = time.perf_counter()
            got = ours()
            dt_gpu = time.perf_counter() - t0
            tm = dict(_cabi.LAST_TIMING)
            if config == "c4":
                same = got["AEI_per_chrom"] == want["AEI_per_chrom"]
                n_out = sum(a + b for a, b in want["AEI_per_chrom"][os.path.basename(ds["bam"])].values())
                what_out = f"{n_out} counted bases in the 12 (REF, ALT) sums"
            else:
                same = got.identical(want, stats_rtol=1e-6)
                what_out = f"{len(want)} rows"
        sample = (f"{len(ds['bams'])} BAM(s) x 4 contigs x 250 kb ({ds['aligned_bases']:.3g} aligned bases), "
                  f"single-threaded oracle port incl. BGZF decode; {what_out}")
        cpu = {"value": ds["aligned_bases"] / dt_cpu, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
               "seconds": dt_cpu}
        breakdown = {"what": ("calc_AEI()" if config == "c4" else "pileup_sites()") + " -> raer_pileup() (.Call layer) on the same BAM sample",
                     "wall_s": dt_gpu,
                     "decode_s": tm.get("t_decode"), "pack_s": tm.get("t_pack"), "h2d_s": tm.get("t_h2d"),
                     "kernel_s": tm.get("t_kernel"), "d2h_s": tm.get("t_d2h"),
                     "aligned_bases_per_s": ds["aligned_bases"] / dt_gpu, "identical_to_oracle": bool(same)}

    # ---- the public call on BAM files, at a size where its fixed costs are amortised: device ingest (BGZF inflate, record
    # parsing, prefilter, packing and mate pairing as kernels) against the host ingest of the same library
    file_level = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_file_level and config in ("c2", "c3"):
        from raer_b200 import FilterParam, api, synth

        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            n_bams = cfg["n_bams"]
            pairs = 500_000 if config == "c2" else 250_000
            fds = synth.make_bulk_dataset(d, n_bams=n_bams, n_contigs=2, contig_len=1_000_000, pairs_per_contig=pairs, seed=991)
            t_write = time.perf_counter() - t0
            fp = FilterParam(**cfg["param"])
            kw = {}
            if config == "c3":
                from raer_b200.glue import Sites
                refs = synth.make_reference(fds["lens"], 991)
                sn, sp = [], []
                for t, nm in enumerate(fds["names"]):
                    pos = synth.known_sites_mask(refs[t], 991 + 7 + t).nonzero().flatten().tolist()
                    sn += [nm] * len(pos)
                    sp += [x + 1 for x in pos]
                kw["sites"] = Sites(sn, sp)
            res, wall, tms = {}, {}, {}
            for mode in ("1", "0"):
                os.environ["RAER_GPU_INGEST"] = mode
                api.pileup_sites(fds["bams"], fds["fasta"], param=fp, **kw)  # warm-up: page cache, first allocations
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    r = api.pileup_sites(fds["bams"], fds["fasta"], param=fp, **kw)
                    dt = time.perf_counter() - t0
                    if best is None or dt < best:
                        best, tms[mode] = dt, dict(_cabi.LAST_TIMING)
                res[mode], wall[mode] = r, best
            os.environ.pop("RAER_GPU_INGEST", None)
            size = sum(os.path.getsize(b) for b in fds["bams"])
        file_level = {
            "what": "pileup_sites() on BAM files (reads the files, no pre-packed batches), best of 3 after one warm-up call",
            "sample": f"{n_bams} BAM(s) x 2 contigs x 1 Mb, {size / 1e6:.0f} MB of BGZF, {fds['aligned_bases']:.3g} aligned bases "
                      f"(written in {t_write:.0f} s)",
            "value": fds["aligned_bases"] / wall["1"], "unit": UNIT, "wall_s": wall["1"],
            "c_call_s": tms["1"]["t_total"], "device_ingest_contigs": tms["1"]["n_device_contigs"],
            "host_fallback_contigs": tms["1"]["n_host_contigs"], "gpu_launches": tms["1"]["gpu_launches"],
            "host_ingest": {"value": fds["aligned_bases"] / wall["0"], "wall_s": wall["0"], "c_call_s": tms["0"]["t_total"],
                            "host_threads": os.cpu_count()},
            "rows": len(res["1"]), "identical_rows_both_ingests": bool(res["1"].identical(res["0"])),
        }

    L.raer_ctx_destroy(ctx)
    if rank == 0:
        out = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if partitioned else "weak",
            "vs_baseline": None, "dtype": "u8/int32 (fp64 statistics)", "data": "synthetic",
            "config": {"workload": cfg["workload"],
                       "partition": None if not partitioned else {
                           "what": "one workload cut by genomic region over the ranks, no data-path collective, rows concatenated "
                                   "in (contig, tile) order (raer_b200.multi.plan_units)",
                           "units_per_rank": [len(p) for p in plan],
                           "tiles_per_contig": max(1, sum(len(p) for p in plan) // max(n_contigs, 1))},
                       "contigs": n_contigs, "scale": args.scale, "aligned_bases_per_gpu": total_bases,
                       "rows_per_step": rows_total[0], "sites_per_gpu": n_sites or None, "input_bytes_gt_L2": True,
                       "l2_note": "each launch streams the reads of one contig (> 126 MB L2); qualities are rewritten from a "
                                  "pristine copy before every launch",
                       "generation_s": t_gen},
            "gpu_launches": int(launches.value), "clocks": clocks, "roofline": roofline,
        }
        if e2e:
            out["e2e"] = e2e
        if partition_check:
            out["partition_check"] = partition_check
        if cpu:
            out["cpu_baseline"] = cpu
        if breakdown:
            out["breakdown"] = breakdown
        if file_level:
            out["file_level"] = file_level
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ single cell (C5)
SC_METRIC = "aligned bases/sec pileup_cells (10x-shaped BAM, 1e4 barcodes, 1e6 sites, sparse nRef/nAlt per site x barcode)"
SC_WORKLOAD = ("C5 pileup_cells: synthetic 10x-shaped BAM, 4e8 single-end 91 bp reads on a 100 Mb reference, 1e4 cell barcodes "
               "(Zipf), UMI consensus, 1e6 sites, fr-second-strand, min_mapq=255, min_variant_reads=1")
SC_PARAM = dict(library_type="fr-second-strand", min_mapq=255, min_variant_reads=1)


def _sc_sample(outdir):
    from raer_b200 import synth
    from raer_b200.glue import Sites

    ds = synth.make_sc_dataset(outdir, n_contigs=2, contig_len=100_000, reads_per_contig=200_000, n_cb=500, n_sites=2000)
    ds["gr"] = Sites(**ds["sites"])
    return ds


def _sc_triplets(sce):
    import numpy as np

    if sce is None:
        return None
    o = np.lexsort((sce.col, sce.row))
    return (sce.shape, sce.row[o].tolist(), sce.col[o].tolist(), sce.nRef[o].tolist(), sce.nAlt[o].tolist())


def _oracle_sc_chrom(job):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    from raer_b200 import FilterParam
    from raer_b200.glue import Sites

    bam, sites, barcodes, chrom, out = job
    r = oracle_backend.pileup_cells(bam, Sites(**sites), barcodes, out, chroms=[chrom], param=FilterParam(**SC_PARAM))
    return 0 if r is None else int(r.nRef.sum() + r.nAlt.sum())


def run_reference_sc(args):
    import multiprocessing as mp
    from oracle import pyoracle  # noqa: F401  (bench's reference arm may execute oracle/)

    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as d:
        ds = _sc_sample(d)
        n_bases = pyoracle.count_aligned_bases(ds["bam"])[0]
        jobs = [(ds["bam"], ds["sites"], ds["barcodes"], c, os.path.join(d, f"o_{c}")) for c in ds["names"]]
        workers = min(len(jobs), cores)
        with mp.get_context("spawn").Pool(workers) as pool:
            for _ in range(max(args.warmup, 0)):
                pool.map(_oracle_sc_chrom, jobs)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                cnt = sum(pool.map(_oracle_sc_chrom, jobs))
            dt = (time.perf_counter() - t0) / args.steps
    val = n_bases / dt
    sample = (f"one 10x-shaped BAM, {len(jobs)} contigs x 100 kb, 4e5 reads ({n_bases:.3g} aligned bases), 500 barcodes, "
              f"4000 sites, one process per contig like BiocParallel; {cnt} counted reads")
    out = {
        "impl": "reference", "metric": SC_METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": SC_WORKLOAD + " -- bounded sample (CPU oracle port of raer; R/htslib absent)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def run_ours_sc(args):
    import torch
    import torch.distributed as dist

    from raer_b200 import _cabi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: raer_b200 has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = _cabi.lib()
    # 4e8 reads on 100 Mb = 80 windows of 1.25 Mb x 5e6 reads (same depth, 364x; a window is one batch)
    n_batches = max(1, int(round(80 * args.scale)))
    win_len, reads_per = 1_250_000, 5_000_000
    n_cb = 10_000
    ctx = L.raer_sc_ctx_create(local)
    if not ctx:
        raise SystemExit(_cabi.last_error())
    stream = torch.cuda.ExternalStream(L.raer_sc_ctx_stream(ctx), device=dev)

    def make_conf(S):
        conf = _cabi.ScConf()
        conf.min_bq, conf.min_mapq = 20, 255
        conf.remove_overlaps, conf.has_umi = 0, 1  # single-end reads: nothing to pair
        conf.n_barcodes, conf.n_umi = n_cb, 1 << 24
        conf.n_sites = S["site_pos"].numel()
        conf.site_pos, conf.site_strand = S["site_pos"].data_ptr(), S["site_strand"].data_ptr()
        conf.site_ref, conf.site_alt = S["site_ref"].data_ptr(), S["site_alt"].data_ptr()
        return conf

    t_gen = time.perf_counter()
    batches, total_bases, n_sites = [], 0, 0
    seed0 = 5001 + 100_000 * rank
    for t in range(n_batches):
        ref = synth.make_reference([win_len], seed0 + t, device=dev)[0]
        T, rb = synth.gen_sc_batch(ref, reads_per, seed0 + 1000 + t, n_cb=n_cb, tid=t)
        total_bases += T["aligned_bases"]
        n_sites += T["site_pos"].numel()
        del ref
        batches.append((T, rb, make_conf(T)))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    launches = C.c_int64(0)
    cells_total, tup_total = [0], [0]

    def step_device():
        nc_sum = nt_sum = 0
        for T, rb, conf in batches:
            nc, nt = C.c_int64(0), C.c_int64(0)
            rc = L.raer_sc_batch_device(ctx, C.byref(conf), C.byref(rb), C.byref(nc), C.byref(nt), C.byref(launches))
            if rc != 0:
                raise SystemExit(f"raer_sc_batch_device failed ({rc}): {_cabi.last_error()}")
            nc_sum += nc.value
            nt_sum += nt.value
        cells_total[0], tup_total[0] = nc_sum, nt_sum

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    L.raer_sc_ctx_profile(ctx, 1)
    launches.value = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    kms, kn = (C.c_double * 4)(), (C.c_int64 * 4)()
    L.raer_sc_ctx_profile_read(ctx, kms, kn)
    L.raer_sc_ctx_profile(ctx, 0)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    bases_all = torch.tensor([float(total_bases)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bases_all)
    value = float(bases_all.item()) / (ms_step * 1e-3)

    # ---- roofline: the kernel class that takes the largest share of the step, per batch
    names = ["k_sc_emit (tuple emission)", "k_rs_hist/scan/scatter (radix sort of the tuples, all passes)",
             "k_sc_umi/k_sc_cell/k_cp_* (consensus + cell reduction)", "k_mate_overlap*"]
    per_batch = [kms[i] / max(args.steps * len(batches), 1) for i in range(4)]
    dom = max(range(4), key=lambda i: per_batch[i])
    cells_per_batch = cells_total[0] / max(len(batches), 1)
    alg = sum(L.raer_sc_algorithmic_bytes(C.byref(rb), int(cells_per_batch)) for _, rb, _ in batches) / len(batches)
    peak, peak_src = measured_peaks()
    achieved = alg / (per_batch[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "launch_ms": per_batch[dom],
                "unit_of_launch": "one batch (5e6 reads); the sort is several launches, timed as one group",
                "ms_per_batch_by_class": {names[i]: per_batch[i] for i in range(4)},
                "kernel_share_of_step": sum(kms) / args.steps / (ms / args.steps)}

    # ---- e2e: pinned host batches through raer_sc_batch_host (H2D + kernels + D2H of the cells)
    e2e = None
    if not args.no_e2e:
        keys = ["hdr", "cigar", "seq", "qual", "cb", "umi", "site_pos", "site_strand", "site_ref", "site_alt"]
        per_b = sum(batches[0][0][k].numel() * batches[0][0][k].element_size() for k in keys)
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        n_host = max(1, min(len(batches), int(0.4 * avail / max(local_world, 1) / max(per_b, 1))))
        host = []
        for T, rb, _ in batches[:n_host]:
            H = {k: torch.empty(T[k].shape, dtype=T[k].dtype, pin_memory=True) for k in keys}
            for k in H:
                H[k].copy_(T[k])
            hb = _cabi.ReadBatch()
            C.memmove(C.byref(hb), C.byref(rb), C.sizeof(hb))
            hb.hdr, hb.cigar, hb.seq, hb.qual = H["hdr"].data_ptr(), H["cigar"].data_ptr(), H["seq"].data_ptr(), H["qual"].data_ptr()
            hb.cb, hb.umi, hb.maxend, hb.pair_b = H["cb"].data_ptr(), H["umi"].data_ptr(), None, None
            hb.file_off = T["file_off"].ctypes.data
            host.append((H, hb, make_conf(H)))
        host = [host[i % n_host] for i in range(len(batches))]
        h2d = sum(H[k].numel() * H[k].element_size() for H, _, _ in host for k in H)
        torch.cuda.synchronize()
        d2h = [0]

        def step_host():
            nb = 0
            for H, hb, conf in host:
                cells = _cabi.ScCells()
                rc = L.raer_sc_batch_host(ctx, C.byref(conf), C.byref(hb), C.byref(cells))
                if rc != 0:
                    raise SystemExit(f"raer_sc_batch_host failed ({rc}): {_cabi.last_error()}")
                nb += cells.n * 16 + cells.n_sites * 8
                L.raer_sc_cells_free(C.byref(cells))
            d2h[0] = nb

        step_host()  # warm-up
        barrier()
        n_e2e = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_host()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        t_dt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(bases_all.item()) / float(t_dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h[0]), "ms_per_step": float(t_dt.item()) * 1e3, "steps": n_e2e,
               "distinct_host_batches": n_host, "batches_per_step": len(batches), "cpus_bound_to_gpu_node": numa,
               "api": "raer_sc_batch_host (C-ABI, pinned host read batches + sites -> nonzero cells in host memory)"}
        del host

    cpu = breakdown = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_backend
        from oracle import pyoracle
        from raer_b200 import FilterParam, api

        with tempfile.TemporaryDirectory() as d:
            ds = _sc_sample(d)
            n_bases = pyoracle.count_aligned_bases(ds["bam"])[0]
            fp = FilterParam(**SC_PARAM)
            t0 = time.perf_counter()
            want = oracle_backend.pileup_cells(ds["bam"], ds["gr"], ds["barcodes"], os.path.join(d, "o"), param=fp)
            dt_cpu = time.perf_counter() - t0
            api.pileup_cells_triplets(ds["bam"], ds["gr"], ds["barcodes"], param=fp)
            t0 = time.perf_counter()
            got = api.pileup_cells_triplets(ds["bam"], ds["gr"], ds["barcodes"], param=fp)
            dt_gpu = time.perf_counter() - t0
            same = _sc_triplets(got) == _sc_triplets(want)
        sample = (f"one 10x-shaped BAM, 2 contigs x 100 kb, 4e5 reads ({n_bases:.3g} aligned bases), 500 barcodes, 4000 sites, "
                  f"single-threaded oracle port incl. BGZF decode; {0 if want is None else len(want.row)} nonzero cells")
        cpu = {"value": n_bases / dt_cpu, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "seconds": dt_cpu}
        breakdown = {"what": "pileup_cells_triplets() -> raer_scpileup_triplets() (.Call layer) on the same BAM sample",
                     "wall_s": dt_gpu, "aligned_bases_per_s": n_bases / dt_gpu, "identical_to_oracle": bool(same)}

    # ---- the public call on a larger 10x-shaped BAM: device ingest against the host ingest of the same library
    file_level = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_file_level:
        from oracle import pyoracle
        from raer_b200 import FilterParam, api, synth
        from raer_b200.glue import Sites

        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            fds = synth.make_sc_dataset(d, n_contigs=2, contig_len=500_000, reads_per_contig=1_000_000, n_cb=2000, n_sites=20_000)
            t_write = time.perf_counter() - t0
            gr = Sites(**fds["sites"])
            fl_bases = pyoracle.count_aligned_bases(fds["bam"])[0]
            fp = FilterParam(**SC_PARAM)
            res, wall, devc = {}, {}, {}
            for mode in ("1", "0"):
                os.environ["RAER_GPU_INGEST"] = mode
                api.pileup_cells_triplets(fds["bam"], gr, fds["barcodes"], param=fp)
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    r = api.pileup_cells_triplets(fds["bam"], gr, fds["barcodes"], param=fp)
                    dt = time.perf_counter() - t0
                    best = dt if best is None or dt < best else best
                res[mode], wall[mode], devc[mode] = r, best, _cabi.sc_last_device_contigs()
            os.environ.pop("RAER_GPU_INGEST", None)
            size = os.path.getsize(fds["bam"])
        file_level = {
            "what": "pileup_cells_triplets() on a BAM file, one call per chromosome like the reference, best of 3 after a warm-up",
            "sample": f"one 10x-shaped BAM, 2 contigs x 500 kb, 2e6 reads, {size / 1e6:.0f} MB of BGZF, {fl_bases:.3g} aligned bases, "
                      f"2000 barcodes, {len(gr)} sites (written in {t_write:.0f} s)",
            "value": fl_bases / wall["1"], "unit": UNIT, "wall_s": wall["1"], "device_ingest_contigs_last_call": devc["1"],
            "host_ingest": {"value": fl_bases / wall["0"], "wall_s": wall["0"]},
            "identical_cells_both_ingests": bool(_sc_triplets(res["1"]) == _sc_triplets(res["0"])),
        }

    L.raer_sc_ctx_destroy(ctx)
    if rank == 0:
        out = {
            "metric": SC_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32/u64 keys", "data": "synthetic",
            "config": {"workload": SC_WORKLOAD, "batches": n_batches, "scale": args.scale, "aligned_bases_per_gpu": total_bases,
                       "reads_per_gpu": n_batches * reads_per, "sites_per_gpu": n_sites, "tuples_per_step": tup_total[0],
                       "cells_per_step": cells_total[0], "input_bytes_gt_L2": True,
                       "l2_note": "each batch streams ~1 GB of reads (> 126 MB L2)", "generation_s": t_gen},
            "gpu_launches": int(launches.value), "clocks": clocks, "roofline": roofline,
        }
        if e2e:
            out["e2e"] = e2e
        if cpu:
            out["cpu_baseline"] = cpu
        if breakdown:
            out["breakdown"] = breakdown
        if file_level:
            out["file_level"] = file_level
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

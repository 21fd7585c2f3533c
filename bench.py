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
N>1: weak scaling, every rank runs its own copy of the workload (own seeds, as if other samples);
no collective on the data path, barrier + max-over-ranks timing only.

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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("RAER_BENCH_SCALE", "1.0")),
                    help="fraction of the 20 contigs of the C2 workload (development only; 1.0 = full config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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


def sample_dataset(outdir, n_contigs, seed=4242):
    """Bounded sample of the C2 workload as real BAM files: same depth (100x per BAM), same read model."""
    from raer_b200 import synth

    return synth.make_bulk_dataset(outdir, n_bams=2, n_contigs=n_contigs, contig_len=250_000, pairs_per_contig=125_000,
                                   seed=seed)


C2_PARAM = dict(library_type="fr-first-strand", only_keep_variants=True)


def _oracle_contig(job):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    from raer_b200 import FilterParam

    bams, fasta, contig = job
    r = oracle_backend.pileup_sites(bams, fasta, region=contig, param=FilterParam(**C2_PARAM))
    return len(r)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    n_contigs = max(2, min(cores, 20))
    with tempfile.TemporaryDirectory() as d:
        ds = sample_dataset(d, n_contigs)
        jobs = [(ds["bams"], ds["fasta"], c) for c in ds["names"]]
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
    sample = (f"2 BAMs x {n_contigs} contigs x 250 kb at 100x per BAM ({ds['aligned_bases']:.3g} aligned bases, "
              f"C2 read model), one process per contig like BiocParallel; {rows} rows")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": "C2 all-sites bulk pileup, bounded sample (CPU oracle port of raer; R/htslib absent)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
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

    n_contigs = max(1, int(round(20 * args.scale)))
    contig_len = 5_000_000
    pairs = 2_500_000  # per contig per BAM: 5e7 pairs per BAM over 20 contigs
    n_bams = 2
    ctx = L.raer_ctx_create(local)
    if not ctx:
        raise SystemExit(_cabi.last_error())
    stream = torch.cuda.ExternalStream(L.raer_ctx_stream(ctx), device=dev)

    # ---- generate the workload in HBM
    t_gen = time.perf_counter()
    batches = []
    seed0 = 1001 + 100_000 * rank
    total_bases = 0
    for t in range(n_contigs):
        ref = synth.make_reference([contig_len], seed0 + t, device=dev)[0]
        per = [synth.gen_contig(ref, pairs, seed0 + 1000 * (b + 1) + t, device=dev) for b in range(n_bams)]
        for r in per:
            op, ln = r.cigar & 0xF, r.cigar >> 4
            total_bases += int((ln * (op == 0)).sum())
        T, rb = synth.pack_device_batch(per, t, 0, contig_len)
        T["ref"] = synth.reference_ascii(ref)
        T["qual0"] = T["qual"].clone()
        del per, ref
        batches.append((T, rb))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    conf = _cabi.BulkConf()
    conf.n_files = n_bams
    conf.min_depth, conf.min_bq = 1, 20
    conf.report_multiallelic, conf.remove_overlaps, conf.want_stats = 1, 1, 1
    if os.environ.get("RAER_BENCH_NO_STATS"):  # development probe: skip the second pass (not the reference behaviour)
        conf.want_stats = 0
    if os.environ.get("RAER_BENCH_EXT_FILTERS"):  # development probe: a typical RNA-editing FilterParam (not the C2 call)
        conf.trim_5p, conf.trim_3p, conf.splice_dist, conf.indel_dist, conf.min_overhang = 5, 5, 4, 4, 10
    if os.environ.get("RAER_BENCH_ALL_SITES"):  # development probe: calc_AEI-like call, a row for every covered site
        for f in range(n_bams):
            conf.only_keep_variants[f] = 0
    if os.environ.get("RAER_BENCH_AEI"):  # development probe: the same with the fused 4 x 4 reduction instead of rows
        for f in range(n_bams):
            conf.only_keep_variants[f] = 0
        conf.aei_mode = 1
    if os.environ.get("RAER_BENCH_NO_OVERLAP"):  # development probe: skip the mate-overlap kernel
        conf.remove_overlaps = 0
    for f in range(n_bams):
        conf.min_mapq[f] = 0
        conf.only_keep_variants[f] = 1
    conf.reg_beg, conf.reg_end = 0, 2**31 - 1

    launches = C.c_int64(0)
    rows_total = [0]

    def step_device():
        rows = 0
        with torch.cuda.stream(stream):
            for T, rb in batches:
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
    rows_per_launch = rows_total[0] / max(len(batches), 1)
    alg = [L.raer_batch_algorithmic_bytes(C.byref(rb), int(rows_per_launch), contig_len) for _, rb in batches]
    alg_per_launch = sum(alg) / len(alg)
    launch_ms = k_ms.value / max(k_n.value, 1)
    achieved = alg_per_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel (not measured live)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_k_pileup_fast.json")))
        traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
        traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_pileup_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": launch_ms,
                "kernel_share_of_step": (k_ms.value / args.steps) / (ms / args.steps)}

    # ---- e2e: pinned host batches through the C-ABI (H2D + kernels + D2H rows)
    e2e = None
    if not args.no_e2e:
        # Pinned host copies of the batches. All ranks of a node share its RAM: keep this rank's share below 40 % of
        # what is available / world and, if that is less than the whole workload, cycle through the batches that fit
        # (same bytes over PCIe and same kernels per step; the JSON says how many distinct batches were used).
        per_batch = sum(T[k].numel() * T[k].element_size() for T, _ in batches[:1] for k in
                        ("hdr", "maxend", "cigar", "seq", "qual0", "pair_b", "ref"))
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        n_host = max(1, min(len(batches), int(0.4 * avail / max(local_world, 1) / max(per_batch, 1))))
        host = []
        h2d = 0
        for T, rb in batches[:n_host]:
            H = {k: torch.empty(T[k].shape, dtype=T[k].dtype, pin_memory=True) for k in
                 ("hdr", "maxend", "cigar", "seq", "qual0", "pair_b", "ref")}
            for k in H:
                H[k].copy_(T[k])
            hb = _cabi.ReadBatch()
            C.memmove(C.byref(hb), C.byref(rb), C.sizeof(hb))
            hb.hdr, hb.maxend, hb.cigar = H["hdr"].data_ptr(), H["maxend"].data_ptr(), H["cigar"].data_ptr()
            hb.seq, hb.qual, hb.pair_b = H["seq"].data_ptr(), H["qual0"].data_ptr(), H["pair_b"].data_ptr()
            hb.file_off = T["file_off"].ctypes.data
            host.append((H, hb))
        host = [host[i % n_host] for i in range(len(batches))]
        h2d = sum(H[k].numel() * H[k].element_size() for H, _ in host for k in H)
        torch.cuda.synchronize()
        d2h = [0]

        pipe = L.raer_pipe_create(local)
        if not pipe:
            raise SystemExit(_cabi.last_error())

        def collect():
            rows = _cabi.Rows()
            rc = L.raer_pipe_collect(pipe, C.byref(rows))
            if rc != 0:
                raise SystemExit(f"raer_pipe_collect failed ({rc}): {_cabi.last_error()}")
            nb = rows.n_rows * (8 + 24 + n_bams * 36)
            L.raer_rows_free(C.byref(rows))
            return nb

        def step_host():
            # two batches in flight: H2D of batch i + 1 overlaps the kernels of batch i and the rows of batch i - 1
            nbytes = 0
            inflight = 0
            for H, hb in host:
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

        step_host()
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
            ds = sample_dataset(d, 4)
            fp = FilterParam(**C2_PARAM)
            t0 = time.perf_counter()
            want = oracle_backend.pileup_sites(ds["bams"], ds["fasta"], param=fp)
            dt_cpu = time.perf_counter() - t0
            api.pileup_sites(ds["bams"], ds["fasta"], param=fp)  # warm-up: first pinned allocations, page cache
            t0 = time.perf_counter()
            got = api.pileup_sites(ds["bams"], ds["fasta"], param=fp)
            dt_gpu = time.perf_counter() - t0
            tm = dict(_cabi.LAST_TIMING)
            same = got.identical(want, stats_rtol=1e-6)
        sample = (f"2 BAMs x 4 contigs x 250 kb at 100x per BAM ({ds['aligned_bases']:.3g} aligned bases), "
                  f"single-threaded oracle port incl. BGZF decode; {len(want)} rows")
        cpu = {"value": ds["aligned_bases"] / dt_cpu, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
               "seconds": dt_cpu}
        breakdown = {"what": "raer_pileup() (.Call layer) on the same BAM sample", "wall_s": dt_gpu,
                     "decode_s": tm.get("t_decode"), "pack_s": tm.get("t_pack"), "h2d_s": tm.get("t_h2d"),
                     "kernel_s": tm.get("t_kernel"), "d2h_s": tm.get("t_d2h"),
                     "aligned_bases_per_s": ds["aligned_bases"] / dt_gpu, "identical_to_oracle": bool(same)}

    L.raer_ctx_destroy(ctx)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32 (fp64 statistics)", "data": "synthetic",
            "config": {"workload": "C2 all-sites bulk pileup: 2 synthetic BAMs x 1e8 2x100bp reads, 100 Mb reference "
                                   "(20 x 5 Mb), fr-first-strand, only_keep_variants",
                       "contigs": n_contigs, "scale": args.scale, "aligned_bases_per_gpu": total_bases,
                       "rows_per_step": rows_total[0], "input_bytes_gt_L2": True,
                       "l2_note": "each launch streams ~1.9 GB of reads (> 126 MB L2); qualities are rewritten from a "
                                  "pristine copy before every launch",
                       "generation_s": t_gen},
            "gpu_launches": int(launches.value), "clocks": clocks, "roofline": roofline,
        }
        if e2e:
            out["e2e"] = e2e
        if cpu:
            out["cpu_baseline"] = cpu
        if breakdown:
            out["breakdown"] = breakdown
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

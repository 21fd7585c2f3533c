#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box: gpurun -- 'bash tools/profile_r2.sh'). Every profiled command is first
# run plain and has to exit 0; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --scale 0.1"
NCU="ncu --set full --clock-control none --import-source on -f"
$B > $O/p_plain_c2.json 2> $O/p_plain_c2.err || { echo "plain c2 run failed"; exit 1; }
# launch list of the default command at reduced contig count (same per-launch sizes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --scale 0.1 > $O/p_ll.json 2> $O/p_ll.err
$NCU -k k_pileup_lane -s 3 -c 1 -o $O/r2_k_pileup_lane $B > $O/p1.log 2>&1
$NCU -k regex:k_mate_overlap -s 6 -c 2 -o $O/r2_k_mate_overlap $B > $O/p2.log 2>&1
RAER_KERNEL=fast $B > $O/p_plain_fast.json 2>&1 && RAER_KERNEL=fast $NCU -k regex:k_pileup_fast -s 3 -c 1 -o $O/r2_k_pileup_fast $B > $O/p3a.log 2>&1
RAER_KERNEL=fast RAER_BENCH_EXT_FILTERS=1 $B > $O/p_plain_ext.json 2>&1 && \
    RAER_KERNEL=fast RAER_BENCH_EXT_FILTERS=1 $NCU -k regex:k_pileup_fast -s 3 -c 1 -o $O/r2_k_pileup_fast_ext $B > $O/p3.log 2>&1
RAER_KERNEL=generic $B > $O/p_plain_generic.json 2>&1 && RAER_KERNEL=generic $NCU -k k_pileup_tiles -s 1 -c 1 -o $O/r2_k_pileup_tiles $B > $O/p4.log 2>&1
S="python bench.py --config c5 --steps 1 --warmup 3 --no-cpu --no-e2e --scale 0.05"
$S > $O/p_plain_c5.json 2> $O/p_plain_c5.err && {
    $NCU -k k_sc_emit -s 2 -c 1 -o $O/r2_k_sc_emit $S > $O/p5.log 2>&1
    $NCU -k k_rs_scatter -s 8 -c 1 -o $O/r2_k_rs_scatter $S > $O/p6.log 2>&1
}
python tools/inflate_bench.py > $O/p_plain_inflate.log 2>&1 && $NCU -k k_gi_inflate -s 1 -c 1 -o $O/r2_k_gi_inflate python tools/inflate_bench.py > $O/p7.log 2>&1
ls -la $O/*.ncu-rep | tail -12

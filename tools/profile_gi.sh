#!/bin/bash
# ncu evidence for the device ingest (run on the GPU box): a full capture of k_gi_inflate and the launch list of one
# file-level pileup_sites() call with the device ingest
set -u
O=gpurun_out
mkdir -p $O
python tools/inflate_bench.py > $O/p_plain_inflate.log 2>&1 || { echo "plain inflate run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -f -k k_gi_inflate -s 1 -c 1 -o $O/r2_k_gi_inflate python tools/inflate_bench.py > $O/p7.log 2>&1
RAER_GPU_INGEST=1 python tools/file_level.py 250000 500000 2 2 > $O/p_plain_file_level.log 2>&1 || { echo "plain file-level run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches_file_level.csv \
    python tools/file_level.py 250000 500000 2 2 > $O/p8.log 2>&1
tail -3 $O/p_plain_file_level.log

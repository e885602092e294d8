#!/bin/bash
# ncu evidence for every kernel of the hot path (B200_PROFILING.md recipe).  Run on the GPU box:
#   bash tools/profile_all.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv
# Each capture follows a plain run of the same command that exited 0.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
BENCH="python bench.py --gb 1 --steps 2 --warmup 1 --no-e2e --no-cpu"
KT="python tools/kernel_times.py 256"
$BENCH > $OUT/${TAG}_plain_bench.json 2> $OUT/${TAG}_plain_bench.err || { echo "plain bench failed"; exit 1; }
$KT > $OUT/${TAG}_plain_kt.json 2> $OUT/${TAG}_plain_kt.err || { echo "plain kernel_times failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > /dev/null 2>&1
for spec in "gear_scan_tma_kernel:scan" "resolve_spec_kernel:resolve" "sha256_kernel:sha256" "parse_kernel:parse" "huffman_kernel:huffman" "encode_kernel:encode" "pack_kernel:pack" "dedup_insert_kernel:dedup"; do
  k=${spec%%:*}; n=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$n $BENCH > $OUT/${TAG}_ncu_$n.log 2>&1
done
for spec in "minhash_kernel:minhash" "lsh_keys_kernel:lshkeys"; do
  k=${spec%%:*}; n=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$n $KT > $OUT/${TAG}_ncu_$n.log 2>&1
done
ls -la $OUT/${TAG}_*

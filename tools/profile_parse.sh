#!/bin/bash
# Launch list + full ncu capture of every parse_kernel launch of one 1 GB step (all size classes), after a plain run.
#   bash tools/profile_parse.sh <tag>  -> gpurun_out/<tag>_launches.csv, <tag>_parse.ncu-rep
set -u
TAG=${1:-rXX}
OUT=gpurun_out
BENCH="python bench.py --gb 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --no-l4"
$BENCH > $OUT/${TAG}_plain_bench.json 2> $OUT/${TAG}_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^parse_kernel -s 4 -c 4 -f -o $OUT/${TAG}_parse $BENCH > $OUT/${TAG}_ncu_parse.log 2>&1
ls -la $OUT/${TAG}_*

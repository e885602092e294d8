#!/bin/bash
# Full ncu captures of the L4 kernels (MinHash, delta encode / apply, base selection) on a 1 GB batch, after a plain run.
#   bash tools/profile_l4.sh <tag>  -> gpurun_out/<tag>_{minhash,delta_encode,delta_apply,votes}.ncu-rep
set -u
TAG=${1:-rXX}
OUT=gpurun_out
CMD="python tools/l4_times.py 1 1"
$CMD > $OUT/${TAG}_plain_l4.json 2> $OUT/${TAG}_plain_l4.err || { echo "plain l4 run failed"; exit 1; }
for spec in "minhash_kernel:minhash" "delta_encode_kernel:delta_encode" "delta_apply_kernel:delta_apply" "heads_kernel:heads" "sha256_kernel:sha256"; do
  k=${spec%%:*}; n=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/${TAG}_$n $CMD > $OUT/${TAG}_ncu_$n.log 2>&1
done
ls -la $OUT/${TAG}_*

#!/bin/bash
# A/B run of an experimental library variant (tools/build_variant.sh NAME ...) against the product library on one GPU box:
#   bash tools/ab_run.sh TAG NAME   ->  gpurun_out/TAG_*  (scan variants, deflate tests + parse phases for both libraries)
# The product library is put back whatever happens.
set -u
TAG=$1; NAME=$2
OUT=gpurun_out; mkdir -p $OUT
LIB=hmse_b200/libhmse_b200.so
cp $LIB /tmp/hmse_base.so
trap 'cp /tmp/hmse_base.so $LIB' EXIT
python tools/scan_variants.py 4 > $OUT/${TAG}_scan_variants.txt 2> $OUT/${TAG}_scan_variants.err; echo "scan rc=$?"; cat $OUT/${TAG}_scan_variants.txt
python tools/parse_phases.py > $OUT/${TAG}_phases_base.json 2> $OUT/${TAG}_phases_base.err; echo "phases base rc=$?"
cp hmse_b200/libhmse_b200_$NAME.so $LIB
python tools/parse_phases.py > $OUT/${TAG}_phases_$NAME.json 2> $OUT/${TAG}_phases_$NAME.err; echo "phases $NAME rc=$?"
python -m pytest tests/test_gpu_deflate.py -m gpu -x -q > $OUT/${TAG}_pytest_$NAME.log 2>&1; echo "pytest $NAME rc=$?"; tail -2 $OUT/${TAG}_pytest_$NAME.log
grep -h '"GB/s"\|cycles_per_byte\|P3' $OUT/${TAG}_phases_base.json $OUT/${TAG}_phases_$NAME.json

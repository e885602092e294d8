#!/bin/bash
# Checks the experimental gear-scan variants (tools/scan_check.py: equality with the default on ragged sizes, then time),
# picks the fastest one that is bit-equal and at least 3 % faster than the default, and runs the benchmark (with its
# full-size CPU-oracle verification) under it.   bash tools/scan_pick.sh TAG "4 3 5" [bench args]
set -u
TAG=$1; VARIANTS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_scan_check.txt
for v in $VARIANTS; do
  timeout -s KILL 40 python tools/scan_check.py $v 4 >> $OUT/${TAG}_scan_check.txt 2> $OUT/${TAG}_scan_check_v$v.err
  echo "variant $v rc=$?"
done
cat $OUT/${TAG}_scan_check.txt
PICK=$(python - <<'P' "$OUT/${TAG}_scan_check.txt"
import json, sys
best, bt = "", None
for l in open(sys.argv[1]):
    try:
        r = json.loads(l)
    except ValueError:
        continue
    if r.get("equal") and r["variant_ms"] < 0.97 * r["default_ms"] and (bt is None or r["variant_ms"] < bt):
        best, bt = r["variant"], r["variant_ms"]
print(best)
P
)
echo "picked variant: '${PICK}'" | tee $OUT/${TAG}_picked.txt
if [ -n "$PICK" ]; then export HMSE_SCAN_VARIANT=$PICK; fi
timeout -s KILL 75 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<'P' "$OUT/${TAG}_bench.json"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), "stages", d.get("stages_ms"), "oracle_equal", (d.get("verify") or {}).get("oracle_equal"))
except Exception as e:
    print("no bench line:", e)
P
timeout -s KILL 30 python -m pytest tests/test_gpu_cdc.py tests/test_gpu_stream.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.log

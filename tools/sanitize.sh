#!/bin/bash
# compute-sanitizer over the GPU parity tests on small inputs (SURVEY.md section 5 "race detection / sanitizers").
#   bash tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [tag]
# ONE tool per gpurun call (B200_PROFILING.md).  Writes gpurun_out/<tag>_sanitizer_<tool>.txt; the summary line
# ("ERROR SUMMARY: 0 errors") is what profiles/ keeps.  The plain run comes first: a test that fails without the tool is
# not worth sanitising.
set -u
TOOL=${1:-memcheck}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
# the parity tests run on <= 8 MiB inputs: kernels of microseconds, so the tool's 10-100x slowdown stays affordable
SEL="cdc or digest or dedup or deflate or inflate or delta or archive or golden"
TESTS="tests/test_gpu_cdc.py tests/test_gpu_digest.py tests/test_gpu_deflate.py tests/test_gpu_inflate.py tests/test_gpu_delta.py tests/test_gpu_archive.py tests/test_gpu_golden.py"
python -m pytest $TESTS -m gpu -x -q -k "$SEL" > $OUT/${TAG}_sanitizer_plain.txt 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_sanitizer_plain.txt; exit 1; }
tail -2 $OUT/${TAG}_sanitizer_plain.txt
EXTRA=""
[ "$TOOL" = "memcheck" ] && EXTRA="--leak-check no --report-api-errors no"
timeout 2400 compute-sanitizer --tool $TOOL $EXTRA --target-processes application-only --error-exitcode 99 --print-limit 50 \
    python -m pytest $TESTS -m gpu -x -q -k "$SEL" > $OUT/${TAG}_sanitizer_${TOOL}.txt 2>&1
RC=$?
echo "compute-sanitizer --tool $TOOL exit code $RC" >> $OUT/${TAG}_sanitizer_${TOOL}.txt
grep -E "ERROR SUMMARY|passed|failed|exit code" $OUT/${TAG}_sanitizer_${TOOL}.txt | tail -8
exit 0

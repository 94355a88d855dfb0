#!/bin/bash
# one sanitizer tool per gpurun call (the profiling guide's rule): $1 = memcheck | racecheck | synccheck | initcheck
T=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_subset.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $T --print-limit 20 python tools/sanitize_subset.py > gpurun_out/sanitize_$T.log 2>&1; echo "sanitizer $T rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize subset ok" gpurun_out/sanitize_$T.log | tail -20

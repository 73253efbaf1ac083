#!/bin/bash
# compute-sanitizer pass over every kernel family that exchanges state through shared memory /
# DSMEM / TMEM (tools/sanitize_targets.py).  Run on the GPU box:
#     bash tools/sanitize.sh [outdir]        # default gpurun_out/sanitizer
# One log per tool; the summaries are copied to profiles/r02_sanitizer_*.log.
set -u
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  echo "== $tool" | tee "$OUT/$tool.log"
  timeout 1500 "$SAN" --tool "$tool" $extra --print-limit 50 --error-exitcode 7 \
      python tools/sanitize_targets.py >> "$OUT/$tool.log" 2>&1
  echo "exit code: $?" | tee -a "$OUT/$tool.log"
  tail -n 4 "$OUT/$tool.log"
done

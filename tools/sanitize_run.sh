#!/bin/bash
# compute-sanitizer passes over one tiny detokenize (tools/sanitize_pass.py); logs under gpurun_out/sanitizer/.
# Usage (on the GPU box): bash tools/sanitize_run.sh [tools...]   default: memcheck racecheck synccheck
set -u
out=gpurun_out/sanitizer
mkdir -p "$out"
tools="${*:-memcheck racecheck synccheck}"
for t in $tools; do
  echo "== compute-sanitizer --tool $t" | tee "$out/$t.log"
  timeout 900 compute-sanitizer --tool "$t" --print-limit 20 --launch-timeout 0 \
      python tools/sanitize_pass.py 1 8 >> "$out/$t.log" 2>&1
  echo "exit code $?" >> "$out/$t.log"
  tail -n 6 "$out/$t.log"
done

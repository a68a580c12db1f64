#!/bin/bash
# Race hunting: N fresh processes, each builds the model and runs a few small passes (tools/stress_small.py).
# usage: cold_start_loop.sh <tag> <lib or ""> <n>
O=gpurun_out/cold; mkdir -p $O
for i in $(seq 1 $3); do
  if [ -n "$2" ]; then export SPARKCODEC_LIB=$2; fi
  timeout 120 python tools/stress_small.py 2 > $O/$1_$i.txt 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then echo "$1 run $i rc=$rc"; grep "timed out" $O/$1_$i.txt | sort | uniq -c | sort -rn | head -40; fi
done
echo "$1: done $3 runs"

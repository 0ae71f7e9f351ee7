#!/bin/bash
# Round-2 GPU session 15: item-level trace of the persistent attention forward (why is it slower at 4096 keys?).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s15_summary.txt
: > $S
for nk in 4096 1024 320; do
  timeout -k 10 200 python tools/attn_itrace.py $nk >> $S 2>&1
done
cat $S

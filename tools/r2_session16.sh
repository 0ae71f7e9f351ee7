#!/bin/bash
# Round-2 GPU session 16: persistent attention forward with tile B staggered behind tile A at every item start.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s16_summary.txt
: > $S
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/r2s16_attn.log 2>&1
echo "attention tests rc=$? $(tail -1 gpurun_out/r2s16_attn.log)" >> $S
for v in 2 4 5; do
  echo "--- attn_bench variant $v" >> $S
  PXA_ATTN_VARIANT=$v timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
done
for v in 4 5; do
  timeout -k 10 200 python tools/attn_itrace.py 4096 $v 2>&1 | cut -c1-420 >> $S
done
timeout -k 10 200 python tools/attn_itrace.py 320 4 2>&1 | cut -c1-420 >> $S
cat $S

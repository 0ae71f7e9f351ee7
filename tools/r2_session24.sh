#!/bin/bash
# Round-2 GPU session 24: how much of the attention epilogue is the output stores?  (nostore = experiment build that skips them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s24_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
for rep in 1 2; do
echo "--- default" >> $S
timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
echo "--- nostore" >> $S
PXA_LIB_PATH=$V/libpixart_sm100_nostore.so timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
done
PXA_LIB_PATH=$V/libpixart_sm100_nostore.so timeout -k 10 200 python tools/attn_itrace.py 320 2>&1 | tail -4 | cut -c1-400 >> $S
cat $S

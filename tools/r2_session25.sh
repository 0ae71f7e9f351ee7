#!/bin/bash
# Round-2 GPU session 25: attention forward with 2 instead of 3 K/V stages (would free 52 KB of smem for an output staging buffer).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s25_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
for rep in 1 2; do
echo "--- default (3 stages)" >> $S
timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
echo "--- kv2 (2 stages)" >> $S
PXA_LIB_PATH=$V/libpixart_sm100_kv2.so timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
done
PXA_LIB_PATH=$V/libpixart_sm100_kv2.so timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/r2s25_t.log 2>&1
echo "kv2 attention tests rc=$? $(tail -1 gpurun_out/r2s25_t.log)" >> $S
cat $S

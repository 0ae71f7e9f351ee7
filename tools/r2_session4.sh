#!/bin/bash
# Round-2 GPU session 4: FMA-pipe exp2 share in the attention forward (PXA_POLY_OF8 = 0..4): parity + isolated timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s4_summary.txt
: > $S
timeout 60 tools/micro/exp_mix 2>&1 | head -2 >> $S
for v in default poly1 poly2 poly3 poly4; do
  if [ "$v" == "default" ]; then unset PXA_LIB_PATH; else export PXA_LIB_PATH=$PWD/pixart_sigma_b200/build/variants/libpixart_sm100_$v.so; fi
  echo "--- $v" >> $S
  timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_benchshape_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" 2>&1 | tail -1 >> $S
  timeout 200 python tools/attn_bwd_bench.py 8 16 4096 2>&1 | tail -1 >> $S
  timeout 200 python tools/attn_bwd_bench.py 2 16 16384 2>&1 | tail -1 >> $S
done
unset PXA_LIB_PATH
grep "bench-geometry flash_attn" gpurun_out/parity.txt | tail -12 >> $S
cat $S

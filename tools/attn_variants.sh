#!/bin/bash
# Time the attention-kernel experiment builds (pixart_sigma_b200/build/variants/libpixart_sm100_<tag>.so) in one GPU call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/attn_variants.txt
for lib in pixart_sigma_b200/build/variants/libpixart_sm100_*.so; do
  v=$(basename $lib .so); v=${v#libpixart_sm100_}
  echo "=== $v" >> gpurun_out/attn_variants.txt
  PXA_LIB_PATH=$PWD/$lib timeout 200 python tools/attn_trace.py 2>&1 | grep -E "kernel|mean|behind|[AB] w0 n=(20|21):" | cut -c1-200 >> gpurun_out/attn_variants.txt
  PXA_LIB_PATH=$PWD/$lib timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k flash_attn -p no:cacheprovider 2>&1 | tail -1 >> gpurun_out/attn_variants.txt
done
cat gpurun_out/attn_variants.txt

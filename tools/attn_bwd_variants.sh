#!/bin/bash
# A/B the attention-backward experiment builds on the GPU box (built here beforehand with build_variant; selected through
# PXA_LIB_PATH): parity tests + isolated kernel timing for each.  usage: tools/attn_bwd_variants.sh  (inside gpurun)
#   python -c "from pixart_sigma_b200 import build as b; b.build_variant('tmastats', ['PXA_BWD_TMA_STATS=1']); b.build_variant('prefetch', ['PXA_BWD_PREFETCH=1']); b.build_variant('both', ['PXA_BWD_TMA_STATS=1', 'PXA_BWD_PREFETCH=1']); b.build_variant('ew2', ['PXA_BWD_EW=2'])"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default tmastats prefetch both ew2; do
  if [ "$v" == "default" ]; then unset PXA_LIB_PATH; else export PXA_LIB_PATH=pixart_sigma_b200/build/variants/libpixart_sm100_$v.so; [ -f "$PXA_LIB_PATH" ] || continue; fi
  echo "--- $v"
  timeout 300 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k flash_attn 2>&1 | tail -1
  timeout 200 python tools/attn_bwd_bench.py
  timeout 200 python tools/attn_bwd_bench.py 8 16 1024
done 2>&1 | tee gpurun_out/attn_bwd_variants.txt

#!/bin/bash
# Round-2 GPU session 32: attention backward after the masked-column fix (default: one CTA per item; variant: persistent grid).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s32_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
timeout -k 10 400 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s32_t.log 2>&1
echo "backward tests (default) rc=$? $(tail -1 gpurun_out/r2s32_t.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s32_t.log | head -8 >> $S
PXA_LIB_PATH=$V/libpixart_sm100_bwdp.so timeout -k 10 400 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "attn" > gpurun_out/r2s32_t1.log 2>&1
echo "backward attention tests (persistent variant) rc=$? $(tail -1 gpurun_out/r2s32_t1.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s32_t1.log | head -8 >> $S
timeout -k 10 500 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s32_t2.log 2>&1
echo "training tests rc=$? $(tail -1 gpurun_out/r2s32_t2.log)" >> $S
timeout -k 10 200 python tools/attn_bwd_bench.py >> $S 2>&1
cat $S

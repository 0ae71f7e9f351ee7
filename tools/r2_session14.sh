#!/bin/bash
# Round-2 GPU session 14: persistent attention forward -- parity (bit-exact vs one CTA per item), timing, c3 step; qk_norm test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s14_summary.txt
: > $S
echo "=== attention tests" >> $S
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py tests/test_benchshape_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/r2s14_attn.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s14_attn.log)" >> $S
grep -E "^FAILED|^ERROR|assert" gpurun_out/r2s14_attn.log | head -8 | cut -c1-300 >> $S
echo "=== qk_norm training test" >> $S
timeout -k 10 300 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k "qk_norm" > gpurun_out/r2s14_qk.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s14_qk.log)" >> $S
for v in 2 4; do
  echo "--- attn_bench variant $v" >> $S
  PXA_ATTN_VARIANT=$v timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
done
for v in 2 4; do
  PXA_ATTN_VARIANT=$v timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s14_c3_v$v.json 2> gpurun_out/r2s14_c3_v$v.err
  echo "c3 attn variant=$v rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s14_c3_v$v.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity'], d['roofline']['attention'], d['clocks'])" 2>&1)" >> $S
done
cat $S

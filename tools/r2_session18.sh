#!/bin/bash
# Round-2 GPU session 18: VAE binding fix, grid-stride ln_modulate, sub-block cycle traces of the attention forward in both launch modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s18_summary.txt
: > $S
timeout -k 10 400 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "vae or groupnorm or resblock or ln_modulate or conv" > gpurun_out/r2s18_t.log 2>&1
echo "vae + ln tests rc=$? $(tail -1 gpurun_out/r2s18_t.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s18_t.log | head -5 >> $S
echo "=== ln_bench" >> $S
timeout -k 10 200 python tools/ln_bench.py >> $S 2>&1
for rep in 1 2; do
  for v in 2 4; do
    echo "=== attn_trace variant $v (run $rep)" >> $S
    timeout -k 10 200 python tools/attn_trace.py $v 2>&1 | grep -v "^  [AB] w0 n=" | cut -c1-700 >> $S
  done
done
PXA_ATTN_VARIANT=2 timeout -k 10 200 python tools/attn_bench.py 2>&1 | head -1 >> $S
PXA_ATTN_VARIANT=4 timeout -k 10 200 python tools/attn_bench.py 2>&1 | head -1 >> $S
cat $S

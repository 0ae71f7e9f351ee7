#!/bin/bash
# Round-2 GPU session 10: fused Mlp kernel (parity under dependency stress, then timing and the c3 step), direct-grad test debug.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s10_summary.txt
: > $S
echo "=== fused mlp tests" >> $S
timeout -k 10 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "mlp_fused" -x > gpurun_out/r2s10_mlp.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s10_mlp.log)" >> $S
grep -E "^FAILED|^ERROR|rel_err|assert" gpurun_out/r2s10_mlp.log | head -8 >> $S
echo "=== direct grad test" >> $S
timeout -k 10 240 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k "direct_gradient" > gpurun_out/r2s10_dg.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s10_dg.log)" >> $S
grep -E "AssertionError|assert " gpurun_out/r2s10_dg.log | head -4 | cut -c1-600 >> $S
echo "=== timing: two launches vs one persistent kernel (M = 32768)" >> $S
timeout -k 10 200 python tools/mlp_fused_bench.py >> $S 2>&1
for f in 0 1; do
  PXA_MLP_FUSED=$f timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s10_c3_mlp$f.json 2> gpurun_out/r2s10_c3_mlp$f.err
  echo "c3 mlp_fused=$f rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s10_c3_mlp$f.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['block_rel_err'], d['clocks'])" 2>&1)" >> $S
done
cat $S

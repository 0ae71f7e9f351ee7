#!/bin/bash
# Round-2 GPU session 11: fused Mlp with fc2 lagging several groups behind fc1; direct-grad test after the idempotent readiness fix.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s11_summary.txt
: > $S
echo "=== fused mlp tests" >> $S
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "mlp_fused" > gpurun_out/r2s11_mlp.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s11_mlp.log)" >> $S
grep -E "^FAILED|^ERROR|rel_err|assert" gpurun_out/r2s11_mlp.log | head -8 >> $S
echo "=== direct grad test + training suite" >> $S
timeout -k 10 500 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s11_dg.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s11_dg.log)" >> $S
grep -E "AssertionError|assert |^FAILED" gpurun_out/r2s11_dg.log | head -6 | cut -c1-600 >> $S
echo "=== timing: two launches vs one persistent kernel (M = 32768)" >> $S
timeout -k 10 300 python tools/mlp_fused_bench.py >> $S 2>&1
cat $S

#!/bin/bash
# Backward-kernel parity tests in one gpurun call (each group under its own timeout so a hang cannot hide the others).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "not flash_attn" > gpurun_out/bwd_elem.log 2>&1
echo "elementwise rc=$? $(tail -1 gpurun_out/bwd_elem.log)" > gpurun_out/bwd_summary.txt
timeout 240 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/bwd_attn.log 2>&1
echo "attention rc=$? $(tail -1 gpurun_out/bwd_attn.log)" >> gpurun_out/bwd_summary.txt
if [ -n "$1" ]; then
  timeout 600 python -m pytest $1 -q -m gpu -p no:cacheprovider -x > gpurun_out/bwd_extra.log 2>&1
  echo "extra rc=$? $(tail -1 gpurun_out/bwd_extra.log)" >> gpurun_out/bwd_summary.txt
fi
cat gpurun_out/bwd_summary.txt
grep -E "^(FAILED|ERROR)|Error|rel_err|assert " gpurun_out/bwd_elem.log gpurun_out/bwd_attn.log | head -60

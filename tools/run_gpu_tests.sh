#!/bin/bash
# Run the GPU kernel tests group by group, each under its own timeout so one hung kernel cannot eat the whole
# gpurun call.  Logs land in gpurun_out/ (merged back to the build container).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/tests_summary.txt
  timeout "$to" python -m pytest "$@" -q -m gpu -p no:cacheprovider > "gpurun_out/test_$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -1 gpurun_out/test_$name.log)" | tee -a gpurun_out/tests_summary.txt
}
: > gpurun_out/tests_summary.txt
run gemm_small 240 tests/test_kernels_gpu.py -k "gemm_bias_bf16 and (128-192-64 or 384-256-128)"
run gemm_all 300 tests/test_kernels_gpu.py -k "gemm"
run ln 300 tests/test_kernels_gpu.py -k "ln_modulate or kv_compress"
run attn_small 240 tests/test_kernels_gpu.py -k "flash_attn_self and (1-1-128-128 or 1-2-256-256)"
run attn_all 400 tests/test_kernels_gpu.py -k "flash_attn"
for extra in "$@"; do run "$(basename "$extra" .py)" 600 "$extra"; done
cat gpurun_out/tests_summary.txt

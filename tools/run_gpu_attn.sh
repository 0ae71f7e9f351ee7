#!/bin/bash
# Attention-kernel iteration loop: kernel tests, cycle trace, experiment builds, model tests, one bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-a}
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/test_attn.log 2>&1; echo "attn tests rc=$? $(tail -1 gpurun_out/test_attn.log)" > gpurun_out/attn_summary.txt
timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "trace rc=$?" >> gpurun_out/attn_summary.txt
if ls pixart_sigma_b200/build/variants/*.so > /dev/null 2>&1; then bash tools/attn_variants.sh > /dev/null 2>&1; fi
timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "mask or block" > gpurun_out/test_model_q.log 2>&1; echo "model rc=$? $(tail -1 gpurun_out/test_model_q.log)" >> gpurun_out/attn_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?" >> gpurun_out/attn_summary.txt
cut -c1-400 gpurun_out/bench_${TAG}.json >> gpurun_out/attn_summary.txt
cat gpurun_out/attn_summary.txt

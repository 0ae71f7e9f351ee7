#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-q}
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "not cta_pair" > gpurun_out/test_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 gpurun_out/test_kernels.log)" > gpurun_out/quick_summary.txt
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "cta_pair and 256-192-64" > gpurun_out/test_pair0.log 2>&1; echo "pair0 rc=$? $(tail -1 gpurun_out/test_pair0.log)" >> gpurun_out/quick_summary.txt
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "cta_pair" > gpurun_out/test_pair.log 2>&1; echo "pair rc=$? $(tail -1 gpurun_out/test_pair.log)" >> gpurun_out/quick_summary.txt
timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "mask or block" > gpurun_out/test_model_q.log 2>&1; echo "model rc=$? $(tail -1 gpurun_out/test_model_q.log)" >> gpurun_out/quick_summary.txt
timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; echo "trace rc=$?" >> gpurun_out/quick_summary.txt
tools/micro/mufu_warps > gpurun_out/mufu_warps.txt 2>&1
timeout 120 python tools/gemm_trace.py 1152 > gpurun_out/gemm_trace.txt 2>&1; echo "gemm trace rc=$?" >> gpurun_out/quick_summary.txt
if [ "$2" == "gemm" ]; then timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; echo "gemm_bench rc=$?" >> gpurun_out/quick_summary.txt; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?" >> gpurun_out/quick_summary.txt
cat gpurun_out/bench_${TAG}.json >> gpurun_out/quick_summary.txt
cat gpurun_out/quick_summary.txt

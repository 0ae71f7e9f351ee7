#!/bin/bash
# Other BASELINE configs through bench.py (c2 512px, c4 2K kv-compress) + gemm epilogue trace. Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?" > gpurun_out/misc_summary.txt
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?" >> gpurun_out/misc_summary.txt
timeout 120 python tools/gemm_trace.py 1152 > gpurun_out/gemm_trace.txt 2>&1; echo "gemm trace rc=$?" >> gpurun_out/misc_summary.txt
cat gpurun_out/misc_summary.txt; cat gpurun_out/bench_c2.json gpurun_out/bench_c4.json | cut -c1-400

#!/bin/bash
# Model-level parity tests, smoke, bench and ncu captures in one gpurun call. Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
echo "=== model tests" > gpurun_out/round_summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_vae_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/test_model.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/test_model.log)" >> gpurun_out/round_summary.txt
echo "=== smoke" >> gpurun_out/round_summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/smoke.log)" >> gpurun_out/round_summary.txt
echo "=== bench c3" >> gpurun_out/round_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "rc=$?" >> gpurun_out/round_summary.txt
cat gpurun_out/bench_${TAG}.json >> gpurun_out/round_summary.txt
if [ "$2" != "noncu" ]; then
  echo "=== ncu launch list" >> gpurun_out/round_summary.txt
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv \
      --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  echo "rc=$?" >> gpurun_out/round_summary.txt
  echo "=== ncu full gemm + attn" >> gpurun_out/round_summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm|flash_attn" -s 340 -c 14 \
      -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "rc=$?" >> gpurun_out/round_summary.txt
fi
cat gpurun_out/round_summary.txt

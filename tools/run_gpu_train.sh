#!/bin/bash
# Training-path measurements in one gpurun call: attention fwd/bwd microbench, c5 bench line, ncu launch list of a training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1t}
timeout 200 python tools/attn_bwd_bench.py > gpurun_out/attn_bwd_bench.txt 2>&1; echo "attn bench rc=$?"; cat gpurun_out/attn_bwd_bench.txt
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err; echo "bench c5 rc=$?"
cat gpurun_out/bench_c5_${TAG}.json; tail -5 gpurun_out/bench_c5_${TAG}.err
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv \
      --log-file gpurun_out/launches_c5_${TAG}.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5.log 2>&1
  echo "ncu rc=$?"
fi

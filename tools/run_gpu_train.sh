#!/bin/bash
# Training-path measurements in one gpurun call: attention fwd/bwd microbench, c5 bench line, ncu launch list of a training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1t}
timeout 200 python tools/attn_bwd_bench.py > gpurun_out/attn_bwd_bench.txt 2>&1; echo "attn bench rc=$?"; cat gpurun_out/attn_bwd_bench.txt
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err; echo "bench c5 rc=$?"
cat gpurun_out/bench_c5_${TAG}.json; tail -5 gpurun_out/bench_c5_${TAG}.err
if [ "$2" == "ncu-bench" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv \
      --log-file gpurun_out/launches_c5_${TAG}.csv python bench.py --workload c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5.log 2>&1
  echo "ncu rc=$?"
fi
if [ "$2" == "prof" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_c5_${TAG}.csv python tools/train_profile.py --depth 4 > gpurun_out/ncu_c5_launch.log 2>&1
  echo "ncu launch list rc=$?"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:"flash_attn_d72_bwd|ln_modulate_bwd|gate_residual|transpose|gelu|colsum|attn_delta" -c 40 \
      -o gpurun_out/prof_c5_${TAG} -f python tools/train_profile.py --depth 1 > gpurun_out/ncu_c5_full.log 2>&1
  echo "ncu full rc=$?"
fi

#!/bin/bash
# Round-end evidence in one gpurun call: the whole -m gpu suite, smoke, the default bench line (c3) + its ncu captures, the
# training bench (c5, with and without activation checkpointing) + ncu captures of a training step.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1f}
S=gpurun_out/final_summary.txt
echo "=== pytest -m gpu" > $S
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/test_all.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/test_all.log)" >> $S
echo "=== smoke" >> $S
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/smoke.log)" >> $S
echo "=== bench c3" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "rc=$?" >> $S; cat gpurun_out/bench_${TAG}.json >> $S
echo "=== bench reference arm" >> $S
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "rc=$?" >> $S; cat gpurun_out/bench_${TAG}_ref.json >> $S
echo "=== bench c5 (checkpointing / no checkpointing)" >> $S
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err
echo "rc=$?" >> $S; cat gpurun_out/bench_c5_${TAG}.json >> $S
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-checkpoint > gpurun_out/bench_c5_${TAG}_nockpt.json 2> gpurun_out/bench_c5_${TAG}_nockpt.err
echo "rc=$?" >> $S; cat gpurun_out/bench_c5_${TAG}_nockpt.json >> $S
if [ "$2" != "noncu" ]; then
  echo "=== ncu c3 launch list + full" >> $S
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv \
      --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  echo "rc=$?" >> $S
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm|flash_attn" -s 340 -c 14 \
      -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "rc=$?" >> $S
  echo "=== ncu training step (depth 4 launch list, depth 1 full on the backward kernels)" >> $S
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_c5_${TAG}.csv python tools/train_profile.py --depth 4 > gpurun_out/ncu_c5_launch.log 2>&1
  echo "rc=$?" >> $S
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:"flash_attn_d72_bwd|gemm_bf16_kernel|ln_modulate_bwd|gate_residual|gelu|colsum|attn_delta" -c 60 \
      -o gpurun_out/prof_c5_${TAG} -f python tools/train_profile.py --depth 1 > gpurun_out/ncu_c5_full.log 2>&1
  echo "rc=$?" >> $S
fi
cat $S

#!/bin/bash
# Round-2 end-of-round evidence in one gpurun call (tag r2f): the whole -m gpu suite, smoke, the default bench line and the
# reference arm, the c5 training lines, ncu launch lists (c3 step, training step) and one --set full capture of a block's kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2f}
S=gpurun_out/final_summary.txt
echo "=== pytest -m gpu" > $S
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_all.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/test_all.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/test_all.log | head -10 >> $S
echo "=== smoke" >> $S
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/smoke.log)" >> $S
echo "=== bench (default line: c3 + train + c4 extras)" >> $S
timeout -k 10 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "rc=$?" >> $S; cat gpurun_out/bench_${TAG}.json >> $S
echo "=== bench reference arm" >> $S
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "rc=$?" >> $S; cat gpurun_out/bench_${TAG}_ref.json >> $S
echo "=== bench c5 (one CUDA graph; with / without activation checkpointing)" >> $S
timeout -k 10 600 python bench.py --workload c5 --steps 8 --warmup 3 > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err
echo "rc=$?" >> $S; cat gpurun_out/bench_c5_${TAG}.json >> $S
timeout -k 10 600 python bench.py --workload c5 --steps 8 --warmup 3 --no-checkpoint > gpurun_out/bench_c5_${TAG}_nockpt.json 2> gpurun_out/bench_c5_${TAG}_nockpt.err
echo "rc=$?" >> $S; cat gpurun_out/bench_c5_${TAG}_nockpt.json >> $S
echo "=== bench vae" >> $S
timeout -k 10 400 python bench.py --workload vae > gpurun_out/bench_vae_${TAG}.json 2> gpurun_out/bench_vae_${TAG}.err
echo "rc=$?" >> $S; cat gpurun_out/bench_vae_${TAG}.json >> $S
if [ "$2" != "noncu" ]; then
  echo "=== ncu c3 launch list + full" >> $S
  PXA_PROFILER_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
      --log-file gpurun_out/launches_${TAG}.csv python bench.py --no-cuda-graph --no-extras --no-parity --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
  echo "rc=$?" >> $S
  # one block's kernels (11 launches) from the third block of the first warm-up forward
  timeout -k 10 900 ncu --set full --clock-control none -k regex:"gemm|flash_attn|ln_modulate" -s 22 -c 11 \
      -o gpurun_out/prof_${TAG} -f python bench.py --no-cuda-graph --no-extras --no-parity --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  echo "rc=$?" >> $S
  ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
  ls -la gpurun_out/prof_${TAG}.ncu-rep >> $S
  # gpurun merges at most 64 MiB back: keep the report only if it is small, the raw CSV always
  [ $(stat -c %s gpurun_out/prof_${TAG}.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ] && rm -f gpurun_out/prof_${TAG}.ncu-rep
  echo "=== ncu training step (depth 4 launch list)" >> $S
  timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_c5_${TAG}.csv python tools/train_profile.py --depth 4 > gpurun_out/ncu_c5_launch.log 2>&1
  echo "rc=$?" >> $S
fi
cat $S

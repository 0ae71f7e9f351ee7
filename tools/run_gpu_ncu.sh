#!/bin/bash
# ncu captures only (the bench / test lines come from run_gpu_final.sh).  The .ncu-rep files are converted to raw CSV on
# the box and deleted: gpurun merges at most 64 MiB back (a --set full report is ~2 MB per kernel).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "c3 launch list rc=$?"
if [ "$2" == "c3full" ]; then
  timeout 600 ncu --set full --clock-control none -k regex:"gemm|flash_attn" -s 340 -c 14 \
      -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "c3 full rc=$?"
  ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null; rm -f gpurun_out/prof_${TAG}.ncu-rep
fi
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_c5_${TAG}.csv python tools/train_profile.py --depth 4 > gpurun_out/ncu_c5_launch.log 2>&1
echo "c5 launch list rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"flash_attn_d72_bwd|gemm_bf16_kernel|ln_modulate_bwd|gate_residual|gelu|colsum" -c 24 \
    -o gpurun_out/prof_c5_${TAG} -f python tools/train_profile.py --depth 1 > gpurun_out/ncu_c5_full.log 2>&1
echo "c5 full rc=$?"
ncu -i gpurun_out/prof_c5_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_c5_${TAG}_raw.csv 2>/dev/null; rm -f gpurun_out/prof_c5_${TAG}.ncu-rep
ls -la gpurun_out/; du -sh gpurun_out

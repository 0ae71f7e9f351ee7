#!/bin/bash
# Round-2 GPU session 8: training-step launch list (where the non-GEMM / non-attention 30 % goes), DRAM bytes per kernel with
# and without L2 chaining (no cache flush between kernels), c3 bench with the reversed second LN pass.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s8_summary.txt
: > $S
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k "ln_modulate or adamw" 2>&1 | tail -1 >> $S
for c in 1 0; do
  PXA_L2_CHAIN=$c timeout 600 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2s8_c3_chain$c.json 2> gpurun_out/r2s8_c3_chain$c.err
  echo "c3 l2_chain=$c rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s8_c3_chain$c.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])" 2>&1)" >> $S
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_c5_r2a.csv python tools/train_profile.py --depth 4 > gpurun_out/ncu_c5_launch.log 2>&1
echo "c5 launch list rc=$?" >> $S
for c in 1 0; do
  PXA_L2_CHAIN=$c timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
      -k regex:"gemm|flash_attn|ln_modulate" -s 700 -c 44 --csv --log-file gpurun_out/r2s8_dram_chain$c.csv \
      python bench.py --no-extras --no-cpu-baseline --no-parity --no-cuda-graph --steps 1 --warmup 3 > gpurun_out/r2s8_ncu_chain$c.log 2>&1
  echo "ncu chain=$c rc=$?" >> $S
done
cat $S

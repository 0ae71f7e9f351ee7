#!/bin/bash
# Round-2 GPU session 9: block-level autograd node + direct gradient accumulation: training parity, c5 bench (old vs new).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s9_summary.txt
: > $S
echo "=== training tests (BlockFn)" >> $S
timeout 1200 python -m pytest tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_benchshape_gpu.py -q -m gpu -p no:cacheprovider -k "not flash_attn_at and not block_at" > gpurun_out/r2s9_train.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s9_train.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s9_train.log | head >> $S
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2s9_smoke.log 2>&1
echo "smoke rc=$? $(tail -2 gpurun_out/r2s9_smoke.log | tr '\n' ' ')" >> $S
for cfg in "1 1" "0 0"; do
  set -- $cfg
  for ck in "" "--no-checkpoint"; do
    PXA_BLOCK_FN=$1 PXA_DIRECT_GRAD=$2 timeout 600 python bench.py --workload c5 --no-cpu-baseline $ck > gpurun_out/r2s9_c5_$1$2$ck.json 2> gpurun_out/r2s9_c5_$1$2$ck.err
    echo "c5 block_fn=$1 direct_grad=$2 $ck rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s9_c5_$1$2$ck.json'));print(d['ms_per_step'], d['value'], d['roofline']['whole_step'], d['config']['peak_mem_gib'], d['config']['loss'])" 2>&1)" >> $S
    tail -2 gpurun_out/r2s9_c5_$1$2$ck.err | cut -c1-300 >> $S
  done
done
cat $S

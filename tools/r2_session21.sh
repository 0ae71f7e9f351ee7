#!/bin/bash
# Round-2 GPU session 21: more in-step A/B under the power cap (attention build switches with FADD2 row sums as the new default).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s21_summary.txt
: > $S
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s21_$tag.json 2> gpurun_out/r2s21_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s21_$tag.json'));r=d['roofline'];print(round(d['ms_per_step'],2), round(d['value'],2), 'gemm', round(r['achieved']), 'attn', round(r['attention']['achieved']), 'parity', round(d['parity']['block_rel_err'],6), d['clocks'])" 2>&1)" >> $S
}
V=$PWD/pixart_sigma_b200/build/variants
run default0 PXA_DUMMY=0
run summma PXA_LIB_PATH=$V/libpixart_sm100_summma.so
run pre0 PXA_LIB_PATH=$V/libpixart_sm100_pre0.so
run attn4 PXA_ATTN_VARIANT=4
run nochain PXA_L2_CHAIN=0
run default1 PXA_DUMMY=0
cat $S
# c3 launch list of the timed step only (the first capture of r2f caught the weight initialisation instead)
PXA_PROFILER_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2f.csv python bench.py --no-cuda-graph --no-extras --no-parity --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$? $(grep -c flash_attn gpurun_out/launches_r2f.csv) attention rows" >> $S
cat $S

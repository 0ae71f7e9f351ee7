#!/bin/bash
# Round-2 GPU session 26: attention output rows through smem + one TMA bulk copy per row (2 K/V stages) vs direct stores (3 stages).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s26_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py tests/test_benchshape_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn or block or forward or model" > gpurun_out/r2s26_t.log 2>&1
echo "attention + model tests rc=$? $(tail -1 gpurun_out/r2s26_t.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s26_t.log | head -8 >> $S
for rep in 1 2; do
echo "--- bulk copy epilogue (default build)" >> $S
timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
echo "--- direct stores" >> $S
PXA_LIB_PATH=$V/libpixart_sm100_direct.so timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
done
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s26_$tag.json 2> gpurun_out/r2s26_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s26_$tag.json'));r=d['roofline'];print(round(d['ms_per_step'],2), round(d['value'],2), 'gemm', round(r['achieved']), 'attn', round(r['attention']['achieved']), 'parity', round(d['parity']['block_rel_err'],6), d['clocks']['sm_mhz'])" 2>&1)" >> $S
}
run bulk0 PXA_DUMMY=0
run direct PXA_LIB_PATH=$V/libpixart_sm100_direct.so
run bulk1 PXA_DUMMY=0
timeout -k 10 300 python -m pytest tests/test_training_gpu.py tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s26_t2.log 2>&1
echo "training + backward tests rc=$? $(tail -1 gpurun_out/r2s26_t2.log)" >> $S
cat $S

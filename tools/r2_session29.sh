#!/bin/bash
# Round-2 GPU session 29: tile tensor store + persistent grid for EVERY key count (the store then drains under the next item) vs
# direct stores with the current auto policy (persistent for Nk <= 1024 only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s29_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
for rep in 1 2; do
echo "--- tile tensor store, persistent for every Nk (PXA_ATTN_VARIANT=4)" >> $S
PXA_ATTN_VARIANT=4 timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
echo "--- direct stores, auto policy" >> $S
PXA_LIB_PATH=$V/libpixart_sm100_direct.so timeout -k 10 300 python tools/attn_bench.py >> $S 2>&1
echo "--- direct stores, persistent for every Nk" >> $S
PXA_ATTN_VARIANT=4 PXA_LIB_PATH=$V/libpixart_sm100_direct.so timeout -k 10 300 python tools/attn_bench.py 2>&1 | head -2 >> $S
done
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s29_$tag.json 2> gpurun_out/r2s29_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s29_$tag.json'));r=d['roofline'];print(round(d['ms_per_step'],2), round(d['value'],2), 'gemm', round(r['achieved']), 'attn', round(r['attention']['achieved']), 'parity', round(d['parity']['block_rel_err'],6), d['clocks']['sm_mhz'])" 2>&1)" >> $S
}
run tile_p4 PXA_ATTN_VARIANT=4
run direct_auto PXA_LIB_PATH=$V/libpixart_sm100_direct.so
run direct_p4 PXA_ATTN_VARIANT=4 PXA_LIB_PATH=$V/libpixart_sm100_direct.so
run tile_p4b PXA_ATTN_VARIANT=4
run direct_autob PXA_LIB_PATH=$V/libpixart_sm100_direct.so
cat $S

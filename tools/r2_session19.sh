#!/bin/bash
# Round-2 GPU session 19: the c3 step is power-capped (1 kW, SM clock ~1.65 GHz) -- A/B of choices that trade time in isolation
# against energy: fused Mlp (half the DRAM bytes), persistent attention at 4096 keys, row sums on FADD2 instead of the tensor pipe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s19_summary.txt
: > $S
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s19_$tag.json 2> gpurun_out/r2s19_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s19_$tag.json'));r=d['roofline'];print(round(d['ms_per_step'],2), round(d['value'],2), 'gemm', round(r['achieved']), 'attn', round(r['attention']['achieved']), 'parity', round(d['parity']['block_rel_err'],6), d['clocks'])" 2>&1)" >> $S
}
run default0 PXA_DUMMY=0
run mlpfused PXA_MLP_FUSED=1
run attn4 PXA_ATTN_VARIANT=4
run nosummma PXA_LIB_PATH=$PWD/pixart_sigma_b200/build/variants/libpixart_sm100_nosummma.so
run fuseln PXA_FUSE_LN=1
run default1 PXA_DUMMY=0
cat $S

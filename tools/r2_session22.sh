#!/bin/bash
# Round-2 GPU session 22: batched kv_linear projection (one GEMM per forward) -- parity + c3 A/B; model / sampler suites.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s22_summary.txt
: > $S
timeout -k 10 600 python -m pytest tests/test_model_gpu.py tests/test_sampler_gpu.py tests/test_benchshape_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s22_t.log 2>&1
echo "model + sampler + benchshape tests rc=$? $(tail -1 gpurun_out/r2s22_t.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s22_t.log | head -8 >> $S
timeout -k 10 200 python __graft_entry__.py --smoke > gpurun_out/r2s22_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r2s22_smoke.log)" >> $S
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s22_$tag.json 2> gpurun_out/r2s22_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s22_$tag.json'));r=d['roofline'];print(round(d['ms_per_step'],2), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], 'gemm', round(r['achieved']), 'attn', round(r['attention']['achieved']), 'parity', round(d['parity']['block_rel_err'],6), round(d['parity'].get('forward_rel_err') or 0,5), d['clocks']['sm_mhz'])" 2>&1)" >> $S
  tail -2 gpurun_out/r2s22_$tag.err >> $S
}
run kvbatch1 PXA_KV_BATCH=1
run kvbatch0 PXA_KV_BATCH=0
run kvbatch1b PXA_KV_BATCH=1
cat $S

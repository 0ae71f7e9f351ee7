#!/bin/bash
# Round-2 GPU session 6: three-tile attention variant (parity + timing vs the two-tile kernel), drop_path / remaining tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s6_summary.txt
: > $S
echo "=== attention tests (variants 2 and 3)" >> $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_benchshape_gpu.py -q -m gpu -p no:cacheprovider -k "flash_attn" > gpurun_out/r2s6_attn.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s6_attn.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s6_attn.log | head -20 >> $S
for v in 2 3; do
  echo "--- variant $v" >> $S
  PXA_ATTN_VARIANT=$v timeout 200 python tools/attn_bwd_bench.py 8 16 4096 2>&1 | tail -1 >> $S
  PXA_ATTN_VARIANT=$v timeout 200 python tools/attn_bwd_bench.py 2 16 16384 2>&1 | tail -1 >> $S
  PXA_ATTN_VARIANT=$v timeout 200 python tools/attn_bwd_bench.py 8 16 1024 2>&1 | tail -1 >> $S
done
echo "=== rest of the suite" >> $S
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "not flash_attn" > gpurun_out/r2s6_suite.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s6_suite.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s6_suite.log | head -20 >> $S
for v in 2 3; do
  PXA_ATTN_VARIANT=$v timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s6_c3_v$v.json 2> gpurun_out/r2s6_c3_v$v.err
  echo "c3 attn variant=$v rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s6_c3_v$v.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['block_rel_err'], d['roofline']['frac'], d['roofline']['attention']['frac'], d['clocks'])" 2>&1)" >> $S
done
cat $S

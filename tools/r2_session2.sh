#!/bin/bash
# Round-2 GPU session 2: full GPU suite with the fused LayerNorm chain, GEMM epilogue variants, c3 bench fused vs unfused.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s2_summary.txt
: > $S
echo "=== full gpu suite" >> $S
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2s2_suite.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s2_suite.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s2_suite.log >> $S
echo "=== gemm variants" >> $S
timeout 600 python tools/gemm_bench2.py >> $S 2>&1
for f in 1 0; do
  PXA_FUSE_LN=$f timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s2_c3_fuse$f.json 2> gpurun_out/r2s2_c3_fuse$f.err
  echo "c3 fuse_ln=$f rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s2_c3_fuse$f.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])" 2>&1)" >> $S
done
echo "=== exp2 MUFU / FMA-pipe mix (tools/micro/exp_mix.cu)" >> $S
timeout 120 tools/micro/exp_mix >> $S 2>&1
cat $S

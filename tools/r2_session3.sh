#!/bin/bash
# Round-2 GPU session 3: suite after the ln_prepare fix + epilogue prefetch, GEMM variants, exp2 mix microbench, new bench.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s3_summary.txt
: > $S
echo "=== full gpu suite" >> $S
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2s3_suite.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s3_suite.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s3_suite.log >> $S
echo "=== exp2 MUFU / FMA-pipe mix (tools/micro/exp_mix.cu)" >> $S
timeout 120 tools/micro/exp_mix >> $S 2>&1
echo "=== gemm variants" >> $S
timeout 600 python tools/gemm_bench2.py >> $S 2>&1
for f in 1 0; do
  PXA_FUSE_LN=$f timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s3_c3_fuse$f.json 2> gpurun_out/r2s3_c3_fuse$f.err
  echo "c3 fuse_ln=$f rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s3_c3_fuse$f.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity'], d['clocks'])" 2>&1)" >> $S
done
timeout 900 python bench.py > gpurun_out/r2s3_bench_full.json 2> gpurun_out/r2s3_bench_full.err
echo "bench default rc=$?" >> $S
python -c "
import json; d=json.load(open('gpurun_out/r2s3_bench_full.json'))
print('c3', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity'], 'cpu', d['cpu_baseline'])
print('train', d['train']['ms_per_step'], d['train']['value'], d['train']['config']['step_mode'])
print('c4', d['c4']['ms_per_step'], d['c4']['value'])
" >> $S 2>&1
tail -5 gpurun_out/r2s3_bench_full.err >> $S
cat $S

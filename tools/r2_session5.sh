#!/bin/bash
# Round-2 GPU session 5: 8-warp pair epilogue, GroupNorm kernel, VAE bench, fused-LN decision, attention trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s5_summary.txt
: > $S
echo "=== full gpu suite" >> $S
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2s5_suite.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s5_suite.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s5_suite.log >> $S
echo "=== gemm variants" >> $S
timeout 600 python tools/gemm_bench2.py 2>&1 | grep -E "==|epilogue|rmw \+ aux\*s|reduce" >> $S
for f in 1 0; do
  PXA_FUSE_LN=$f timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s5_c3_fuse$f.json 2> gpurun_out/r2s5_c3_fuse$f.err
  echo "c3 fuse_ln=$f rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s5_c3_fuse$f.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['block_rel_err'], d['roofline']['frac'], d['roofline']['attention']['frac'], d['clocks'])" 2>&1)" >> $S
done
timeout 600 python bench.py --workload vae > gpurun_out/r2s5_vae.json 2> gpurun_out/r2s5_vae.err
echo "vae rc=$?: $(cat gpurun_out/r2s5_vae.json | cut -c1-1500)" >> $S
tail -3 gpurun_out/r2s5_vae.err >> $S
echo "=== attention trace" >> $S
timeout 200 python tools/attn_trace.py 2>&1 | cut -c1-260 >> $S
cat $S

#!/bin/bash
# Session 33 (round 2, re-entry): GPU validation of the fp32_attention mode (P as bf16 hi + lo), the whole SDXL-VAE assembly and the
# T5 encoder; full -m gpu suite, smoke, attention micro-benchmark, default bench line, t5 / vae workloads.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/s33_summary.txt
echo "=== pytest -m gpu (whole suite, no -x)" > $S
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/s33_tests.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/s33_tests.log)" >> $S
grep -E "^FAILED|^ERROR|fp32_p |fp32_attention|SDXL-VAE|^T5 " gpurun_out/s33_tests.log >> $S
grep -E "fp32_attention" gpurun_out/parity.txt >> $S
echo "=== smoke" >> $S
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s33_smoke.log 2>&1
echo "rc=$? $(tail -2 gpurun_out/s33_smoke.log)" >> $S
echo "=== attention micro-benchmark (bf16 P vs hi+lo P)" >> $S
timeout 200 python tools/attn_bench.py > gpurun_out/s33_attn_bench.txt 2>&1
echo "rc=$?" >> $S; cat gpurun_out/s33_attn_bench.txt >> $S
echo "=== bench default (c3 + train + c4)" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s33_bench.json 2> gpurun_out/s33_bench.err
echo "rc=$?" >> $S; cat gpurun_out/s33_bench.json >> $S; tail -3 gpurun_out/s33_bench.err >> $S
echo "=== bench t5" >> $S
timeout 300 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s33_bench_t5.json 2> gpurun_out/s33_bench_t5.err
echo "rc=$?" >> $S; cat gpurun_out/s33_bench_t5.json >> $S; tail -3 gpurun_out/s33_bench_t5.err >> $S
echo "=== bench vae" >> $S
timeout 300 python bench.py --workload vae --steps 5 --warmup 3 > gpurun_out/s33_bench_vae.json 2> gpurun_out/s33_bench_vae.err
echo "rc=$?" >> $S; cat gpurun_out/s33_bench_vae.json >> $S; tail -3 gpurun_out/s33_bench_vae.err >> $S
echo "=== bench c5, fp32_attention off (A/B of the hi+lo P forward inside the training step)" >> $S
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-fp32-attention --no-cpu-baseline > gpurun_out/s33_bench_c5_nofp32.json 2> gpurun_out/s33_bench_c5_nofp32.err
echo "rc=$?" >> $S; python - <<'P' >> $S
import json
try:
    d = json.loads(open('gpurun_out/s33_bench_c5_nofp32.json').read().strip().splitlines()[-1])
    print('c5 fp32_attention off:', d['ms_per_step'], d['value'], d['config'].get('fp32_attention'))
    d = json.loads(open('gpurun_out/s33_bench.json').read().strip().splitlines()[-1])
    print('c5 (train sub-line of the default bench):', d['train']['ms_per_step'], d['train']['value'], d['train']['config'].get('fp32_attention'))
except Exception as e:
    print('parse failed', e)
P
tail -c 3000 $S

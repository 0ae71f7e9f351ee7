#!/bin/bash
# Round-2 GPU session 7: L2 chaining (reverse tile / batch order) on and off, GroupNorm rewrite, new tests, ncu DRAM bytes of fc2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s7_summary.txt
: > $S
echo "=== new / changed tests" >> $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_vae_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s7_tests.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s7_tests.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s7_tests.log | head >> $S
for c in 1 0; do
  PXA_L2_CHAIN=$c timeout 600 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2s7_c3_chain$c.json 2> gpurun_out/r2s7_c3_chain$c.err
  echo "c3 l2_chain=$c rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s7_c3_chain$c.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['attention']['frac'], d['clocks'])" 2>&1)" >> $S
done
timeout 600 python bench.py --workload vae > gpurun_out/r2s7_vae.json 2> gpurun_out/r2s7_vae.err
echo "vae rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s7_vae.json'));print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['groupnorm_silu'])" 2>&1)" >> $S
echo "=== ncu: dram bytes of the block's kernels with and without L2 chaining (1 step)" >> $S
for c in 1 0; do
  PXA_L2_CHAIN=$c timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:"gemm|flash_attn|ln_modulate" -s 700 -c 44 --csv --log-file gpurun_out/r2s7_dram_chain$c.csv \
      python bench.py --no-extras --no-cpu-baseline --no-parity --no-cuda-graph --steps 1 --warmup 3 > gpurun_out/r2s7_ncu_chain$c.log 2>&1
  echo "ncu chain=$c rc=$?" >> $S
done
cat $S

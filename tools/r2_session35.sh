#!/bin/bash
# Session 35 (last of round 2): gate the two new code paths under short timeouts (convolution over a virtual width, T5 attention
# with the Toeplitz bias in smem), then the whole -m gpu suite, the t5 / vae workloads and the default bench line of the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/s35_summary.txt
echo "=== gate 1: conv3x3 incl. non-tiling widths" > $S
timeout 150 python -m pytest tests/test_vae_gpu.py -q -m gpu -p no:cacheprovider -k "conv3x3" > gpurun_out/s35_conv.log 2>&1
G1=$?; echo "rc=$G1 $(tail -1 gpurun_out/s35_conv.log)" >> $S
grep -E "^FAILED|^ERROR|Error" gpurun_out/s35_conv.log | head -20 >> $S
echo "=== gate 2: T5 attention kernel (dense + Toeplitz)" >> $S
timeout 150 python -m pytest tests/test_t5_gpu.py -q -m gpu -p no:cacheprovider -k "attention_kernel" > gpurun_out/s35_t5k.log 2>&1
G2=$?; echo "rc=$G2 $(tail -1 gpurun_out/s35_t5k.log)" >> $S
grep -E "^FAILED|^ERROR|Error" gpurun_out/s35_t5k.log | head -20 >> $S
if [ $G1 -ne 0 ] || [ $G2 -ne 0 ]; then echo "gate failed: stopping" >> $S; tail -c 2000 $S; exit 0; fi
echo "=== pytest -m gpu (whole suite)" >> $S
timeout 500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/s35_tests.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/s35_tests.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/s35_tests.log | head -20 >> $S
grep -o "SDXL-VAE[^;]*rel_err [0-9.e-]*[^.]*" gpurun_out/s35_tests.log | head -5 >> $S
grep -o "T5 [a-z0-9-]* attn=[a-z]*: last_hidden_state[^,]*, [0-9]* kernel launches" gpurun_out/s35_tests.log >> $S
echo "=== new kernels alone + ncu --set full" >> $S
timeout 120 python tools/new_kernels_bench.py > gpurun_out/s35_new_kernels.txt 2>&1
echo "rc=$?" >> $S; cat gpurun_out/s35_new_kernels.txt >> $S
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"flash_attn_d72|t5_attn" -c 2 -o gpurun_out/prof_r2g_new -f python tools/new_kernels_bench.py --ncu > gpurun_out/s35_ncu.log 2>&1
echo "ncu rc=$?" >> $S
echo "=== bench t5" >> $S
timeout 240 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s35_bench_t5.json 2> gpurun_out/s35_bench_t5.err
echo "rc=$?" >> $S; cat gpurun_out/s35_bench_t5.json >> $S; tail -2 gpurun_out/s35_bench_t5.err >> $S
echo "=== bench vae" >> $S
timeout 240 python bench.py --workload vae --steps 5 --warmup 3 > gpurun_out/s35_bench_vae.json 2> gpurun_out/s35_bench_vae.err
echo "rc=$?" >> $S; cat gpurun_out/s35_bench_vae.json >> $S; tail -2 gpurun_out/s35_bench_vae.err >> $S
echo "=== bench default (c3 + train + c4)" >> $S
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err
echo "rc=$?" >> $S; cat gpurun_out/s35_bench.json >> $S; tail -2 gpurun_out/s35_bench.err >> $S
tail -c 1500 $S

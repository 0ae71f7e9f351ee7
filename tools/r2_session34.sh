#!/bin/bash
# Session 34: the T5 attention kernel (pxa_t5_attn_d64_bf16) and the fused q|k|v GEMM: kernel + encoder tests, t5 bench with the
# native and the PyTorch attention core.  The new kernel runs first under a short timeout.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/s34_summary.txt
echo "=== pytest t5 attention kernel (short timeout)" > $S
timeout 150 python -m pytest tests/test_t5_gpu.py -q -m gpu -p no:cacheprovider -s -k "attention_kernel" > gpurun_out/s34_kernel.log 2>&1
KRC=$?
echo "rc=$KRC $(tail -1 gpurun_out/s34_kernel.log)" >> $S
grep -E "^FAILED|^ERROR|Error|assert |rel_err" gpurun_out/s34_kernel.log | head -30 >> $S
if [ $KRC -eq 0 ]; then SEL=""; else SEL="-k torch"; fi
echo "=== pytest tests/test_t5_gpu.py $SEL" >> $S
timeout 400 python -m pytest tests/test_t5_gpu.py -q -m gpu -p no:cacheprovider -s $SEL --deselect tests/test_t5_gpu.py::test_t5_attention_kernel_matches_torch > gpurun_out/s34_tests.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/s34_tests.log)" >> $S
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/s34_tests.log | head -40 >> $S
grep -o "T5 [a-z0-9-]* attn=[a-z]*: last_hidden_state[^,]*, [0-9]* kernel launches" gpurun_out/s34_tests.log >> $S
if [ $KRC -eq 0 ]; then
  echo "=== bench t5 (native attention)" >> $S
  timeout 240 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s34_bench_t5.json 2> gpurun_out/s34_bench_t5.err
  echo "rc=$?" >> $S; cat gpurun_out/s34_bench_t5.json >> $S; tail -3 gpurun_out/s34_bench_t5.err >> $S
fi
echo "=== bench t5 (PyTorch attention core, fused q|k|v GEMM)" >> $S
PXA_T5_ATTN=torch timeout 240 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s34_bench_t5_torch.json 2> gpurun_out/s34_bench_t5_torch.err
echo "rc=$?" >> $S; cat gpurun_out/s34_bench_t5_torch.json >> $S; tail -3 gpurun_out/s34_bench_t5_torch.err >> $S
tail -c 3000 $S

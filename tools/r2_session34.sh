#!/bin/bash
# Session 34: the T5 attention kernel (pxa_t5_attn_d64_bf16) and the fused q|k|v GEMM: kernel + encoder tests, t5 bench with the
# native and the PyTorch attention core.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/s34_summary.txt
echo "=== pytest tests/test_t5_gpu.py" > $S
timeout 500 python -m pytest tests/test_t5_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/s34_tests.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/s34_tests.log)" >> $S
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/s34_tests.log | head -40 >> $S
grep -o "T5 [a-z0-9-]* attn=[a-z]*: last_hidden_state[^,]*, [0-9]* kernel launches" gpurun_out/s34_tests.log >> $S
echo "=== bench t5 (native attention)" >> $S
timeout 300 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s34_bench_t5.json 2> gpurun_out/s34_bench_t5.err
echo "rc=$?" >> $S; cat gpurun_out/s34_bench_t5.json >> $S; tail -3 gpurun_out/s34_bench_t5.err >> $S
echo "=== bench t5 (PyTorch attention core)" >> $S
PXA_T5_ATTN=torch timeout 300 python bench.py --workload t5 --steps 5 --warmup 3 > gpurun_out/s34_bench_t5_torch.json 2> gpurun_out/s34_bench_t5_torch.err
echo "rc=$?" >> $S; cat gpurun_out/s34_bench_t5_torch.json >> $S; tail -3 gpurun_out/s34_bench_t5_torch.err >> $S
tail -c 2500 $S

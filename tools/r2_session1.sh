#!/bin/bash
# Round-2 GPU session 1: full GPU suite (incl. the bench-geometry parity tests and the experimental GEMM epilogues),
# attention-backward experiment switches, L2-resident MLP grouping experiment, fused-MLP training op.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s1_summary.txt
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $S 2>&1
nproc >> $S
echo "=== full gpu suite (PXA_EXPERIMENTAL=1)" >> $S
PXA_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2s1_suite.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s1_suite.log)" >> $S
echo "=== mlp grouping" >> $S
timeout 300 python tools/mlp_group_bench.py >> $S 2>&1
echo "=== attn bwd variants" >> $S
timeout 900 bash tools/attn_bwd_variants.sh > gpurun_out/r2s1_bwdvar.log 2>&1
cat gpurun_out/attn_bwd_variants.txt >> $S
echo "=== fused MLP training op" >> $S
PXA_EXPERIMENTAL_FUSED_MLP=1 timeout 600 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s1_fusedmlp.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s1_fusedmlp.log)" >> $S
for fm in 0 1; do
  PXA_EXPERIMENTAL_FUSED_MLP=$fm timeout 600 python bench.py --workload c5 --cuda-graph --no-checkpoint --no-cpu-baseline > gpurun_out/r2s1_c5_fm$fm.json 2> gpurun_out/r2s1_c5_fm$fm.err
  echo "c5 graph nockpt fused_mlp=$fm rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s1_c5_fm$fm.json'));print(d['ms_per_step'], d['value'])" 2>&1)" >> $S
done
cat $S

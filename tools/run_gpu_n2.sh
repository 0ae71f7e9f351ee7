#!/bin/bash
# 2-GPU check of the batch-sharded bench path (run with `gpurun --gpus 2`). Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n2.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 rc=$?"; cat gpurun_out/bench_n2.json | cut -c1-300; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
echo "n2 ref rc=$?"; cat gpurun_out/bench_n2_ref.json | cut -c1-200

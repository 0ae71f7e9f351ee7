#!/bin/bash
# Session 37 (2 GPUs): the default bench line of the final build launched exactly as the driver's scaling run does at N = 2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s37_bench_n2.json 2> gpurun_out/s37_bench_n2.err
echo "rc=$?"; tail -2 gpurun_out/s37_bench_n2.err | cut -c1-300
python - <<'P'
import json
d = json.loads(open('gpurun_out/s37_bench_n2.json').read().strip().splitlines()[-1])
print('c3', d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])
print('train', d['train']['ms_per_step'], d['train']['value'], d['train']['allreduce'], d['train']['config']['fp32_attention'])
print('c4', d['c4']['ms_per_step'], d['c4']['value'])
P

#!/bin/bash
# Round-2 GPU session (2 GPUs): the c5 training step under DDP -- eager (bucket all-reduces behind the backward), one CUDA graph
# with the all-reduce after the replay, one CUDA graph with the collectives captured inside; fp32 and bf16 gradients on the wire.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2n2_summary.txt
: > $S
run() {  # tag, extra args...
  local tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --workload c5 --no-cpu-baseline "$@" > gpurun_out/r2n2_$tag.json 2> gpurun_out/r2n2_$tag.err
  echo "$tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2n2_$tag.json'));print(d['ms_per_step'], d['value'], d['config']['step_mode'], d['allreduce'], d['config']['parallelism'])" 2>&1 | cut -c1-400)" >> $S
  tail -2 gpurun_out/r2n2_$tag.err | cut -c1-300 >> $S
}
run graph --train-mode graph
run graph_bf16 --train-mode graph --bf16-reduce
run overlap --train-mode graph-overlap
run overlap_bf16 --train-mode graph-overlap --bf16-reduce
run eager --train-mode eager
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 > gpurun_out/r2n2_default.json 2> gpurun_out/r2n2_default.err
echo "default rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2n2_default.json'));print(d['value'], d['train']['ms_per_step'], d['train']['value'], d['train']['allreduce'], d['c4']['value'])" 2>&1 | cut -c1-400)" >> $S
cat $S

#!/bin/bash
# Round-2 GPU session 23 (2 GPUs): the default bench line and the c5 line under torchrun, as the driver launches them.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s23_summary.txt
: > $S
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2s23_n2_default.json 2> gpurun_out/r2s23_n2_default.err
echo "default N=2 rc=$?" >> $S
python - >> $S <<'PY'
import json
d = json.load(open("gpurun_out/r2s23_n2_default.json"))
print("c3", d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
for k in ("train", "c4"):
    e = d.get(k) or {}
    print(k, e.get("ms_per_step"), e.get("value"), e.get("allreduce"))
PY
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r2s23_n2_ref.json 2> gpurun_out/r2s23_n2_ref.err
echo "reference arm N=2 rc=$? $(cut -c1-200 gpurun_out/r2s23_n2_ref.json)" >> $S
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --no-checkpoint --bf16-reduce > gpurun_out/r2s23_n2_c5_bf16.json 2> gpurun_out/r2s23_n2_c5_bf16.err
echo "c5 no-ckpt bf16-reduce N=2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2s23_n2_c5_bf16.json'));print(round(d['ms_per_step'],2), round(d['value'],2), d.get('allreduce'))" 2>&1)" >> $S
cat $S

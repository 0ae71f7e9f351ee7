#!/bin/bash
# Secondary bench lines: sampling loop (eager / CUDA graph), the other BASELINE configs.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sampling-loop > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "loop rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sampling-loop --cuda-graph > gpurun_out/bench_loop_graph.json 2> gpurun_out/bench_loop_graph.err; echo "loop graph rc=$?"
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --sampling-loop > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --sampling-loop --cuda-graph > gpurun_out/bench_c2_graph.json 2> gpurun_out/bench_c2_graph.err; echo "c2 graph rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
for f in bench_loop bench_loop_graph bench_c2 bench_c2_graph bench_c4; do echo "== $f"; cut -c1-260 gpurun_out/$f.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read())
    print("   ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "loop", d.get("sampling_loop"))
except Exception as e:
    print("   parse error", e)
PY
done

#!/bin/bash
# Session 36: ncu launch list of the c3 step of the final build (round-end evidence r2g).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 400 --csv \
    --log-file gpurun_out/launches_r2g.csv python bench.py --no-cuda-graph --no-extras --no-parity --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s36_ncu_launch.log 2>&1
echo "rc=$? $(wc -l < gpurun_out/launches_r2g.csv) lines"

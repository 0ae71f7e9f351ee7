#!/bin/bash
# Round-2 GPU session 13: L2 reduce-add throughput micro-benchmark, qk_norm training test, DRAM bytes of the fused Mlp with a 27 MB ring.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s13_summary.txt
: > $S
echo "=== red_bw" >> $S
timeout -k 10 120 tools/micro/red_bw >> $S 2>&1
echo "=== qk_norm training test" >> $S
timeout -k 10 300 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k "qk_norm" > gpurun_out/r2s13_qk.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s13_qk.log)" >> $S
grep -E "AssertionError|assert |^FAILED|Error" gpurun_out/r2s13_qk.log | head -6 | cut -c1-600 >> $S
grep "qk_norm" gpurun_out/parity.txt | tail -1 >> $S
for cfg in "4 1 3 3" "2 4 6 3"; do
  echo "=== DRAM bytes, fused Mlp group/lag/ring/ksplit = $cfg" >> $S
  timeout -k 10 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2s13_mlp_dram.csv -k regex:"gemm|mlp" python tools/mlp_dram.py $cfg > gpurun_out/r2s13_ncu.log 2>&1
  echo "ncu rc=$?" >> $S
  python - >> $S <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2s13_mlp_dram.csv") if l.startswith('"'))]
h = rows[0]; ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]][:40]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
for (i, k), m in d.items():
    print(f"{i:3d} {k:42s} read {m.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  write {m.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB  {m.get('gpu__time_duration.sum', 0) / 1e3:8.1f} us")
PY
done
cat $S

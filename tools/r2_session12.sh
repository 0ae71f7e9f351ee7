#!/bin/bash
# Round-2 GPU session 12: fused Mlp with fc2 split along K (equal-cost tiles) -- parity, timing, DRAM bytes, c3 step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s12_summary.txt
: > $S
echo "=== fused mlp tests" >> $S
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "mlp_fused" > gpurun_out/r2s12_mlp.log 2>&1
echo "rc=$? $(tail -1 gpurun_out/r2s12_mlp.log)" >> $S
grep -E "^FAILED|^ERROR|rel_err|assert" gpurun_out/r2s12_mlp.log | head -8 >> $S
echo "=== timing: two launches vs one persistent kernel (M = 32768)" >> $S
timeout -k 10 300 python tools/mlp_fused_bench.py > gpurun_out/r2s12_bench.txt 2>&1
cat gpurun_out/r2s12_bench.txt >> $S
# best configuration by time -> DRAM bytes under ncu and the c3 step
BEST=$(grep "one persistent" gpurun_out/r2s12_bench.txt | sed -E 's/.*group ([0-9]+) lag ([0-9]+) ring ([0-9]+) fc2 k-splits ([0-9]+).*: +([0-9.]+) us.*/\5 \1 \2 \3 \4/' | sort -n | head -1 | cut -d' ' -f2-)
echo "best: $BEST" >> $S
timeout -k 10 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2s12_mlp_dram.csv -k regex:"gemm|mlp" python tools/mlp_dram.py $BEST > gpurun_out/r2s12_ncu.log 2>&1
echo "ncu rc=$?" >> $S
python - >> $S <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2s12_mlp_dram.csv") if l.startswith('"'))]
h = rows[0]; ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]][:40]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
for (i, k), m in d.items():
    print(f"{i:3d} {k:42s} read {m.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  write {m.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB  {m.get('gpu__time_duration.sum', 0) / 1e3:8.1f} us")
PY
set -- $BEST
for f in 0 1; do
  PXA_MLP_GROUP=$1 PXA_MLP_LAG=$2 PXA_MLP_RING=$3 PXA_MLP_FUSED=$f timeout -k 10 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2s12_c3_mlp$f.json 2> gpurun_out/r2s12_c3_mlp$f.err
  echo "c3 mlp_fused=$f rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s12_c3_mlp$f.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['block_rel_err'], d['clocks'])" 2>&1)" >> $S
done
cat $S

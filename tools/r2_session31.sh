#!/bin/bash
# Round-2 GPU session 31: persistent attention backward (both passes) -- parity, isolated timing, c5 step A/B vs one CTA per item.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/r2s31_summary.txt
: > $S
V=$PWD/pixart_sigma_b200/build/variants
timeout -k 10 400 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "attn" > gpurun_out/r2s31_t.log 2>&1
echo "backward attention tests (persistent) rc=$? $(tail -1 gpurun_out/r2s31_t.log)" >> $S
grep -E "^FAILED|^ERROR" gpurun_out/r2s31_t.log | head -8 >> $S
PXA_LIB_PATH=$V/libpixart_sm100_bwdnp.so timeout -k 10 400 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -k "attn" > gpurun_out/r2s31_t1.log 2>&1
echo "backward attention tests (one CTA per item) rc=$? $(tail -1 gpurun_out/r2s31_t1.log)" >> $S
for rep in 1 2; do
echo "--- persistent" >> $S
timeout -k 10 200 python tools/attn_bwd_bench.py >> $S 2>&1
echo "--- one CTA per item" >> $S
PXA_LIB_PATH=$V/libpixart_sm100_bwdnp.so timeout -k 10 200 python tools/attn_bwd_bench.py >> $S 2>&1
done
timeout -k 10 500 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2s31_t2.log 2>&1
echo "training tests rc=$? $(tail -1 gpurun_out/r2s31_t2.log)" >> $S
for tag in persistent np persistent2; do
  if [ $tag == np ]; then export PXA_LIB_PATH=$V/libpixart_sm100_bwdnp.so; else unset PXA_LIB_PATH; fi
  timeout -k 10 500 python bench.py --workload c5 --no-checkpoint --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2s31_c5_$tag.json 2> gpurun_out/r2s31_c5_$tag.err
  echo "c5 no-ckpt $tag rc=$?: $(python -c "import json;d=json.load(open('gpurun_out/r2s31_c5_$tag.json'));print(round(d['ms_per_step'],2), round(d['value'],2), d['roofline']['whole_step']['frac'], d['clocks']['sm_mhz'])" 2>&1)" >> $S
done
unset PXA_LIB_PATH
cat $S

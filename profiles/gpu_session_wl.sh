#!/bin/bash
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
echo "== pytest -m gpu" > gpurun_out/session_wl.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 >> gpurun_out/session_wl.log
for wl in term_top10_1M_1seg and2_top10_10M_1seg; do
  echo "== $wl" >> gpurun_out/session_wl.log
  timeout 900 python bench.py --workload $wl --steps 6 --warmup 3 >> gpurun_out/session_wl.log 2>&1
done
echo "== or20_top10_500M_64seg" >> gpurun_out/session_wl.log
timeout 900 python bench.py --workload or20_top10_500M_64seg --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/session_wl.log 2>&1
echo "== default" >> gpurun_out/session_wl.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline >> gpurun_out/session_wl.log 2>&1
tail -c 2500 gpurun_out/session_wl.log

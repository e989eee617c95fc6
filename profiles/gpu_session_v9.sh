#!/bin/bash
# N=1 lease: GPU tests and the bench lines of every BASELINE workload + batch sizes 1 / 64 / 1024 of the default workload.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_v9.log
echo "== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -30 >> $L
echo "== default bench" >> $L
timeout 600 python bench.py --steps 12 --warmup 3 >> $L 2>&1
for nq in 1 64 1024; do
  echo "== default workload, nq=$nq" >> $L
  timeout 600 python bench.py --nq $nq --steps 12 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
for wl in term_top10_1M_1seg and2_top10_10M_1seg mixed_top10_100M_8seg; do
  echo "== $wl" >> $L
  timeout 900 python bench.py --workload $wl --steps 8 --warmup 3 >> $L 2>&1
done
echo "== or20_top10_500M_64seg" >> $L
timeout 900 python bench.py --workload or20_top10_500M_64seg --steps 3 --warmup 3 --cpu-sample 32 >> $L 2>&1
grep -o '"value": [0-9.]*, "unit": "queries/s", "n_gpus": [0-9]*\|"workload": "[a-z0-9_A-Z]*"\|"queries_per_step": [0-9]*\|"e2e": {"value": [0-9.]*\|"mismatches": [0-9]*' $L | paste - - - - - | tail -12

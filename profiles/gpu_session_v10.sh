#!/bin/bash
# N=1 lease: GPU tests; k_phrase next to the tile engine (second stream) on / off; single-term queries on k_term vs the tile engine.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_v10.log
echo "== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -30 >> $L
for cfg in "TQ_X=0" "TQ_PHRASE_SIDE_STREAM=0" "TQ_TILE_OPS=6" ; do
  echo "== mixed, $cfg" >> $L
  env $cfg timeout 600 python bench.py --workload mixed_top10_100M_8seg --steps 8 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
for cfg in "TQ_X=0" "TQ_TILE_OPS=6" ; do
  echo "== term, $cfg" >> $L
  env $cfg timeout 600 python bench.py --workload term_top10_1M_1seg --steps 12 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
grep -v '^{"metric' $L | tail -12
grep -o '"value": [0-9.]*, "unit": "queries/s", "n_gpus": [0-9]*\|"workload": "[a-z0-9_A-Z]*"\|"e2e": {"value": [0-9.]*\|"mismatches": [0-9]*\|"phrase": [0-9.]*\|"term": [0-9.]*, "and"' $L | paste - - - - - - | tail -6

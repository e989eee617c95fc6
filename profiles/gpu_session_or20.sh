#!/bin/bash
# N=1 lease: configs[4] (20-term unions) on one GPU under different group sizes / window policies.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_or20.log
: > $L
for cfg in "TQ_X=0" "TQ_TILE_MAX_QUERIES=256" "TQ_TILE_MAX_QUERIES=128" "TQ_TILE_MAX_QUERIES=64" "TQ_TILE_WINDOWS=8 TQ_TILE_LIGHT_MAX=24" "TQ_TILE_MAX_QUERIES=128 TQ_TILE_WINDOWS=8 TQ_TILE_LIGHT_MAX=24" "TQ_TILE_MAX_QUERIES=128 TQ_TILE_LIGHT_MAX=100000"; do
  echo "-- $cfg" >> $L
  env $cfg TQ_TILE_COUNTERS=1 timeout 300 python profiles/probe_tile.py or20_top10_500M_64seg 512 1 0 2>&1 | grep '"step": 2' | cut -c1-520 >> $L
done
echo "== pytest tile hooks" >> $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tile" 2>&1 | tail -3 >> $L
cat $L

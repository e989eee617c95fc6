#!/bin/bash
# One GPU lease, several measurements; everything lands in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
echo "== sweeps" > gpurun_out/session.log
for lm in 8 16 32 64 192; do
  echo "-- TQ_TILE_LIGHT_MAX=$lm" >> gpurun_out/session.log
  TQ_TILE_LIGHT_MAX=$lm timeout 300 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 2>&1 | grep '"step": 2' | cut -c1-420 >> gpurun_out/session.log
done
for u in 444 1776 3552; do
  echo "-- TQ_TILE_UNITS=$u" >> gpurun_out/session.log
  TQ_TILE_UNITS=$u timeout 300 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 2>&1 | grep '"step": 2' | cut -c1-420 >> gpurun_out/session.log
done
echo "== ncu full k_tile + k_score_lists" >> gpurun_out/session.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_score_lists" --launch-skip 5 --launch-count 5 -f -o gpurun_out/prof_tile_r2b python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 >> gpurun_out/session.log 2>&1
tail -c 4000 gpurun_out/session.log

#!/bin/bash
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
echo "== probe (tile engine, or5 100M)" > gpurun_out/session.log
timeout 600 python profiles/probe_tile.py or5_top100_100M_8seg 512 2 16 2>&1 | cut -c1-460 >> gpurun_out/session.log
echo "== pytest -m gpu" >> gpurun_out/session.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -30 >> gpurun_out/session.log
echo "== sweeps" >> gpurun_out/session.log
for cfg in "TQ_TILE_SAMPLE_DIV=8" "TQ_TILE_SAMPLE_DIV=4" "TQ_TILE_SAMPLE_DIV=32" "TQ_TILE_BIG_MIN=6" "TQ_TILE_UNITS=1332" "TQ_TILE_LIGHT_MAX=192"; do
  echo "-- $cfg" >> gpurun_out/session.log
  env $cfg timeout 300 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 2>&1 | grep '"step": 2' | cut -c1-330 >> gpurun_out/session.log
done
echo "== launch list" >> gpurun_out/session.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/launches_r2_tile.csv python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 > /dev/null 2>&1
grep -E "k_tile|k_score|k_theta|k_final" gpurun_out/launches_r2_tile.csv | awk -F'","' '{print $5, $(NF)}' | tail -9 >> gpurun_out/session.log
echo "== ncu full, exact launch C + k_score_lists" >> gpurun_out/session.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_score_lists" --launch-skip 8 --launch-count 2 -f -o gpurun_out/prof_tile_r2c python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 > /dev/null 2>&1
tail -c 3500 gpurun_out/session.log

#!/bin/bash
# One GPU lease, several measurements; everything lands in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
# the snapshot may have been taken between an edit and its rebuild: make the libraries match the sources
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
echo "== probe (tile engine, or5 100M)" > gpurun_out/session.log
timeout 600 python profiles/probe_tile.py or5_top100_100M_8seg 512 3 16 >> gpurun_out/session.log 2>&1
echo "== launch list" >> gpurun_out/session.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2_tile.csv python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 > /dev/null 2>&1
grep -E "k_tile|k_score|k_theta|k_final" gpurun_out/launches_r2_tile.csv | awk -F'","' '{print $5, $(NF)}' | head -12 >> gpurun_out/session.log
echo "== pytest -m gpu" >> gpurun_out/session.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -40 >> gpurun_out/session.log
echo "== bench" >> gpurun_out/session.log
timeout 600 python bench.py --steps 8 --warmup 3 >> gpurun_out/session.log 2>&1
tail -c 3000 gpurun_out/session.log

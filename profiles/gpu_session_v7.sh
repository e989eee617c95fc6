#!/bin/bash
# N=1 lease: GPU tests, the default bench line, counters of configs[4] on one GPU, the launch list of the bench command and a
# full ncu capture of the last exact k_tile launch + k_score_lists.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_v7.log
echo "== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -30 >> $L
echo "== default bench" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 >> $L 2>&1
echo "== probe or5" >> $L
timeout 600 python profiles/probe_tile.py or5_top100_100M_8seg 512 2 0 2>&1 | cut -c1-460 >> $L
echo "== probe or20" >> $L
timeout 600 python profiles/probe_tile.py or20_top10_500M_64seg 512 1 0 2>&1 | cut -c1-460 >> $L
echo "== launch list of: python bench.py --steps 2 --warmup 1 --no-cpu-baseline" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-queries 0 > gpurun_out/bench_under_ncu.log 2>&1
grep -c "k_tile" gpurun_out/launches_r2_bench.csv >> $L
echo "== ncu full: k_tile (last exact launch) + k_score_lists" >> $L
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_score_lists" --launch-skip 8 --launch-count 2 -f -o gpurun_out/prof_tile_v7 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep >> $L
tail -c 3000 $L

#!/bin/bash
# N=1 lease: GPU tests, default bench line, threshold-potential probe of configs[4], launch list of the bench command with DRAM bytes.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_v8.log
echo "== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -30 >> $L
echo "== default bench" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 >> $L 2>&1
echo "== probe or20 (+ thresholds imported)" >> $L
PROBE_THETA=1 timeout 600 python profiles/probe_tile.py or20_top10_500M_64seg 512 1 0 2>&1 | cut -c1-520 >> $L
echo "== probe or5 (+ thresholds imported)" >> $L
PROBE_THETA=1 timeout 600 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 2>&1 | cut -c1-520 >> $L
echo "== launch list + DRAM bytes of: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-queries 0" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-queries 0 > gpurun_out/bench_under_ncu.log 2>&1
grep -c "k_tile" gpurun_out/launches_r2_bench_dram.csv >> $L
tail -c 3000 $L

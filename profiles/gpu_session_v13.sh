#!/bin/bash
# N=1 lease, final code of the round: GPU suite, bench lines of every BASELINE workload + the reference arm, ncu launch list of
# the bench command (time + DRAM bytes per launch), ncu --set full of k_score_lists + the four k_tile launches of one step.
mkdir -p gpurun_out
L=gpurun_out/session_v13.log
echo "== pytest -m gpu" > $L
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -8 ) >> $L 2>&1
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> $L
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/r2_bench_n1.json 2>> $L
echo "default exit=$?" >> $L
for wl in and2_top10_10M_1seg term_top10_1M_1seg mixed_top10_100M_8seg; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_${wl}_n1.json 2>> $L
  echo "$wl exit=$?" >> $L
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> $L
echo "reference exit=$?" >> $L
echo "== ncu launch list of the bench command" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_or5_100M_v9.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 0 > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit=$?" >> $L
echo "== ncu full: k_score_lists + the four k_tile launches of one step" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_score_lists" --launch-skip 10 --launch-count 5 -f -o gpurun_out/prof_tile_v9 python profiles/probe_tile.py or5_top100_100M_8seg 512 1 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep >> $L
grep -v "^\*\*\*\|OMP_NUM\|^$" $L | tail -30
for f in gpurun_out/r2_bench_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if d.get("impl") == "reference":
        print(sys.argv[1], "reference value", round(d["value"], 1), d["cpu_baseline"]["cores"], "threads")
    else:
        print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "serial", round(d["pipeline"]["serial_value"]),
              "parity", d.get("parity", {}).get("mismatches"), "cpu", round(d.get("cpu_baseline", {}).get("value", 0), 1), "kern", d["roofline"]["all_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

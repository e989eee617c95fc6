#!/bin/bash
# N=1 lease: full GPU suite (doc-range split, tq_multi split, async pack), default workload with two batches in flight vs one,
# and2 (configs[1]) line.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_v11.log
echo "== pytest -m gpu" > $L
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 ) >> $L 2>&1
echo "== default workload, 2 in flight" >> $L
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_or5_n1.json 2>> $L
echo "exit=$?" >> $L
echo "== default workload, 1 in flight" >> $L
timeout 600 python bench.py --steps 12 --warmup 3 --in-flight 1 --no-cpu-baseline --parity-queries 0 > gpurun_out/bench_or5_n1_serial.json 2>> $L
echo "exit=$?" >> $L
echo "== and2" >> $L
timeout 600 python bench.py --workload and2_top10_10M_1seg --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_and2_n1.json 2>> $L
echo "exit=$?" >> $L
grep -v "^\*\*\*\|OMP_NUM\|^$" $L | tail -60
for f in gpurun_out/bench_or5_n1.json gpurun_out/bench_or5_n1_serial.json gpurun_out/bench_and2_n1.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "serial", round(d["pipeline"]["serial_value"]),
          "e2e_serial", round(d["e2e"]["serial_value"]), "parity", d.get("parity", {}).get("mismatches"), "kern", d["roofline"]["all_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

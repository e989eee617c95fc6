#!/bin/bash
# N=1 lease: A/B of k_tile stage B2 (early test + survivor compaction in _lib vs the previous flat path in _lib_base), tile counters, GPU suite.
mkdir -p gpurun_out
L=gpurun_out/session_v12.log
echo "== pytest -m gpu" > $L
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -15 ) >> $L 2>&1
for lib in _lib _lib_base; do
  echo "== probe $lib" >> $L
  TANTIVY_B200_LIB=$PWD/tantivy_b200/$lib/libtantivy_b200.so timeout 300 python profiles/probe_tile.py or5_top100_100M_8seg 512 6 16 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -5 >> $L
done
for lib in _lib _lib_base; do
  echo "== bench $lib" >> $L
  TANTIVY_B200_LIB=$PWD/tantivy_b200/$lib/libtantivy_b200.so timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_or5_$lib.json 2>> $L
done
for wl in mixed_top10_100M_8seg and2_top10_10M_1seg; do
  for lib in _lib _lib_base; do
    TANTIVY_B200_LIB=$PWD/tantivy_b200/$lib/libtantivy_b200.so timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_$lib.json 2>> $L
  done
done
grep -v "^\*\*\*\|OMP_NUM\|^$" $L | tail -45
for f in gpurun_out/bench_*_lib*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "serial", round(d["pipeline"]["serial_value"]),
          "parity", d.get("parity", {}).get("mismatches"), "kern", d["roofline"]["all_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

#!/bin/bash
# Two-GPU lease: tq_multi tests on two devices, N=2 lines with two batches in flight (default, and2 = one segment split by doc range,
# mixed), the forced-overflow path of CrossGpuMerger.complete, the reference arm under torchrun.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_n2d.log
echo "== pytest multi_handle" > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_handle" 2>&1 | tail -5 >> $L
run() {  # name, extra env, args...
  name=$1; shift; envs=$1; shift
  echo "== $name" >> $L
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 "$@" > gpurun_out/$name.json 2>> $L
  echo "exit=$?" >> $L
}
run bench_or5_n2 TQ_X=0 --steps 12 --warmup 3
run bench_and2_n2 TQ_X=0 --workload and2_top10_10M_1seg --steps 12 --warmup 3
run bench_mixed_n2 TQ_X=0 --workload mixed_top10_100M_8seg --steps 8 --warmup 3
run bench_or5_n2_overflow TQ_TILE_CAND_FLOOR=4 --steps 3 --warmup 3 --docs-per-segment 1000000
grep -v "^\*\*\*\|OMP_NUM\|^$\|^W0\|^\[W" $L | tail -40
for f in bench_or5_n2 bench_and2_n2 bench_mixed_n2 bench_or5_n2_overflow; do
  python - "gpurun_out/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "serial", round(d["pipeline"]["serial_value"]),
          "e2e_serial", round(d["e2e"]["serial_value"]), "parity", d.get("parity", {}).get("mismatches"), "repeats", d["pipeline"]["overflow_repeats"],
          "fallback_steps", d["workload_stats"]["tile_fallback_steps"], "kern", d["roofline"]["all_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done

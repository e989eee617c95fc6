#!/bin/bash
# Eight-GPU lease: the scaling points of the default workload and BASELINE.json configs[3] / configs[4] over 8 ranks.
N=${1:-8}
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
: > gpurun_out/session_n8.log
run() {  # n workload extra...
  local n=$1; shift; local wl=$1; shift
  echo "== $wl N=$n" >> gpurun_out/session_n8.log
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 --workload $wl "$@" >> gpurun_out/session_n8.log 2>&1
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload $wl "$@" >> gpurun_out/session_n8.log 2>&1; fi
}
run $N or5_top100_100M_8seg --steps 12 --warmup 3
run 4 or5_top100_100M_8seg --steps 12 --warmup 3
run 2 or5_top100_100M_8seg --steps 12 --warmup 3
run 1 or5_top100_100M_8seg --steps 12 --warmup 3 --no-cpu-baseline
run $N mixed_top10_100M_8seg --steps 6 --warmup 3
run $N or20_top10_500M_64seg --steps 4 --warmup 3
grep -o '"value": [0-9.]*, "unit": "queries/s", "n_gpus": [0-9]*\|"workload": "[a-z0-9_A-Z]*"\|"e2e": {"value": [0-9.]*\|"mismatches": [0-9]*' gpurun_out/session_n8.log | paste - - - - | tail -8

#!/bin/bash
# Two-GPU lease: full GPU suite, term / mixed lines at N=1 with the default routing, clean exit + lines of the N=2 runs.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_n2c.log
echo "== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -8 >> $L
for wl in term_top10_1M_1seg mixed_top10_100M_8seg; do
  echo "== $wl N=1" >> $L
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 >> $L 2>&1
  echo "exit=$?" >> $L
done
for wl in or5_top100_100M_8seg mixed_top10_100M_8seg; do
  echo "== $wl N=2" >> $L
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 >> $L 2>&1
  echo "exit=$?" >> $L
done
echo "== reference arm under torchrun N=2" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> $L 2>&1
echo "exit=$?" >> $L
grep -v '^{"metric\|^{"impl' $L | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -25
grep -o '"value": [0-9.]*, "unit": "queries/s", "n_gpus": [0-9]*\|"workload": "[a-z0-9_A-Z]*"\|"e2e": {"value": [0-9.]*\|"mismatches": [0-9]*' $L | paste - - - - | tail -6

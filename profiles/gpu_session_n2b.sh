#!/bin/bash
# Two-GPU lease: the cross-GPU path after the fetch / key-count changes, and the phrase kernel after its verification rework.
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
L=gpurun_out/session_n2b.log
echo "== pytest phrase + multi" > $L
timeout 900 python -m pytest tests/test_gpu_phrase.py tests/test_gpu_parity.py -m gpu -q -k "phrase or multi or topkeys or mixed" 2>&1 | tail -5 >> $L
echo "== mixed N=1" >> $L
timeout 600 python bench.py --workload mixed_top10_100M_8seg --steps 8 --warmup 3 --no-cpu-baseline >> $L 2>&1
echo "== or5 N=2, 30 keys per rank in the exchange" >> $L
TANTIVY_B200_EXCHANGE_KEYS=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 >> $L 2>&1
echo "== or5 N=2" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 >> $L 2>&1
echo "== mixed N=2" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload mixed_top10_100M_8seg --steps 8 --warmup 3 >> $L 2>&1
grep -v '^{"metric' $L | tail -20
grep -o '"value": [0-9.]*, "unit": "queries/s", "n_gpus": [0-9]*\|"workload": "[a-z0-9_A-Z]*"\|"e2e": {"value": [0-9.]*\|"mismatches": [0-9]*\|"phrase": [0-9.]*' $L | paste - - - - - | tail -6

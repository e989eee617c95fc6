#!/bin/bash
# Two-GPU lease: the in-process multi handle on two devices, and the sharded bench (NCCL key exchange + packed all-gather).
mkdir -p gpurun_out
(make -C tantivy_b200/csrc -s 2>&1 | grep -E "error|Error" ; make -C oracle -s 2>&1 | grep -E "error|Error") > gpurun_out/build.log 2>&1
echo "== multi handle, two devices" > gpurun_out/session_n2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_handle" 2>&1 | tail -15 >> gpurun_out/session_n2.log
echo "== pytest -m gpu (all)" >> gpurun_out/session_n2.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -25 >> gpurun_out/session_n2.log
echo "== bench N=2" >> gpurun_out/session_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 >> gpurun_out/session_n2.log 2>&1
echo "== bench N=1 (same box)" >> gpurun_out/session_n2.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline >> gpurun_out/session_n2.log 2>&1
tail -c 6000 gpurun_out/session_n2.log

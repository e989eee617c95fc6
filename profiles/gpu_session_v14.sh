#!/bin/bash
# N=1 lease: CTAs per exact launch (TQ_TILE_UNITS) with two batches in flight -- the launch tails no longer idle the GPU, so longer
# units (fewer cursor prologues) may pay now.
mkdir -p gpurun_out
for u in 444 666 888 1332; do
  TQ_TILE_UNITS=$u timeout 200 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --parity-queries 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('units $u value', round(d['value']), 'serial', round(d['pipeline']['serial_value']), 'e2e', round(d['e2e']['value']), d['roofline']['all_kernels_ms_per_step'])"
done

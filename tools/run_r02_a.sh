#!/bin/bash
# r02 call A: reference digest at 1024^3 on the box's host cores, then GPU tests and the bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g > gpurun_out/mem.txt
echo "== golden 1024 (reference OpenMP JFA on host)"
( time timeout 1500 python tests/golden/make_golden_large.py --only bunny1348128_union_bimba_n1024 ) 2>&1 | tail -8
cp tests/golden/ref_digests_large.json gpurun_out/ref_digests_large.json
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json

#!/bin/bash
# r02: two-row flood form at 3 CTAs per SM (24 warps) for strides <= 8 against the shipped four-row form (12 warps)
set -u
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["parity"]["status"])'
echo "== variant r2c3 (two rows per thread, 3 CTAs/SM for k <= 8)"
VPB_LIB=$PWD/gpurun_variants/libvpb200_r2c3.so VPB_F5_RPT=2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --config4 off 2>/dev/null | python -c "$show" | tee gpurun_out/r02_r2c3.txt
echo "== shipped"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --config4 off 2>/dev/null | python -c "$show" | tee -a gpurun_out/r02_r2c3.txt

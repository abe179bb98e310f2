#!/bin/bash
# r02 call V (1 GPU): packed vs scalar candidate arithmetic in the flood pass (stride 1 only / every stride)
set -u
mkdir -p gpurun_out
for V in packed scalar1 scalar_all; do
  export VPB_LIB=$PWD/gpurun_variants/libvpb200_$V.so
  echo "== $V"
  timeout 600 python -m pytest tests -m gpu -x -q -k "metric_config_1024 or config3_512 or tiled_pass or random_grids" 2>&1 | tail -1
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config4 off 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$V', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['roofline']['ms_early_seed_plus_3_passes'], d['parity']['status'])" | tee -a gpurun_out/r02_scalar_ab.txt
done

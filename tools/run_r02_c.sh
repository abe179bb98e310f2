#!/bin/bash
# r02 call C: flood5 with arithmetic decode at k >= 16: variants timing + ncu --set full of two variants
set -u
mkdir -p gpurun_out
export VPB_VARIANT_TESTS="stage_calls or fused_early or oracle or random or metric_config"
timeout 1500 python tools/variants.py run 1024 2>&1 | tail -12
V=cuda_mesh_voxelization_b200/build/variants
for name in f5_r2 f5_r4_c3; do
  VPB_LIB=$V/libvpb200_$name.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jfa_pass_flood5|jfa_early' -c 8 -f -o gpurun_out/r02_$name \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep

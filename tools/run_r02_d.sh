#!/bin/bash
# r02 call D: ncu --set full of flood5 variants + jfa_early; summaries made ON the box (the reports are too big to bring back)
set -u
mkdir -p gpurun_out /tmp/ncu
V=cuda_mesh_voxelization_b200/build/variants
for name in f5_r2 f5_r4_c3; do
  VPB_LIB=$V/libvpb200_$name.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jfa_pass_flood5|jfa_early' -c 8 -f -o /tmp/ncu/$name \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep > gpurun_out/r02_${name}_summary.txt 2>&1
  for li in 0 1 6 7; do python tools/ncu_sass_costs.py /tmp/ncu/$name.ncu-rep $li > gpurun_out/r02_${name}_costs_launch$li.txt 2>&1; done
  python tools/ncu_hot_lines.py /tmp/ncu/$name.ncu-rep 6 40 > gpurun_out/r02_${name}_hot_k2.txt 2>&1
  python tools/ncu_hot_lines.py /tmp/ncu/$name.ncu-rep 0 40 > gpurun_out/r02_${name}_hot_early.txt 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page details --launch-skip 6 --launch-count 1 > gpurun_out/r02_${name}_details_k2.txt 2>&1
done
ncu -i /tmp/ncu/f5_r2.ncu-rep --page details --launch-skip 0 --launch-count 1 > gpurun_out/r02_early_details.txt 2>&1
cat gpurun_out/r02_f5_r2_summary.txt gpurun_out/r02_f5_r4_c3_summary.txt
du -sh gpurun_out

#!/bin/bash
# r02 final: ncu --set full of the JFA kernels of one step of the shipped library, summarised on the box
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jfa_pass_flood5|jfa_early' -c 8 -f -o /tmp/ncu/jfa python bench.py --steps 1 --warmup 0 --no-cpu-baseline --config4 off > gpurun_out/ncu_jfa.log 2>&1
python tools/ncu_summary.py /tmp/ncu/jfa.ncu-rep > gpurun_out/r02_final_jfa_summary.txt 2>&1; cut -c1-250 gpurun_out/r02_final_jfa_summary.txt
python tools/ncu_sass_costs.py /tmp/ncu/jfa.ncu-rep 6 > gpurun_out/r02_final_costs_k2.txt 2>&1; head -12 gpurun_out/r02_final_costs_k2.txt
python tools/ncu_sass_costs.py /tmp/ncu/jfa.ncu-rep 0 > gpurun_out/r02_final_costs_early.txt 2>&1; head -5 gpurun_out/r02_final_costs_early.txt

#!/bin/bash
# r02 call J: full GPU suite with the fused CSG+shell path, bench, vox/CSG probe + ncu of those kernels, ncu records of the bench
set -u
mkdir -p gpurun_out /tmp/ncu
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('ms/step %.2f value %.2f e2e %.2f parity %s launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['status'], d['gpu_launches']))
print('early', d['roofline']['ms_early_seed_plus_3_passes'], {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, 'frac', round(d['roofline']['frac'],4))
print(d['cpu_baseline'])"
echo "== vox/csg probe"; timeout 600 python tools/vox_csg_probe.py 2>&1 | tail -6 | tee gpurun_out/r02_vox_csg_probe.txt
timeout 900 ncu --set full --clock-control none -k regex:'vox_|csg_|surf_|shell_' -c 30 -f -o /tmp/ncu/vox python tools/vox_csg_probe.py > gpurun_out/ncu_vox.log 2>&1
ncu -i /tmp/ncu/vox.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); H=r[0]; U=r[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct']
idx=[H.index(w) for w in want if w in H]
print(' | '.join(H[i] for i in idx)); print(' | '.join(U[i] for i in idx))
for row in r[2:]: print(' | '.join(row[i][:60] for i in idx))
" > gpurun_out/r02_vox_csg_ncu.txt 2>&1
head -40 gpurun_out/r02_vox_csg_ncu.txt | cut -c1-260
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-150
echo "== ncu full of the JFA kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jfa_pass_flood5|jfa_early' -c 8 -f -o /tmp/ncu/jfa python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_jfa.log 2>&1
python tools/ncu_summary.py /tmp/ncu/jfa.ncu-rep > gpurun_out/r02_final_jfa_summary.txt 2>&1; cat gpurun_out/r02_final_jfa_summary.txt | cut -c1-250
for li in 0 1 4 6 7; do python tools/ncu_sass_costs.py /tmp/ncu/jfa.ncu-rep $li > gpurun_out/r02_final_costs_launch$li.txt 2>&1; done
python tools/ncu_hot_lines.py /tmp/ncu/jfa.ncu-rep 6 40 > gpurun_out/r02_final_hot_k2.txt 2>&1
python tools/ncu_hot_lines.py /tmp/ncu/jfa.ncu-rep 1 40 > gpurun_out/r02_final_hot_k64.txt 2>&1
python tools/ncu_hot_lines.py /tmp/ncu/jfa.ncu-rep 7 40 > gpurun_out/r02_final_hot_k1.txt 2>&1
du -sh gpurun_out

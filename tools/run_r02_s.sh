#!/bin/bash
# r02 call S (1 GPU): full GPU suite, the bench line with the config-4 extra run, ncu launch list, ncu --set full of the CSG / shell kernels
set -u
mkdir -p gpurun_out /tmp/ncu
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_s_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('ms/step %.2f value %.2f e2e %.2f parity %s launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['status'], d['gpu_launches']))
print('early', d['roofline']['ms_early_seed_plus_3_passes'], {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, 'frac', round(d['roofline']['frac'],4))
print(d['config'].get('extra_runs'))"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config4 off > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-150
echo "== ncu full of csg / shell kernels"
timeout 900 ncu --set full --clock-control none -k regex:'csg_|shell_' -c 12 -f -o /tmp/ncu/csg python tools/vox_csg_probe.py 1348128 > gpurun_out/ncu_csg.log 2>&1
ncu -i /tmp/ncu/csg.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); H=r[0]; U=r[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
idx=[H.index(w) for w in want if w in H]
print(' | '.join(H[i] for i in idx)); print(' | '.join(U[i] for i in idx))
for row in r[2:]: print(' | '.join(row[i][:60] for i in idx))
" > gpurun_out/r02_csg_shell_ncu.txt 2>&1
cat gpurun_out/r02_csg_shell_ncu.txt | cut -c1-260

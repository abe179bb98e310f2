#!/bin/bash
# r02 final 1-GPU call: full GPU suite, the bench line, ncu launch list
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_final_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_1gpu_final.json').read().strip().splitlines()[-1])
print('ms/step %.2f value %.2f e2e %.2f parity %s launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['status'], d['gpu_launches']))
print('early', d['roofline']['ms_early_seed_plus_3_passes'], {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, 'frac', round(d['roofline']['frac'],4))
print(d['cpu_baseline']); print(d['config'].get('extra_runs')); print(d['clocks'])"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config4 off > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-150

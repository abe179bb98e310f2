#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_early or stage_calls or metric_config or slab" 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f'%d['ms_per_step'], 'early', round(d['roofline']['ms_early_seed_plus_3_passes'] or 0,2), d['parity']['status'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jfa_early' -c 1 -f -o /tmp/ncu/early python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_early.log 2>&1
python tools/ncu_summary.py /tmp/ncu/early.ncu-rep > gpurun_out/r02_early2_summary.txt 2>&1
python tools/ncu_sass_costs.py /tmp/ncu/early.ncu-rep 0 > gpurun_out/r02_early2_costs.txt 2>&1
ncu -i /tmp/ncu/early.ncu-rep --page details > gpurun_out/r02_early2_details.txt 2>&1
ncu -i /tmp/ncu/early.ncu-rep --page source --csv --print-source sass | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=None; out=[]
for r in rows:
    if r and r[0]=='Address': H=r; continue
    if H and len(r)>=len(H) and r[0].startswith('0x'):
        try: w=int(r[H.index('L1 Wavefronts Shared')]); e=int(r[H.index('L1 Wavefronts Shared Excessive')]); n=int(r[H.index('Instructions Executed')])
        except ValueError: continue
        if w: out.append((w,e,n,r[1].strip()))
tot=sum(o[0] for o in out)
print('total shared wavefronts',tot,'excessive',sum(o[1] for o in out))
for w,e,n,s in sorted(out,reverse=True)[:40]: print(f'{w:12d} {e:12d} {n:10d} {s[:90]}')
" > gpurun_out/r02_early2_wavefronts.txt 2>&1
cat gpurun_out/r02_early2_summary.txt; head -45 gpurun_out/r02_early2_wavefronts.txt

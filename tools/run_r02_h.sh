#!/bin/bash
# r02 call H (8 GPUs): bench with the 2048^3 extra run + digests, then the concurrent D2H probe
set -u
mkdir -p gpurun_out
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${G}gpu.json 2> gpurun_out/r02_bench_${G}gpu.err
tail -2 gpurun_out/r02_bench_${G}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${G}gpu.json'))
print('ms/step %.2f value %.2f e2e %s parity %s' % (d['ms_per_step'], d['value'], d['e2e'] and round(d['e2e']['value'],2), d['parity']))
print('stages', d['config'].get('stage_ms_by_rank'))
print('passes', d['roofline']['ms_per_pass_by_k'], 'early', d['roofline']['ms_early_seed_plus_3_passes'])
for e in d['config'].get('extra_runs', []): print('extra', e['n'], e['csg'], 'ms/step %.2f value %.2f' % (e['ms_per_step'], e['value']), e['stage_ms_by_rank'])
PY
timeout 300 $TR tools/d2h_probe.py 550 > gpurun_out/r02_d2h_probe_${G}gpu.txt 2>&1; head -30 gpurun_out/r02_d2h_probe_${G}gpu.txt

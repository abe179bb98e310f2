#!/bin/bash
# r02 call I (G GPUs): halo-overlap A/B + multi-GPU parity check
set -u
mkdir -p gpurun_out
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multi_gpu_check.py 2>/dev/null | grep -v NCCL | tail -3
for ov in 1 0; do
  VPB_HALO_OVERLAP=$ov timeout 600 $TR bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --extra-2048 off 2>/dev/null | tail -1 > gpurun_out/r02_bench_${G}gpu_ov$ov.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${G}gpu_ov$ov.json').read().strip().splitlines()[-1])
print('overlap=$ov ms/step %.2f value %.2f e2e %s parity %s' % (d['ms_per_step'], d['value'], d['e2e'] and round(d['e2e']['value'],2), d['parity']['status']))
for k,v in d['config'].get('stage_ms_by_rank',{}).items(): print('  ', k, v)
print('   passes', {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()})
PY
done

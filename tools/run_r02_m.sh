#!/bin/bash
# r02 call M (G GPUs): where the parity-split passes wait; side-stream priority A/B
G=${1:-4}
mkdir -p gpurun_out
for P in 1 0; do
VPB_SPLIT_TRACE=1 VPB_SIDE_PRIORITY=$P VPB_HALO=split timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$G bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --no-cpu-baseline 2>gpurun_out/bench${G}_prio$P.err | tee gpurun_out/r02_bench_${G}gpu_split_prio$P.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('prio=$P', d['ms_per_step'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d.get('parity'))"
grep trace gpurun_out/bench${G}_prio$P.err | sort | head -4 | cut -c1-400
done

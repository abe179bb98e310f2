#!/bin/bash
# r02 call L (G GPUs): split-mode halo exchange (two half-slab launches per pass, flags instead of barriers) against the push mode
G=${1:-2}
mkdir -p gpurun_out
echo "== multi_gpu_check ($G ranks, split)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>gpurun_out/multi_check_$G.err | tee gpurun_out/r02_multi_check_${G}_split.txt
tail -3 gpurun_out/multi_check_$G.err | cut -c1-300
for H in split push; do
VPB_HALO=$H timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$G bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off 2>gpurun_out/bench${G}_$H.err | tee gpurun_out/r02_bench_${G}gpu_$H.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('halo=$H', d['ms_per_step'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['roofline']['ms_early_seed_plus_3_passes'], d['e2e']['value'], d.get('parity'))"
tail -2 gpurun_out/bench${G}_$H.err | cut -c1-300
done

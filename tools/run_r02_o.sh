#!/bin/bash
# r02 call O (G GPUs): z-cyclic first phase (VPB_CYCLIC=1) against the slab-only split mode (VPB_CYCLIC=0)
G=${1:-4}
mkdir -p gpurun_out
for C in 1 0; do
VPB_CYCLIC=$C VPB_HALO=split timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$G bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --no-cpu-baseline 2>gpurun_out/bench${G}_cyc$C.err | tee gpurun_out/r02_bench_${G}gpu_cyclic$C.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cyclic=$C', d['ms_per_step'], d['value'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['roofline']['ms_early_seed_plus_3_passes'], d['e2e']['value'], d.get('parity'))"
tail -2 gpurun_out/bench${G}_cyc$C.err | cut -c1-300
done

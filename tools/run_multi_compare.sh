#!/bin/bash
# usage (on a G-GPU box): tools/run_multi_compare.sh G  -> parity check of the slab pipeline, then bench with the DMA halo pull and with NCCL send/recv
G=${1:-2}
mkdir -p gpurun_out
echo "== slab/early parity tests on one GPU"; timeout 600 python -m pytest tests -m gpu -x -q -k "slab or early" 2>&1 | tail -3
echo "== multi_gpu_check ($G ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>gpurun_out/multi_check_$G.err | tee gpurun_out/multi_check_$G.txt
tail -3 gpurun_out/multi_check_$G.err | cut -c1-300
for H in dma nccl; do
VPB_HALO=$H timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$G bench.py --gpus $G --steps 3 --warmup 3 2>gpurun_out/bench${G}_$H.err | tee gpurun_out/bench${G}_$H.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('halo=$H', d['ms_per_step'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['roofline']['ms_early_seed_plus_3_passes'], d['e2e'])"
tail -2 gpurun_out/bench${G}_$H.err | cut -c1-300
done

#!/bin/bash
# usage (on a G-GPU box): tools/run_halo_ab.sh G "modes" [extra bench args] -> parity check, then one bench run per halo transport mode
G=${1:-8}; MODES=${2:-"push pull"}; EXTRA=${3:-}
mkdir -p gpurun_out
echo "== multi_gpu_check ($G ranks, default transport)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>gpurun_out/multi_check_$G.err | tee gpurun_out/multi_check_$G.txt
tail -2 gpurun_out/multi_check_$G.err | cut -c1-300
for H in $MODES; do
VPB_HALO=$H timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline $EXTRA 2>gpurun_out/bench${G}_$H.err | grep '^{' | tee gpurun_out/bench${G}_$H.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('halo=$H', d['ms_per_step'], d['value'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['roofline']['ms_early_seed_plus_3_passes'], d['e2e'])"
grep -i "error\|unavailable" gpurun_out/bench${G}_$H.err | head -3 | cut -c1-300
done

for P in 1 0; do
VPB_PEER=$P python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$P bench.py --gpus 4 --steps 3 --warmup 3 2>gpurun_out/bench4_$P.err | tee gpurun_out/bench4_peer$P.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('peer=$P', d['ms_per_step'], d['config'].get('stage_ms_by_rank'), {k:round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()})"
done

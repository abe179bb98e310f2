#!/bin/bash
# r02 call U (G GPUs): parity check + bench with the occupancy pushed through symmetric memory (default) and with the NCCL all-gather
G=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>gpurun_out/multi_check_$G.err | tee gpurun_out/r02_multi_check_${G}_final.txt
tail -2 gpurun_out/multi_check_$G.err | cut -c1-300
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["config"].get("stage_ms_by_rank"), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["roofline"]["ms_early_seed_plus_3_passes"], d["e2e"]["value"], d.get("parity")); print(json.dumps(d["config"].get("extra_runs", []))[:1200])'
for M in push nccl; do
echo "== gather=$M"
VPB_GATHER=$M timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off 2>gpurun_out/bench${G}_$M.err | tee gpurun_out/r02_bench_${G}gpu_gather_$M.json | python -c "$show"
tail -2 gpurun_out/bench${G}_$M.err | cut -c1-300
done

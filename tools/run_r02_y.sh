#!/bin/bash
# r02 call Y (G GPUs): the driver's command line with the final defaults, then the all-direct transpose variant
G=${1:-8}
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["config"].get("stage_ms_by_rank"), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["roofline"]["ms_early_seed_plus_3_passes"], "e2e", d["e2e"]["value"], d.get("parity",{}).get("status")); print(json.dumps(d["config"].get("extra_runs", []))[:1800])'
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $G --steps 5 --warmup 3 2>gpurun_out/bench${G}_final.err | tee gpurun_out/r02_bench_${G}gpu_final.json | python -c "$show"
tail -1 gpurun_out/bench${G}_final.err | cut -c1-300
echo "== transpose=direct"
VPB_TRANSPOSE=direct timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --config4 off 2>gpurun_out/bench${G}_direct.err | tee gpurun_out/r02_bench_${G}gpu_transpose_direct.json | python -c "$show"
tail -1 gpurun_out/bench${G}_direct.err | cut -c1-300

#!/bin/bash
G=${1:-4}
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["config"].get("stage_ms_by_rank"), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["roofline"]["ms_early_seed_plus_3_passes"], "e2e", d["e2e"]["value"], d.get("parity",{}).get("status")); print(json.dumps(d["config"].get("extra_runs", []))[:600])'
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>gpurun_out/multi_check_$G.err | tee gpurun_out/r02_multi_check_${G}_final.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $G --steps 5 --warmup 3 2>gpurun_out/bench${G}_final.err | tee gpurun_out/r02_bench_${G}gpu_final.json | python -c "$show"
tail -1 gpurun_out/bench${G}_final.err | cut -c1-300

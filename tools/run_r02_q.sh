#!/bin/bash
# r02 call Q (8 GPUs): the driver's command line (z-cyclic + parity-split default, 2048^3 extra run), then the slab-only split mode
G=${1:-8}
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["config"].get("stage_ms_by_rank"), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["roofline"]["ms_early_seed_plus_3_passes"], d["e2e"]["value"], d.get("parity"), [ (x["ms_per_step"], x["value"]) for x in d["config"].get("extra_runs", [])])'
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $G --steps 5 --warmup 3 2>gpurun_out/bench${G}_default.err | tee gpurun_out/r02_bench_${G}gpu_default.json | python -c "$show"
tail -2 gpurun_out/bench${G}_default.err | cut -c1-300
VPB_CYCLIC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --no-cpu-baseline 2>gpurun_out/bench${G}_cyc0.err | tee gpurun_out/r02_bench_${G}gpu_cyclic0.json | python -c "$show"

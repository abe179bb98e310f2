#!/bin/bash
# r02 call X (G GPUs): transpose variants of the z-cyclic phase (copy engines / own slab by kernel stores / all by kernel stores)
G=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cyclic" 2>&1 | tail -2
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], d["config"].get("stage_ms_by_rank"), {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["roofline"]["ms_early_seed_plus_3_passes"], "e2e", d["e2e"]["value"], d.get("parity",{}).get("status"))'
for T in copy own direct; do
echo "== transpose=$T"
VPB_CYCLIC=1 VPB_TRANSPOSE=$T timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --config4 off 2>gpurun_out/bench${G}_$T.err | tee gpurun_out/r02_bench_${G}gpu_transpose_$T.json | python -c "$show"
tail -1 gpurun_out/bench${G}_$T.err | cut -c1-300
done

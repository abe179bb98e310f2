#!/bin/bash
# r02 (8 GPUs): the parts of the last cyclic pass on one stream against two alternating streams
G=${1:-8}
mkdir -p gpurun_out
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["value"], {k:round(v,2) for k,v in d["roofline"]["ms_per_pass_by_k"].items()}, d["config"].get("stage_ms_by_rank",{}).get("flood"), d.get("parity",{}).get("status"))'
for S in 2 1; do
echo "== streams=$S"
VPB_CYCLIC_STREAMS=$S timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2953$S bench.py --gpus $G --steps 5 --warmup 3 --extra-2048 off --config4 off 2>gpurun_out/bench${G}_s$S.err | tee gpurun_out/r02_bench_${G}gpu_streams$S.json | python -c "$show"
tail -1 gpurun_out/bench${G}_s$S.err | cut -c1-200
done

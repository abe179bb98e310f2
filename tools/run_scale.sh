#!/bin/bash
# usage (on the GPU box): tools/run_scale.sh N_GPUS "extra bench args" tag   -> gpurun_out/bench_<tag>_<N>gpu.json
G=$1; EXTRA=$2; TAG=$3
mkdir -p gpurun_out
if [ "$G" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_${TAG}_${G}gpu.json 2> gpurun_out/bench_${TAG}_${G}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_${TAG}_${G}gpu.json 2> gpurun_out/bench_${TAG}_${G}gpu.err
fi
tail -3 gpurun_out/bench_${TAG}_${G}gpu.err | cut -c1-400
cat gpurun_out/bench_${TAG}_${G}gpu.json | cut -c1-3000

#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, GPU parity tests, bench, and the ncu launch list.
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
echo "== bench" ; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
fi
if [ "${NCUFULL:-0}" = "1" ]; then
echo "== ncu --set full of the fused early kernel + the flood passes"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:jfa_early\|jfa_pass_flood -c ${NCUFULL_COUNT:-8} -f -o gpurun_out/flood_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi

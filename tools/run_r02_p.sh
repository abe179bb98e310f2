#!/bin/bash
set -u
G=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cyclic or slab" 2>&1 | tail -5
bash tools/run_r02_o.sh $G

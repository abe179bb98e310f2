#!/bin/bash
# r02 call K: GPU suite with the marching CSG+shell kernel, vox/CSG probe
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k_pytest.txt
echo "== vox/csg probe"; timeout 600 python tools/vox_csg_probe.py 2>&1 | tail -8 | tee gpurun_out/r02_vox_csg_probe.txt

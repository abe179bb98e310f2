#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cyclic or slab or wide_state_z" 2>&1 | tail -15

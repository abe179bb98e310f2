#!/bin/bash
set -u
make -C apps/cli 2>&1 | tail -1
timeout 900 python -m pytest tests/test_cli_gpu.py -x -q 2>&1 | tail -15

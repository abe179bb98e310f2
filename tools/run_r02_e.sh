#!/bin/bash
set -u
mkdir -p gpurun_out
export VPB_VARIANT_TESTS="stage_calls or fused_early or oracle or random or metric_config"
timeout 1500 python tools/variants.py run 1024 2>&1 | tail -12

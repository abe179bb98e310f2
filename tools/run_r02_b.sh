#!/bin/bash
# r02 call B: flood5 (TMA-staged) bring-up: parity tests + per-pass timing of three variants, flood4 for reference
set -u
mkdir -p gpurun_out
export VPB_VARIANT_TESTS="not cli and not benchmarks_runner and not 2048 and not config4"
echo "== flood4 reference timing"
VPB_JFA_KERNEL=flood4 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flood4 ms/step %.2f'%d['ms_per_step'], d['roofline']['ms_per_pass_by_k'], d['parity']['status'])"
echo "== variants"
timeout 2400 python tools/variants.py run 1024 2>&1 | tail -60

#!/bin/bash
# r02 call F: full GPU suite on the new default library (flood5 dual form + early v2), bench, A/B of the forms
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms/step %.2f value %.2f e2e %.2f parity %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['status']))
print('early', d['roofline']['ms_early_seed_plus_3_passes'], 'passes', {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, 'frac', d['roofline']['frac'])
PY
for v in "VPB_JFA_KERNEL=flood4" "VPB_F5_RPT=2" "VPB_F5_RPT=4"; do
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ms/step %.2f'%d['ms_per_step'], 'early', round(d['roofline']['ms_early_seed_plus_3_passes'] or 0,2), {k: round(v,2) for k,v in d['roofline']['ms_per_pass_by_k'].items()}, d['parity']['status'])"
done

#!/bin/bash
# A/B timing of environment switches on the GPU box: tools/ab_env.sh WORKLOAD TIME_STEPS "VAR=val ..." ...   (ION_LIB=... selects a variant build)
wl=$1; nt=$2; shift 2
for v in "$@"; do
  env $v python bench.py --workload $wl --steps 3 --warmup 3 --time-steps $nt --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('[$v]', '$wl', 'us/step %.2f' % d['us_per_time_step'], 'frac %.3f' % d['hbm_roofline_frac_step'], 'parity', (d.get('parity') or {}).get('max_rel_err'), 'kernels', d['roofline']['kernels_us'])
except Exception as e:
    print('[$v]', '$wl', 'FAILED', e)"
done

#!/bin/bash
# A/B timing on the GPU box: tools/ab.sh WORKLOAD TIME_STEPS LIB...   (LIB = "main" or a variant name built by tools/build_variant.sh)
wl=$1; nt=$2; shift 2
for v in "$@"; do
  if [ "$v" = main ]; then unset ION_LIB; else export ION_LIB=$PWD/ionization_b200/_lib/exp_$v.so; fi
  python bench.py --workload $wl --steps 5 --warmup 3 --time-steps $nt --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', '$wl', 'us/step %.2f' % d['us_per_time_step'], 'kernels', d['roofline']['kernels_us'])"
done
